mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
L=gpu-raytracer_b200/csrc
python tools/gpu_variants_r2.py "{\"fast\": \"$L/libptb.so\", \"exactnodes\": \"$L/libptb.so+PTB_NODE_TEST_EXACT=1\", \"fetch_1_4\": \"$L/libptb_f14.so\", \"fetch_0_1\": \"$L/libptb_f01.so\", \"postpone3\": \"$L/libptb_p3.so\", \"postpone4\": \"$L/libptb_p4.so\", \"postpone8\": \"$L/libptb_p8.so\"}" 1 2>&1 | tee $O/variants.log
timeout 1500 python -m pytest tests -m gpu -q -x > $O/all_gpu_tests.log 2>&1; tail -15 $O/all_gpu_tests.log
