# fused sort + diffuse shade (k_sort<0>): A/B through the env switch, then the whole GPU suite with the fused path as default
mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
L=gpu-raytracer_b200/csrc
python tools/gpu_variants_r2.py "{\"fused\": \"$L/libptb.so\", \"unfused\": \"$L/libptb.so+PTB_FUSE_SORT_SHADE=0\", \"fused2\": \"$L/libptb.so\"}" 1,0 2>&1 | tee $O/variants.log
timeout 1200 python -m pytest tests -m gpu -q -x > $O/all_gpu_tests.log 2>&1; tail -15 $O/all_gpu_tests.log
