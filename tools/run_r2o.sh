mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
L=gpu-raytracer_b200/csrc
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -x -k "refit or instance_updates" > $O/refit_tests.log 2>&1; tail -5 $O/refit_tests.log
python tools/gpu_variants_r2.py "{\"sort4\": \"$L/libptb.so\", \"sort5\": \"$L/libptb_s5.so\", \"sort6\": \"$L/libptb_s6.so\", \"sort8\": \"$L/libptb_s8.so\"}" 1 2>&1 | tee $O/variants.log
timeout 1500 python -m pytest tests -m gpu -q > $O/all_gpu_tests.log 2>&1; tail -8 $O/all_gpu_tests.log
