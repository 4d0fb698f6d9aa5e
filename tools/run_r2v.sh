# occluder cache of the merged tree (shadow rays): A/B through the env switch, then the whole GPU suite
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
L=gpu-raytracer_b200/csrc
python tools/gpu_variants_r2.py "{\"occ_on\": \"$L/libptb.so\", \"occ_off\": \"$L/libptb.so+PTB_OCCLUDER_CACHE=0\", \"occ_on2\": \"$L/libptb.so\"}" 1 2>&1 | tee $O/variants.log
timeout 1200 python -m pytest tests -m gpu -q -x > $O/all_gpu_tests.log 2>&1; tail -15 $O/all_gpu_tests.log
