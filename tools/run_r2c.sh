mkdir -p gpurun_out/r2c
python tools/gpu_variants_r2.py '{"base":"gpu-raytracer_b200/csrc/libptb.so","half":"gpu-raytracer_b200/csrc/libptb_half.so"}' 1 > gpurun_out/r2c/variants.log 2>&1
cat gpurun_out/r2c/variants.log
timeout 300 python tools/gpu_svgf_diag.py sponza 8 > gpurun_out/r2c/svgf_diag.log 2>&1; tail -60 gpurun_out/r2c/svgf_diag.log
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_facade.py -m gpu -q > gpurun_out/r2c/configs_tests.log 2>&1; tail -40 gpurun_out/r2c/configs_tests.log
timeout 600 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2c/other_tests.log 2>&1; tail -30 gpurun_out/r2c/other_tests.log
