mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -x -k "refit or instance_updates" > $O/refit_tests.log 2>&1; tail -25 $O/refit_tests.log
timeout 1500 python -m pytest tests -m gpu -q > $O/all_gpu_tests.log 2>&1; tail -8 $O/all_gpu_tests.log
