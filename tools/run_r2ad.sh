# optimiser on: the rest of the parity tests that run in the default (merged) traversal mode
mkdir -p gpurun_out/r2ad
timeout 70 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_facade.py -m gpu -q -x -k "not (full_size_frame or config2 or config4 or woop)" > gpurun_out/r2ad/tests.log 2>&1; tail -4 gpurun_out/r2ad/tests.log
