mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bvh4" > $O/bvh4_tests.log 2>&1; tail -25 $O/bvh4_tests.log
timeout 1500 python -m pytest tests -m gpu -q > $O/all_gpu_tests.log 2>&1; tail -8 $O/all_gpu_tests.log
