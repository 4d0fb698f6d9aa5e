# wave lanes with stream priorities: bit-identity test, rank probe with 1 and 2 lanes
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -q -x -k "wave_lanes or multi_pass" > $O/lanes_tests.log 2>&1; tail -5 $O/lanes_tests.log
timeout 300 python tools/gpu_rank_probe.py 1 2>&1 | tee $O/probe_lanes1.log
timeout 300 python tools/gpu_rank_probe.py 2 2>&1 | tee $O/probe_lanes2.log
