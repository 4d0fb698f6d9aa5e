mkdir -p gpurun_out/r2e
for m in nosync third sync; do DIAG_MODE=$m timeout 300 python tools/gpu_svgf_diag.py sponza 8 > gpurun_out/r2e/svgf_diag_$m.log 2>&1; echo "== $m"; grep -c " ok" gpurun_out/r2e/svgf_diag_$m.log; grep -v " ok" gpurun_out/r2e/svgf_diag_$m.log | head -24; done
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -k "config2 or config4" > gpurun_out/r2e/svgf_tests.log 2>&1; tail -5 gpurun_out/r2e/svgf_tests.log
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -k "lanes or svgf or merge" > gpurun_out/r2e/prop_tests.log 2>&1; tail -5 gpurun_out/r2e/prop_tests.log
python tools/gpu_rank_probe.py 1 > gpurun_out/r2e/rank_probe_lanes1.log 2>&1; cat gpurun_out/r2e/rank_probe_lanes1.log
python tools/gpu_rank_probe.py 2 > gpurun_out/r2e/rank_probe_lanes2.log 2>&1; cat gpurun_out/r2e/rank_probe_lanes2.log
