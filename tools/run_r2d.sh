mkdir -p gpurun_out/r2d
timeout 300 python tools/gpu_svgf_diag.py sponza 8 > gpurun_out/r2d/svgf_diag.log 2>&1; grep -c " ok" gpurun_out/r2d/svgf_diag.log; grep -v " ok" gpurun_out/r2d/svgf_diag.log | head -40
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_facade.py -m gpu -q -s > gpurun_out/r2d/configs_tests.log 2>&1; grep -E "^\[|passed|failed|FAILED|Error" gpurun_out/r2d/configs_tests.log | head -60
timeout 900 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2d/other_tests.log 2>&1; tail -30 gpurun_out/r2d/other_tests.log
for c in 2 4; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d/bench_c$c.json 2> gpurun_out/r2d/bench_c$c.err; tail -c 300 gpurun_out/r2d/bench_c$c.err; done
python - <<'PY'
import json
for c in (2,4):
    try:
        d=json.load(open(f"gpurun_out/r2d/bench_c{c}.json")); print(c, d["value"], d["ms_per_step"], d["stage_ms_per_step"])
    except Exception as e: print(c, "no result", e)
PY
