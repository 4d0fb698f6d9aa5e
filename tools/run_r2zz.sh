# last check of HEAD: smoke(), the parity file of the GPU suite, the default bench line
mkdir -p gpurun_out/r2zz
O=gpurun_out/r2zz
python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 300 python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2zz/bench_default.json").read().splitlines() if l.startswith("{")][-1])
print("bench", round(d["value"],1), d["unit"], round(d["ms_per_step"],3), "ms e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "clocks", d["clocks"])
PY
