mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -x -k "refit or instance_updates" > $O/refit_tests.log 2>&1; tail -25 $O/refit_tests.log
timeout 300 python bench.py --config 1 --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_c1.json 2> $O/bench_c1.err; tail -c 300 $O/bench_c1.err
python - <<'PY'
import json
txt=open("gpurun_out/r2l/bench_c1.json").read(); d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); print("bench_c1", round(d["value"],1), round(d["ms_per_step"],3), {k:round(v,2) for k,v in d["stage_ms_per_step"].items()})
PY
timeout 1500 python -m pytest tests -m gpu -q > $O/all_gpu_tests.log 2>&1; tail -8 $O/all_gpu_tests.log
