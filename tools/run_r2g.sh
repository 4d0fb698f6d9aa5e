mkdir -p gpurun_out/r2g
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2g/all_gpu_tests.log 2>&1; tail -15 gpurun_out/r2g/all_gpu_tests.log
for c in 1 2; do timeout 300 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2g/bench_c$c.json 2> gpurun_out/r2g/bench_c$c.err; tail -c 300 gpurun_out/r2g/bench_c$c.err; done
PTB_SVGF_TMA=0 timeout 300 python bench.py --config 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2g/bench_c2_notma.json 2> gpurun_out/r2g/bench_c2_notma.err
python - <<'PY'
import json
for n in ("bench_c1","bench_c2","bench_c2_notma"):
    try:
        txt=open(f"gpurun_out/r2g/{n}.json").read(); d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); print(n, round(d["value"],1), round(d["ms_per_step"],3), {k:round(v,2) for k,v in d["stage_ms_per_step"].items()})
    except Exception as e: print(n, "no result", e)
PY
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g/launches_bench_config1.csv python bench.py --config 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2g/ncu_bench.log 2>&1; tail -2 gpurun_out/r2g/ncu_bench.log
# full capture: the 8 trace launches of one 9-pass wave frame + every kernel of one SVGF + TAA frame
PTB_PROF_WAVE=9 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2g/prof_full python tools/prof_all.py > gpurun_out/r2g/ncu_full.log 2>&1; tail -3 gpurun_out/r2g/ncu_full.log
ls -la gpurun_out/r2g
