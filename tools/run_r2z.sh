# end-of-round verification of the final code: GPU suite, every BASELINE config with both arms, launch list + full ncu capture
mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q > $O/all_gpu_tests.log 2>&1; tail -4 $O/all_gpu_tests.log
echo "tests done at $(( $(date +%s) - S )) s"
for c in 1 0 2 3 4; do
  timeout 300 python bench.py --config $c > $O/bench_config$c.json 2> $O/bench_config$c.err
  timeout 400 python bench.py --impl reference --config $c > $O/bench_reference_config$c.json 2> $O/bench_reference_config$c.err
  python - $c <<'PY'
import json, sys
c = sys.argv[1]
def load(p):
    try: return json.loads([l for l in open(p).read().splitlines() if l.startswith("{")][-1])
    except Exception as e: return {"error": str(e)}
a, b = load(f"gpurun_out/r2z/bench_config{c}.json"), load(f"gpurun_out/r2z/bench_reference_config{c}.json")
print("config", c, a.get("metric"), "ptb", round(a.get("value", 0), 1), "ms", round(a.get("ms_per_step", 0), 3), "e2e", round((a.get("e2e") or {}).get("value", 0), 1),
      "| reference", round(b.get("value", 0), 1), "e2e", round((b.get("e2e") or {}).get("value", 0), 1), "| strict", round((a.get("two_level_only") or {}).get("value", 0), 1))
PY
done
echo "bench done at $(( $(date +%s) - S )) s"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_config1.csv python bench.py --config 1 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; tail -2 $O/ncu_bench.log
PTB_PROF_WAVE=9 timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/prof_all python tools/prof_all.py > $O/ncu_full.log 2>&1; tail -3 $O/ncu_full.log
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > $O/prof_all_raw.csv 2> $O/export.err
python tools/ncu_summarize.py $O/prof_all_raw.csv $O/r2_final
gzip -9 $O/prof_all_raw.csv
echo "all done at $(( $(date +%s) - S )) s"; du -sh $O
