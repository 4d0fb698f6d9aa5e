# 4-GPU box, final code: bench lines at N = 4 for the headline and the 4K SVGF configuration
mkdir -p gpurun_out/r2z4
O=gpurun_out/r2z4
nvidia-smi -L > $O/gpus.txt 2>&1
for c in 1 4; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2962$c bench.py --gpus 4 --config $c --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_4gpu_c$c.json 2> $O/bench_4gpu_c$c.err
  python - "$c" <<'PY'
import json,sys
c=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/r2z4/bench_4gpu_c{c}.json").read().splitlines() if l.startswith("{")][-1]); print("config",c,"N=4 value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),"gather_check",(d.get("gather_check") or {}).get("equal"),"stages",{k:round(v,2) for k,v in (d.get("stage_ms_per_step") or {}).items()})
except Exception as e: print("config",c,"no result",e)
PY
done
