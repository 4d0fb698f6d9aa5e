# 8-GPU box: BASELINE configs [1], [3], [4] at N = 8 and at N = 1 on the same box (scaling of the configs that are written as 8-GPU)
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
nvidia-smi -L > $O/gpus.txt 2>&1
for c in 1 4 3; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$c bench.py --gpus 8 --config $c --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_8gpu_c$c.json 2> $O/bench_8gpu_c$c.err
  tail -c 300 $O/bench_8gpu_c$c.err
  timeout 300 python bench.py --gpus 1 --config $c --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_1gpu_c$c.json 2> $O/bench_1gpu_c$c.err
  python - "$c" <<'PY'
import json,sys
c=sys.argv[1]
def load(p):
    txt=open(p).read(); return json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
try:
    a=load(f"gpurun_out/r2q/bench_8gpu_c{c}.json"); b=load(f"gpurun_out/r2q/bench_1gpu_c{c}.json")
    print("config",c,"N=8",round(a["value"],1),"ms",round(a["ms_per_step"],3),"| N=1",round(b["value"],1),"ms",round(b["ms_per_step"],3),"| speed-up",round(a["value"]/b["value"],2),"efficiency",round(a["value"]/b["value"]/8,3),"gather_check",a.get("gather_check"),"stages N=8",{k:round(v,2) for k,v in a.get("stage_ms_per_step",{}).items()})
except Exception as e: print("config",c,"no result",e)
PY
done
