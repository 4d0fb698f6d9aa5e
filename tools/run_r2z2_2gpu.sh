# 2-GPU box, final code: IPC data plane test, bench lines at N = 2 for configs 1, 3, 4 with the gather self-check
mkdir -p gpurun_out/r2z2
O=gpurun_out/r2z2
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -q -k "between_processes" > $O/ipc_test.log 2>&1; tail -3 $O/ipc_test.log
for c in 1 3 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$c bench.py --gpus 2 --config $c --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_2gpu_c$c.json 2> $O/bench_2gpu_c$c.err
  tail -c 300 $O/bench_2gpu_c$c.err
  python - "$c" <<'PY'
import json,sys
c=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/r2z2/bench_2gpu_c{c}.json").read().splitlines() if l.startswith("{")][-1]); print("config",c,"N=2 value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),"gather_check",d.get("gather_check"),"stages",{k:round(v,2) for k,v in (d.get("stage_ms_per_step") or {}).items()})
except Exception as e: print("config",c,"no result",e)
PY
done
