# 2-GPU box: the IPC data plane, the tile-local SVGF across processes, bench lines at N = 2
mkdir -p gpurun_out/r2f
nvidia-smi -L > gpurun_out/r2f/gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -k "between_processes" > gpurun_out/r2f/ipc_test.log 2>&1; tail -3 gpurun_out/r2f/ipc_test.log
for c in 1 3 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --config $c --steps 4 --warmup 3 > gpurun_out/r2f/bench_2gpu_c$c.json 2> gpurun_out/r2f/bench_2gpu_c$c.err
  tail -c 600 gpurun_out/r2f/bench_2gpu_c$c.err
  python - "$c" <<'PY'
import json,sys
c=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/r2f/bench_2gpu_c{c}.json")); print("config",c,"N=2 value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),"gather_check",d.get("gather_check"),"stages",d.get("stage_ms_per_step"))
except Exception as e: print("config",c,"no result",e)
PY
done
