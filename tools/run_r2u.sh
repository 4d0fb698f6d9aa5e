mkdir -p gpurun_out/r2u
timeout 600 python tools/gpu_bvh_kinds.py 2>&1 | tee gpurun_out/r2u/bvh_kinds.log
