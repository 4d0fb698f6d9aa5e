# smaller trace CTAs (feasible at full occupancy now that a CTA stages 10 KB): do finer-grained CTA retirements shorten the tails of a 1/8 shard?
mkdir -p gpurun_out/r2aa
O=gpurun_out/r2aa
L=$PWD/gpu-raytracer_b200/csrc
for v in "" _b128 _b128s32 _b64; do
  echo "== libptb$v.so"
  PTB_LIB_PATH=$L/libptb$v.so timeout 200 python tools/gpu_rank_probe.py 2>&1 | tee $O/probe$v.log
done
