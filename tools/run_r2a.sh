mkdir -p gpurun_out/r2a
bash tools/run_r2b.sh
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r2a/configs_tests.log 2>&1; echo "configs rc=$?" >> gpurun_out/r2a/configs_tests.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_configs.py > gpurun_out/r2a/other_tests.log 2>&1; echo "other rc=$?" >> gpurun_out/r2a/other_tests.log
for c in 0 1 2 3 4; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2a/bench_c$c.json 2> gpurun_out/r2a/bench_c$c.err
  timeout 300 python bench.py --impl reference --config $c --steps 3 --warmup 3 > gpurun_out/r2a/bench_ref_c$c.json 2> gpurun_out/r2a/bench_ref_c$c.err
done
tail -3 gpurun_out/r2a/*.log; tail -c 400 gpurun_out/r2a/bench_c*.err
