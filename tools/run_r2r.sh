# session 3 of round 2: full GPU suite with durations (how long does the driver's run take?), stream-kernel A/Bs, default bench both arms
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
L=gpu-raytracer_b200/csrc
S=$(date +%s)
python tools/gpu_variants_r2.py "{\"base\": \"$L/libptb.so\", \"sortprefetch\": \"$L/libptb_sp.so\", \"accbatch9\": \"$L/libptb_ab9.so\", \"sp_ab12\": \"$L/libptb_sp_ab12.so\", \"base2\": \"$L/libptb.so\"}" 1 2>&1 | tee $O/variants.log
echo "variants done at $(( $(date +%s) - S )) s"
timeout 1500 python -m pytest tests -m gpu -q --durations=30 > $O/all_gpu_tests.log 2>&1; tail -45 $O/all_gpu_tests.log
echo "tests done at $(( $(date +%s) - S )) s"
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 1500 $O/bench_default.json
echo "bench done at $(( $(date +%s) - S )) s"
timeout 300 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 600 $O/bench_reference.json
echo "all done at $(( $(date +%s) - S )) s"
