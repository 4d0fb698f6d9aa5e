# insertion-based optimisation of the merged tree: on (default) vs off, then the parity tests that trace through the merged tree
mkdir -p gpurun_out/r2ac
L=gpu-raytracer_b200/csrc
python tools/gpu_variants_r2.py "{\"optimized\": \"$L/libptb.so\", \"as_built\": \"$L/libptb.so+PTB_MERGE_OPTIMIZE=0\"}" 1 2>&1 | tee gpurun_out/r2ac/variants.log
timeout 100 python -m pytest tests -m gpu -q -x -k "static_merge_is_invisible or full_size_frame or config2 or woop or instance_updates or refit" > gpurun_out/r2ac/tests.log 2>&1; tail -4 gpurun_out/r2ac/tests.log
