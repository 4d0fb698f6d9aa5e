# trace-kernel knobs: what is staged in shared memory (merged tree top instead of the TLAS when both exist; how many nodes), shadow-ray
# specific dynamic-fetch / postponing thresholds
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
L=gpu-raytracer_b200/csrc
python tools/gpu_variants_r2.py "{\"base\": \"$L/libptb.so\", \"stage_merged_top\": \"$L/libptb_smt.so\", \"stage_merged_top_128\": \"$L/libptb_smt128.so\", \"stage_merged_top_384\": \"$L/libptb_smt384.so\", \"stage_tlas_96\": \"$L/libptb_st96.so\", \
\"shadow_fetch_4_16\": \"$L/libptb_s416.so\", \"shadow_fetch_1_4\": \"$L/libptb_s14.so\", \"shadow_postpone_3\": \"$L/libptb_sp3.so\", \"shadow_postpone_8\": \"$L/libptb_sp8.so\", \"shade_prefetch\": \"$L/libptb_shp.so\", \"base2\": \"$L/libptb.so\"}" 1 2>&1 | tee $O/variants.log
