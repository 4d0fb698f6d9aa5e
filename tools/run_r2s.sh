# stream-kernel variants on the new defaults (sort prefetch 1, accumulate batch 9) + merged-tree builder parameters (env knobs)
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
L=gpu-raytracer_b200/csrc
python tools/gpu_variants_r2.py "{\"base\": \"$L/libptb.so\", \"sortprefetch2\": \"$L/libptb_sp2.so\", \"shadeprefetch\": \"$L/libptb_shp.so\", \
\"bins128\": \"$L/libptb.so+PTB_MERGE_BINS=128\", \
\"bins128_tc0.3\": \"$L/libptb.so+PTB_MERGE_BINS=128+PTB_MERGE_TRICOST=0.3\", \
\"bins128_tc0.1\": \"$L/libptb.so+PTB_MERGE_BINS=128+PTB_MERGE_TRICOST=0.1\", \
\"bins96_tc0.1\": \"$L/libptb.so+PTB_MERGE_TRICOST=0.1\", \
\"bins64_tc0.1\": \"$L/libptb.so+PTB_MERGE_BINS=64+PTB_MERGE_TRICOST=0.1\", \
\"bins128_tc0.1_a1e-5\": \"$L/libptb.so+PTB_MERGE_BINS=128+PTB_MERGE_TRICOST=0.1+PTB_MERGE_ALPHA=1e-5+PTB_MERGE_DUP=3\", \
\"bins256_tc1\": \"$L/libptb.so+PTB_MERGE_BINS=256\", \
\"base2\": \"$L/libptb.so\"}" 1 2>&1 | tee $O/variants.log
