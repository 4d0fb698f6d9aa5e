# shared memory vs L1: fewer staged nodes / fewer shared-memory stack entries leave more of the 228 KB to the L1
mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
L=gpu-raytracer_b200/csrc
python tools/gpu_variants_r2.py "{\"base\": \"$L/libptb.so\", \"stage_tlas_64\": \"$L/libptb_st64.so\", \"stage_tlas_32\": \"$L/libptb_st32.so\", \"stage_tlas_8\": \"$L/libptb_st8.so\", \"stage_merged_64\": \"$L/libptb_smt64.so\", \"stage_merged_32\": \"$L/libptb_smt32.so\", \
\"stage32_stack8\": \"$L/libptb_st32k8.so\", \"stage32_stack6\": \"$L/libptb_st32k6.so\", \"stage32_stack12\": \"$L/libptb_st32k12.so\", \"base2\": \"$L/libptb.so\"}" 1 2>&1 | tee $O/variants.log
