# re-tune after the round's changes: byte -> float route of the node test (XU vs ALU+FMA pipes; the split-BVH tree shifted the balance), shade / sort occupancy
mkdir -p gpurun_out/r2ab
L=gpu-raytracer_b200/csrc
python tools/gpu_variants_r2.py "{\"base\": \"$L/libptb.so\", \"cvt_all_i2f\": \"$L/libptb_cvt00.so\", \"cvt_xmin_magic\": \"$L/libptb_cvt01.so\", \"cvt_4_magic\": \"$L/libptb_cvt1b.so\", \"cvt_all_magic\": \"$L/libptb_cvt3f.so\", \
\"shade_3_ctas\": \"$L/libptb_sh3.so\", \"shade_5_ctas\": \"$L/libptb_sh5.so\", \"sort_5_ctas\": \"$L/libptb_so5.so\", \"base2\": \"$L/libptb.so\"}" 1 2>&1 | tee gpurun_out/r2ab/variants.log
