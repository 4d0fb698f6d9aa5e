mkdir -p gpurun_out/r2b
python tools/gpu_variants_r2.py '{"base":"gpu-raytracer_b200/csrc/libptb.so","defer3":"gpu-raytracer_b200/csrc/libptb_defer.so","defer4":"gpu-raytracer_b200/csrc/libptb_defer4.so","defer2":"gpu-raytracer_b200/csrc/libptb_defer2.so"}' > gpurun_out/r2b/variants.log 2>&1
python tools/gpu_variants_r2.py '{"base+woop":"gpu-raytracer_b200/csrc/libptb.so+woop"}' 1 >> gpurun_out/r2b/variants.log 2>&1
cat gpurun_out/r2b/variants.log
