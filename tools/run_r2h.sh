mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
timeout 1200 python -m pytest tests -m gpu -q > $O/all_gpu_tests.log 2>&1; tail -15 $O/all_gpu_tests.log
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_config1.csv python bench.py --config 1 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; tail -2 $O/ncu_bench.log
# full capture, every kernel of one 9-pass wave frame + one SVGF + TAA frame; exported to CSV on the box (the report itself is > 64 MiB)
PTB_PROF_WAVE=9 timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/prof_all python tools/prof_all.py > $O/ncu_full.log 2>&1; tail -3 $O/ncu_full.log
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > $O/prof_all_raw.csv 2> $O/export.err
ls -la /tmp/prof_all.ncu-rep
# source-level capture of two closest-hit trace launches (bounce 0 and 1) only
PTB_PROF_WAVE=9 PTB_PROF_SVGF=0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_trace8 -c 2 -o $O/prof_trace8_src python tools/prof_all.py > $O/ncu_src.log 2>&1; tail -3 $O/ncu_src.log
ncu -i $O/prof_trace8_src.ncu-rep --page source --csv > $O/prof_trace8_source.csv 2>> $O/export.err
gzip -9 $O/prof_all_raw.csv $O/prof_trace8_source.csv
du -sh $O; ls -la $O
