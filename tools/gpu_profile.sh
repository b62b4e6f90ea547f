# ncu evidence for profiles/: variants (if built), --set full of the two loop kernels, launch list of a 10-iteration step
mkdir -p gpurun_out/prof
tools/run_variants.sh > gpurun_out/prof/variants.log 2>&1; cat gpurun_out/prof/variants.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass_ -s 8 -c 2 -o gpurun_out/prof/prof_r2_final -f python bench.py --steps 1 --warmup 3 --iters 4 --no-cpu-baseline --no-parity --no-traffic > gpurun_out/prof/ncu.log 2>&1
tail -2 gpurun_out/prof/ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/launches_256_iters10.csv python bench.py --steps 1 --warmup 3 --iters 10 --no-cpu-baseline --no-parity --no-traffic > gpurun_out/prof/launch.log 2>&1
tail -2 gpurun_out/prof/launch.log | cut -c1-300
