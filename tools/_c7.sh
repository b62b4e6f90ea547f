mkdir -p gpurun_out/r2c7
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass_ -s 8 -c 2 -o gpurun_out/r2c7/prof_r2_pa2 -f python bench.py --steps 1 --warmup 3 --iters 4 --no-cpu-baseline --no-parity --no-traffic > gpurun_out/r2c7/ncu.log 2>&1
tail -3 gpurun_out/r2c7/ncu.log
ls -la gpurun_out/r2c7/
