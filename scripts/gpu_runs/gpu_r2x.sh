mkdir -p gpurun_out
FSE_ROW_SKIP=0 FSE_TICK_PARTS=1 ncu --set full --clock-control none --import-source on -k regex:tick_pass_kernel --launch-skip 60 --launch-count 2 -o gpurun_out/r2x_pass_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2x_ncu_full.log 2>&1; tail -2 gpurun_out/r2x_ncu_full.log
ncu -i gpurun_out/r2x_pass_full.ncu-rep --page raw --csv > gpurun_out/r2x_raw.csv 2>/dev/null
python scripts/ncu_summary.py full gpurun_out/r2x_raw.csv gpurun_out/r2x_full.md
