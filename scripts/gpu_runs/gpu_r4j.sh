# final code: ncu launch list + DRAM bytes of the bench command, one --set full capture of a whole-phase pass-1 / pass-2 launch (plain instantiation)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r4j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r4j_ncu_bench.log 2>&1; tail -2 gpurun_out/r4j_ncu_bench.log
FSE_ROW_SKIP=0 FSE_P2_SPLIT=0 FSE_TICK_PARTS=1 ncu --set full --clock-control none --import-source on -k regex:tick_pass_kernel --launch-skip 60 --launch-count 2 -o gpurun_out/r4j_pass_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r4j_ncu_full.log 2>&1; tail -2 gpurun_out/r4j_ncu_full.log
ncu -i gpurun_out/r4j_pass_full.ncu-rep --page raw --csv > gpurun_out/r4j_raw.csv 2>/dev/null
python scripts/ncu_summary.py launches gpurun_out/r4j_launches.csv gpurun_out/r4j_launches.md gpurun_out/r4j_traffic.json
python scripts/ncu_summary.py full gpurun_out/r4j_raw.csv gpurun_out/r4j_full.md
cat gpurun_out/r4j_launches.md | head -30
