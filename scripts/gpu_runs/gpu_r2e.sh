mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_gputests.log; tail -4 gpurun_out/r2e_gputests.log
for skip in 1 0; do
  FSE_ROW_SKIP=$skip python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_mixed_skip$skip.json 2> gpurun_out/r2e_mixed_skip$skip.err
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload column --size 2048 > gpurun_out/r2e_column.json 2>&1
FSE_FUSED_MAX_CHUNKS=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload column --size 2048 > gpurun_out/r2e_column_perpass.json 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload sparse --size 16384 --height 8192 --active 0 > gpurun_out/r2e_sparse_noactive.json 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload air > gpurun_out/r2e_air.json 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_ncu_bench.log 2>&1
for f in gpurun_out/r2e_mixed_skip1.json gpurun_out/r2e_mixed_skip0.json gpurun_out/r2e_column.json gpurun_out/r2e_column_perpass.json gpurun_out/r2e_sparse_noactive.json gpurun_out/r2e_air.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'], d['state']['hash'], d.get('bodies')); print(d['roofline'].get('phase_ms_by_iteration'))
except Exception as e: print('ERR', e)
"; done
