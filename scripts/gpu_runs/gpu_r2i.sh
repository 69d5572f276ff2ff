mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2i_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_gputests.log; tail -25 gpurun_out/r2i_gputests.log | cut -c1-250
for skip in 2 0 1; do
FSE_ROW_SKIP=$skip python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_mixed_skip$skip.json 2> gpurun_out/r2i_mixed_skip$skip.err
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload column --size 2048 > gpurun_out/r2i_column.json 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload sparse --size 65536 --height 16384 > gpurun_out/r2i_sparse_cfg5.json 2>&1
for f in gpurun_out/r2i_mixed_skip2.json gpurun_out/r2i_mixed_skip0.json gpurun_out/r2i_mixed_skip1.json gpurun_out/r2i_column.json gpurun_out/r2i_sparse_cfg5.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['state']['hash'], d['config'].get('awake_chunks')); print(d['roofline'].get('phase_ms_by_iteration'))
except Exception as e: print('ERR', e)
"; done
