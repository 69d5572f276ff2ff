# row-skip validation (after the pass-3 fix) + A/B timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_gputests.log; tail -15 gpurun_out/r2c_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; tail -3 gpurun_out/r2c_smoke.log
for skip in 1 0; do
  FSE_ROW_SKIP=$skip python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_mixed_skip$skip.json 2> gpurun_out/r2c_mixed_skip$skip.err
  FSE_ROW_SKIP=$skip python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload air > gpurun_out/r2c_air_skip$skip.json 2> gpurun_out/r2c_air_skip$skip.err
done
FSE_ROW_SKIP=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload column --size 2048 > gpurun_out/r2c_column.json 2>&1
FSE_ROW_SKIP=1 FSE_FUSED_MAX_CHUNKS=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload column --size 2048 > gpurun_out/r2c_column_perpass.json 2>&1
for f in gpurun_out/r2c_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
for skip in 1 0; do
  FSE_ROW_SKIP=$skip python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload sparse --size 16384 --height 8192 --active 0 > gpurun_out/r2c_sparse_noactive_skip$skip.json 2>&1
  FSE_ROW_SKIP=$skip FSE_FUSED_MAX_CHUNKS=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload column --size 2048 > gpurun_out/r2c_column_perpass_skip$skip.json 2>&1
done
for f in gpurun_out/r2c_sparse*.json gpurun_out/r2c_column_perpass_skip*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
