mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_gputests.log; tail -25 gpurun_out/r2g_gputests.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -3 gpurun_out/r2g_smoke.log
python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_mixed.json 2> gpurun_out/r2g_mixed.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload air > gpurun_out/r2g_air.json 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload sparse --size 16384 --height 8192 --active 0 > gpurun_out/r2g_sparse_noactive.json 2>&1
for f in gpurun_out/r2g_mixed.json gpurun_out/r2g_air.json gpurun_out/r2g_sparse_noactive.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['state']['hash']); print(d['roofline'].get('phase_ms_by_iteration'))
except Exception as e: print('ERR', e)
"; done
