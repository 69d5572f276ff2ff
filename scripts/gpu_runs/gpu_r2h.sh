mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_gputests.log; tail -25 gpurun_out/r2h_gputests.log | cut -c1-250
for skip in 2 0; do
FSE_ROW_SKIP=$skip python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_mixed_skip$skip.json 2> gpurun_out/r2h_mixed_skip$skip.err
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload bodies > gpurun_out/r2h_bodies.json 2> gpurun_out/r2h_bodies.err
for f in gpurun_out/r2h_mixed_skip2.json gpurun_out/r2h_mixed_skip0.json gpurun_out/r2h_bodies.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['state']['hash'], d.get('bodies')); print(d['roofline'].get('phase_ms_by_iteration'))
except Exception as e: print('ERR', e)
"; done
