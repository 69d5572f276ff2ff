mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2y_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_gputests.log; tail -12 gpurun_out/r2y_gputests.log
for sp in 1 0; do
FSE_P2_SPLIT=$sp python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2y_mixed_split$sp.json 2> gpurun_out/r2y_mixed_split$sp.err
python -c "
import json
d=json.loads(open('gpurun_out/r2y_mixed_split$sp.json').read().strip().splitlines()[-1]); print('split $sp', d['value'], d['ms_per_step'], d['state']['hash'], d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline'].get('phase_ms_by_iteration'))"
done
