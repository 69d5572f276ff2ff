mkdir -p gpurun_out
for sp in 0 1 0 1; do
FSE_P2_SPLIT=$sp python bench.py --no-cpu-baseline > gpurun_out/r3e_default_split$sp.json 2> gpurun_out/r3e.err
python -c "
import json
d=json.loads(open('gpurun_out/r3e_default_split$sp.json').read().strip().splitlines()[-1]); print('split $sp', round(d['value'],3), round(d['ms_per_step'],3), d['state']['hash'], round(d['e2e']['ms_per_step'],2)); print(d['roofline'].get('phase_ms_by_iteration'))"
done
