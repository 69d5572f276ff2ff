# longest-first order with the lightest chunks next to the heaviest on every SM (FSE_LPT_MIX), single launch per pass
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r4f_$name.json 2> gpurun_out/r4f.err
python -c "
import json
d=json.loads(open('gpurun_out/r4f_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],3), round(d['ms_per_step'],3), d['state']['hash'], round(d['e2e']['ms_per_step'],2)); print(d['roofline'].get('phase_ms_by_iteration'))"; }
run default FSE_X=0
run parts1 FSE_TICK_PARTS=1
run parts1_mix148 FSE_TICK_PARTS=1 FSE_LPT_MIX=148
run parts1_mix296 FSE_TICK_PARTS=1 FSE_LPT_MIX=296
run parts1_mix74 FSE_TICK_PARTS=1 FSE_LPT_MIX=74
