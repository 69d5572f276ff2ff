mkdir -p gpurun_out
for deal in 0 1; do for parts in 3 4; do
FSE_LPT_DEAL=$deal FSE_TICK_PARTS=$parts python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_mixed_deal${deal}_parts${parts}.json 2> gpurun_out/r2k_mixed.err
python -c "
import json
d=json.loads(open('gpurun_out/r2k_mixed_deal${deal}_parts${parts}.json').read().strip().splitlines()[-1]); print('deal',$deal,'parts',$parts, d['value'], d['ms_per_step'], d['state']['hash']); print(d['roofline'].get('phase_ms_by_iteration'))"
done; done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "parts or benched" 2>&1 | tail -3
FSE_LPT_DEAL=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "parts or benched" 2>&1 | tail -3
