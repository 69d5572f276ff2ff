mkdir -p gpurun_out
for parts in 1 2 3 4; do
FSE_TICK_PARTS=$parts python bench.py --steps 8 --warmup 4 --no-cpu-baseline --size 32768 --height 4352 > gpurun_out/r2l_slab_parts${parts}.json 2> gpurun_out/r2l.err
python -c "
import json
d=json.loads(open('gpurun_out/r2l_slab_parts${parts}.json').read().strip().splitlines()[-1]); print('parts',$parts, d['value'], d['ms_per_step']); print(d['roofline'].get('phase_ms_by_iteration'))"
done
