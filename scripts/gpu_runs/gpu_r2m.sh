mkdir -p gpurun_out
N=$1
for bmin in 64 100000; do
FSE_STRIP_BOUNDARY_MIN=$bmin FSE_STRIP_TIMELINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2m_n${N}_bmin$bmin.json 2> gpurun_out/r2m_n${N}_bmin$bmin.err
python -c "
import json
d=json.loads(open('gpurun_out/r2m_n${N}_bmin$bmin.json').read().strip().splitlines()[-1]); print('bmin',$bmin, d['value'], d['ms_per_step'], d['state']['hash'], d['e2e']['value']); print(d['roofline'].get('phase_ms_by_iteration')); print({k:(round(v['mean_over_ranks'],3), round(v['max_over_ranks'],3)) for k,v in d['strip_timeline'].items() if k!='what'})"
tail -2 gpurun_out/r2m_n${N}_bmin$bmin.err
done
