mkdir -p gpurun_out
N=$1
FSE_STRIP_TIMELINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --steps 12 --warmup 3 > gpurun_out/r2s_n$N.json 2> gpurun_out/r2s_n$N.err
tail -3 gpurun_out/r2s_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/r2s_n$N.json').read().strip().splitlines()[-1]); print('n$N', d['value'], d['ms_per_step'], d['state']['hash'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks']); print(d['roofline'].get('phase_ms_by_iteration')); print({k:(round(v['mean_over_ranks'],3), round(v['max_over_ranks'],3)) for k,v in d['strip_timeline'].items() if k!='what'})"
if [ "$N" = "4" ]; then python -m pytest tests/test_strips_gpu.py -m gpu -x -q > gpurun_out/r2s_strips4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_strips4.log; tail -4 gpurun_out/r2s_strips4.log; fi
