# 2 GPUs: strip parity tests (tick + particles + temperature + explosion + eraser across cuts), strong-scaling bench line at N=2
mkdir -p gpurun_out
python -m pytest tests/test_strips_gpu.py -m gpu -x -q > gpurun_out/r2p_strips.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_strips.log; tail -12 gpurun_out/r2p_strips.log
FSE_STRIP_TIMELINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2p_n2.json 2> gpurun_out/r2p_n2.err
tail -3 gpurun_out/r2p_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2p_n2.json').read().strip().splitlines()[-1]); print('n2', d['value'], d['ms_per_step'], d['state']['hash'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['what'][:200]); print(d['roofline'].get('phase_ms_by_iteration')); print({k:(round(v['mean_over_ranks'],3), round(v['max_over_ranks'],3)) for k,v in d['strip_timeline'].items() if k!='what'})"
