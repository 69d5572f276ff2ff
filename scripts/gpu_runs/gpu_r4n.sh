# final code of the round on one GPU: GPU tests, smoke, default bench line (no CPU baseline: the reference arm was taken in gpu_r4i.sh)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/r4n_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4n_gputests.log; tail -7 gpurun_out/r4n_gputests.log | cut -c1-300
timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r4n_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r4n_smoke.log; tail -3 gpurun_out/r4n_smoke.log | cut -c1-250
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r4n_bench.json 2> gpurun_out/r4n_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r4n_bench.json').read().strip().splitlines()[-1]); print('default', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'], d['gpu_launches'], d['state']['hash'])"
