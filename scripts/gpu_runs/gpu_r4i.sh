# final code of the round: what the driver runs (GPU tests, smoke, default bench line, reference arm)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r4i_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4i_gputests.log; tail -8 gpurun_out/r4i_gputests.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r4i_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r4i_smoke.log; tail -3 gpurun_out/r4i_smoke.log
timeout 600 python bench.py > gpurun_out/r4i_bench_default.json 2> gpurun_out/r4i_bench_default.err
python -c "
import json
d=json.loads(open('gpurun_out/r4i_bench_default.json').read().strip().splitlines()[-1]); print('default', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'], d['state']['hash'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r4i_bench_reference.json 2> gpurun_out/r4i_bench_reference.err; tail -c 600 gpurun_out/r4i_bench_reference.json
