# outline rewrite validation + bodies workload + e2e stage breakdown
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bridge.py tests/test_gpu_render.py -m gpu -x -q > gpurun_out/r2q_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_gputests.log; tail -12 gpurun_out/r2q_gputests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload bodies > gpurun_out/r2q_bodies.json 2> gpurun_out/r2q_bodies.err; tail -2 gpurun_out/r2q_bodies.err
python -c "
import json
d=json.loads(open('gpurun_out/r2q_bodies.json').read().strip().splitlines()[-1]); print('bodies', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['bodies'])"
python scripts/e2e_stages.py > gpurun_out/r2q_e2e_stages.json 2> gpurun_out/r2q_e2e_stages.err; tail -2 gpurun_out/r2q_e2e_stages.err; cat gpurun_out/r2q_e2e_stages.json
