mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2r_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_gputests.log; tail -12 gpurun_out/r2r_gputests.log
python scripts/e2e_stages.py > gpurun_out/r2r_e2e_stages.json 2> gpurun_out/r2r_e2e_stages.err; tail -2 gpurun_out/r2r_e2e_stages.err; head -30 gpurun_out/r2r_e2e_stages.json
python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2r_mixed.json 2> gpurun_out/r2r_mixed.err
python -c "
import json
d=json.loads(open('gpurun_out/r2r_mixed.json').read().strip().splitlines()[-1]); print('mixed', d['value'], d['ms_per_step'], d['state']['hash'], d['e2e']['value'], d['e2e']['ms_per_step'])"
