# validation of physics_check / probe / layers / cooperative spiral + e2e effect of the particle change
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2o_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_gputests.log; tail -15 gpurun_out/r2o_gputests.log
python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_mixed.json 2> gpurun_out/r2o_mixed.err
python -c "
import json
d=json.loads(open('gpurun_out/r2o_mixed.json').read().strip().splitlines()[-1]); print('mixed', d['value'], d['ms_per_step'], d['state']['hash'], d['e2e']['value'], d['e2e']['ms_per_step'])"
python scripts/bench_aux.py > gpurun_out/r2o_aux.json 2> gpurun_out/r2o_aux.err; tail -3 gpurun_out/r2o_aux.err; python -c "
import json
for r in json.load(open('gpurun_out/r2o_aux.json'))['rows']: print(r['kernel'], r['ms'], r['frac_of_peak'])"
