mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2u_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_gputests.log; tail -25 gpurun_out/r2u_gputests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload bodies > gpurun_out/r2u_bodies.json 2> gpurun_out/r2u_bodies.err; tail -2 gpurun_out/r2u_bodies.err
python -c "
import json
d=json.loads(open('gpurun_out/r2u_bodies.json').read().strip().splitlines()[-1]); print('bodies', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['bodies'])"
python scripts/bench_aux.py > gpurun_out/r2u_aux.json 2> gpurun_out/r2u_aux.err; tail -3 gpurun_out/r2u_aux.err; python -c "
import json
for r in json.load(open('gpurun_out/r2u_aux.json'))['rows']: print(r['kernel'], r['ms'], r['frac_of_peak'])"
