# render planes / scroll / temperature swap validation + baseline bench + role cycles of the current kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2n_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_gputests.log; tail -15 gpurun_out/r2n_gputests.log
python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_mixed.json 2> gpurun_out/r2n_mixed.err
python -c "
import json
d=json.loads(open('gpurun_out/r2n_mixed.json').read().strip().splitlines()[-1]); print('mixed', d['value'], d['ms_per_step'], d['state']['hash'], d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline'].get('phase_ms_by_iteration'))"
python scripts/bench_aux.py > gpurun_out/r2n_aux.json 2> gpurun_out/r2n_aux.err; tail -5 gpurun_out/r2n_aux.err; python -c "
import json
for r in json.load(open('gpurun_out/r2n_aux.json'))['rows']: print(r['kernel'], r['ms'], r['frac_of_peak'])"
FSE_B200_LIB=$PWD/_variants/libfse_role.so python scripts/role_cycles.py 8192 mixed > gpurun_out/r2n_role_mixed.txt 2>&1; cat gpurun_out/r2n_role_mixed.txt
FSE_B200_LIB=$PWD/_variants/libfse_role.so python scripts/role_cycles.py 4096 water > gpurun_out/r2n_role_water.txt 2>&1; cat gpurun_out/r2n_role_water.txt
