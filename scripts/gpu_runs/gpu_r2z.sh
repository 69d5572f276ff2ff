# what the driver runs at round end: smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -3 gpurun_out/r2z_smoke.log
( time python bench.py ) > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err; tail -4 gpurun_out/r2z_bench_default.err
python -c "
import json
d=json.loads(open('gpurun_out/r2z_bench_default.json').read().strip().splitlines()[-1]); print('default', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])"
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; tail -4 gpurun_out/r2z_bench_reference.err; cat gpurun_out/r2z_bench_reference.json | cut -c1-900
