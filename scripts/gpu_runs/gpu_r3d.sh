# final code: GPU test suite, memcheck of every kernel family, default bench line + reference arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r3d_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3d_gputests.log; tail -4 gpurun_out/r3d_gputests.log
for m in per_pass split fused aux; do
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py $m > gpurun_out/r3d_san_memcheck_${m}.log 2>&1; echo "rc=$?" >> gpurun_out/r3d_san_memcheck_${m}.log
  grep "ERROR SUMMARY\|sanitize_small" gpurun_out/r3d_san_memcheck_${m}.log | head -3
done
timeout 600 python bench.py > gpurun_out/r3d_bench_default.json 2> gpurun_out/r3d_bench_default.err
python -c "
import json
d=json.loads(open('gpurun_out/r3d_bench_default.json').read().strip().splitlines()[-1]); print('default', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'])"
timeout 600 python bench.py --workload bodies --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3d_bodies.json 2> gpurun_out/r3d_bodies.err
python -c "
import json
d=json.loads(open('gpurun_out/r3d_bodies.json').read().strip().splitlines()[-1]); print('bodies', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['bodies'])"
