# after keeping the mbarrier phases across segments and double-buffering the CCL sweeps: tests (with a timeout: a phase mismatch would hang), racecheck again
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r3c_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3c_gputests.log; tail -6 gpurun_out/r3c_gputests.log
for m in per_pass split aux; do
  timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_small.py $m > gpurun_out/r3c_san_racecheck_${m}.log 2>&1; echo "rc=$?" >> gpurun_out/r3c_san_racecheck_${m}.log
  grep "Race reported\|RACECHECK SUMMARY\|sanitize_small" gpurun_out/r3c_san_racecheck_${m}.log | head -8
done
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r3c_mixed.json 2> gpurun_out/r3c_mixed.err
python -c "
import json
d=json.loads(open('gpurun_out/r3c_mixed.json').read().strip().splitlines()[-1]); print('mixed', d['value'], d['ms_per_step'], d['state']['hash'], d['e2e']['value'], d['e2e']['ms_per_step'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload sparse --size 16384 --height 8192 --active 0 > gpurun_out/r3c_sparse_noactive.json 2> gpurun_out/r3c_sparse.err
python -c "
import json
d=json.loads(open('gpurun_out/r3c_sparse_noactive.json').read().strip().splitlines()[-1]); print('sparse noactive', d['value'], d['ms_per_step'])"
