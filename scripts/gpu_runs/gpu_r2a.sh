set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_gputests.log; tail -3 gpurun_out/r2a_gputests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.json
for m in per_pass fused aux; do
  for tool in memcheck racecheck; do
    timeout 500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py $m > gpurun_out/r2a_san_${tool}_${m}.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_san_${tool}_${m}.log
    tail -4 gpurun_out/r2a_san_${tool}_${m}.log
  done
done
