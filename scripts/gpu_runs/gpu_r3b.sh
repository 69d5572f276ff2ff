# compute-sanitizer over every kernel family of the final code (tiny worlds; the tools slow kernels down 10-100x)
mkdir -p gpurun_out
for m in per_pass split fused aux; do
  for tool in memcheck racecheck; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py $m > gpurun_out/r3b_san_${tool}_${m}.log 2>&1; echo "rc=$?" >> gpurun_out/r3b_san_${tool}_${m}.log
    tail -4 gpurun_out/r3b_san_${tool}_${m}.log
  done
done
