# single-GPU paths of the calls whose kernels the strip work touched last (tools incl. the vacuum, entities, bridge, physicsCheck) + memcheck of the helper calls
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -k "tool or entit or bridge or physics or vacuum or raster" > gpurun_out/r4q_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4q_tests.log; tail -3 gpurun_out/r4q_tests.log | cut -c1-300
timeout 100 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py aux > gpurun_out/r4q_san_memcheck_aux.log 2>&1; echo "rc=$?" >> gpurun_out/r4q_san_memcheck_aux.log
grep "ERROR SUMMARY\|sanitize_small\|^rc=" gpurun_out/r4q_san_memcheck_aux.log | head -4
