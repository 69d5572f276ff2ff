# memcheck + racecheck of the helper calls (tools, entities, bodies, physicsCheck, scroll ...) after the strip work touched their kernels
mkdir -p gpurun_out
for tool in memcheck racecheck; do
timeout 150 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py aux > gpurun_out/r4o_san_${tool}_aux.log 2>&1; echo "rc=$?" >> gpurun_out/r4o_san_${tool}_aux.log
grep "ERROR SUMMARY\|RACECHECK SUMMARY\|sanitize_small\|rc=\|Error\|error" gpurun_out/r4o_san_${tool}_aux.log | head -6
done
