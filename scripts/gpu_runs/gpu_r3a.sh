mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r3a_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3a_gputests.log; tail -6 gpurun_out/r3a_gputests.log
python scripts/e2e_stages.py > gpurun_out/r3a_e2e_stages.json 2> gpurun_out/r3a_e2e_stages.err; tail -2 gpurun_out/r3a_e2e_stages.err; head -26 gpurun_out/r3a_e2e_stages.json
python scripts/bench_aux.py > gpurun_out/r3a_aux.json 2> gpurun_out/r3a_aux.err; python -c "
import json
for r in json.load(open('gpurun_out/r3a_aux.json'))['rows']: print(r['kernel'], r['ms'], r['frac_of_peak'])" | tail -3
