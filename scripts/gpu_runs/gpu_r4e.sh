# particle deposit rounds hand their losers on as a list: parity tests that tick particles + the stage timing of a game tick
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "particle or long_runs or golden or smoke or split_modes or fullsize" > gpurun_out/r4e_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4e_tests.log; tail -4 gpurun_out/r4e_tests.log
timeout 600 python scripts/e2e_stages.py > gpurun_out/r4e_e2e_stages.json 2> gpurun_out/r4e.err; tail -c 1500 gpurun_out/r4e_e2e_stages.json
