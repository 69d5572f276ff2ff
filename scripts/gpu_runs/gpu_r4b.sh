# which material family the mixed world's tick time is spent on (scripts/mixed_ablate.py)
mkdir -p gpurun_out
timeout 800 python scripts/mixed_ablate.py 8192 > gpurun_out/r4b_mixed_ablate.txt 2> gpurun_out/r4b.err; cat gpurun_out/r4b_mixed_ablate.txt; tail -3 gpurun_out/r4b.err
