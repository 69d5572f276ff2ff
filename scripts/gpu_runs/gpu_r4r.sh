# final state of the round on 2 GPUs: both strip tests at 2 ranks
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_strips_gpu.py -m gpu -q -k "2" > gpurun_out/r4r_strips2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4r_strips2.log; tail -4 gpurun_out/r4r_strips2.log | cut -c1-300
