# full-size (8192^2) property tests of the CUDA path
mkdir -p gpurun_out
( time timeout 800 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x --durations=5 ) > gpurun_out/r4d_fullsize.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4d_fullsize.log; tail -25 gpurun_out/r4d_fullsize.log
