# the whole game tick on 2 strips with the vacuum added to the tools
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_strips_gpu.py -m gpu -q -x -k "bodies and 2" > gpurun_out/r4p_strips2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4p_strips2.log; tail -5 gpurun_out/r4p_strips2.log | cut -c1-300; grep -n 'fse error\|AssertionError' gpurun_out/r4p_strips2.log | head -3 | cut -c1-400
