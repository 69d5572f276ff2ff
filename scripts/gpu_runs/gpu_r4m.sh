# the whole game tick on 4 strips (entities, bodies, tools, physicsCheck across three cuts) vs the single-world oracle
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_strips_gpu.py -m gpu -q -x -k "bodies and 4" > gpurun_out/r4m_strips4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4m_strips4.log; tail -5 gpurun_out/r4m_strips4.log | cut -c1-300; grep -n 'fse error' gpurun_out/r4m_strips4.log | head -3 | cut -c1-400
