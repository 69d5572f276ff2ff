# rigid-body bridge on multi-rank strips vs the single-world oracle (2 ranks; 4 when the box has them) + the existing strip test
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_strips_gpu.py -m gpu -q -x > gpurun_out/r4k_strips.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4k_strips.log; tail -40 gpurun_out/r4k_strips.log | cut -c1-400
