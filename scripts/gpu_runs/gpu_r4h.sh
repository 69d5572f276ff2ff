# strips on the final code: tick + particles + temperature + explosion + eraser + scroll across the cuts (2 ranks; 4 when the box has them)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_strips_gpu.py -m gpu -q > gpurun_out/r4h_strips.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4h_strips.log; tail -6 gpurun_out/r4h_strips.log
