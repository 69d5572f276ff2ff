mkdir -p gpurun_out
timeout 600 python scripts/debug_strip_bodies.py 2 ${1:-6} > gpurun_out/r4l_debug.txt 2>&1; cat gpurun_out/r4l_debug.txt | cut -c1-300 | tail -40
