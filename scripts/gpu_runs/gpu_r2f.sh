mkdir -p gpurun_out
python -m pytest tests/test_strips_gpu.py -m gpu -x -q > gpurun_out/r2f_strips.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_strips.log; tail -30 gpurun_out/r2f_strips.log | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2f_scale_n2.json 2> gpurun_out/r2f_scale_n2.err
tail -c 1500 gpurun_out/r2f_scale_n2.json; tail -5 gpurun_out/r2f_scale_n2.err
