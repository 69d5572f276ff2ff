mkdir -p gpurun_out
N=$1
if [ "$N" = "1" ]; then
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --size 32768 --height 33024 > gpurun_out/r2j_cfg3_n1.json 2> gpurun_out/r2j_cfg3_n1.err
  tail -c 1200 gpurun_out/r2j_cfg3_n1.json; tail -3 gpurun_out/r2j_cfg3_n1.err
else
  FSE_STRIP_TIMELINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2j_cfg3_n$N.json 2> gpurun_out/r2j_cfg3_n$N.err
  tail -c 2500 gpurun_out/r2j_cfg3_n$N.json; tail -5 gpurun_out/r2j_cfg3_n$N.err
  if [ "$N" = "4" ]; then python -m pytest tests/test_strips_gpu.py -m gpu -x -q 2>&1 | tail -3; fi
fi
