#!/bin/bash
# multi-GPU bench line (launched the way the driver launches it)
N=$1
mkdir -p gpurun_out
set -x
nvidia-smi topo -m > gpurun_out/r02n_topo_n$N.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02n_bench_n$N.json 2> gpurun_out/r02n_bench_n$N.err; tail -3 gpurun_out/r02n_bench_n$N.err; head -c 400 gpurun_out/r02n_bench_n$N.json
