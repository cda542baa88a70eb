#!/bin/bash
# the driver's multi-GPU launch of bench.py, for a record under gpurun_out/:   gpurun --gpus N -- 'bash tools/bench_multi_gpu.sh N TAG'
mkdir -p gpurun_out
N=$1; TAG=${2:-run}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo rc=$?
cut -c1-300 gpurun_out/bench_${TAG}_n$N.json; tail -2 gpurun_out/bench_${TAG}_n$N.err
