mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_c_n$N.json 2> gpurun_out/bench_r2_c_n$N.err; echo rc=$?
cut -c1-300 gpurun_out/bench_r2_c_n$N.json; tail -3 gpurun_out/bench_r2_c_n$N.err
