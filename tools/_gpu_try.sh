mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_e.json 2> gpurun_out/bench_r2_e.err; echo bench rc=$?
