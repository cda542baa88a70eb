mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; echo rc=$?
cat gpurun_out/bench_r2_a.json; tail -5 gpurun_out/bench_r2_a.err
