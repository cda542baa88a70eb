mkdir -p gpurun_out
python tools/dev_fim_tc.py 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; echo rc=$?; cat gpurun_out/bench_r2_b.json; tail -3 gpurun_out/bench_r2_b.err
