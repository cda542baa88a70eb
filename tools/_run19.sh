mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_d.json 2> gpurun_out/bench_r2_d.err; echo bench rc=$?
cut -c1-900 gpurun_out/bench_r2_d.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r2_d.csv python tools/profile_target.py 4096 2 > gpurun_out/pt_r2d.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:rollout_ws_kernelILb0 -s 2 -c 1 -o gpurun_out/prof_ws_r2d python tools/profile_target.py 1024 3 > gpurun_out/pw_r2d.log 2>&1
ls -la gpurun_out/*r2d* | tail -3
