mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r2_a.csv python tools/profile_target.py 4096 2 > gpurun_out/pt_r2.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:rollout_ws_kernelILb0 -s 2 -c 1 -o gpurun_out/prof_ws_r2a python tools/profile_target.py 1024 3 > gpurun_out/pw_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
