mkdir -p gpurun_out
python tools/dev_rollout_time.py t3 4096 1023 2>&1 | tail -2
python tools/ws_timeline.py run 1024 2>&1 | tail -24
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
