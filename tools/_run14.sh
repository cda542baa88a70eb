mkdir -p gpurun_out
python tools/dev_rollout_time.py fold 4096 1023 2>&1 | tail -3
python tools/ws_timeline.py run 1024 2>&1 | tail -28
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_readme_bowl.py -m gpu -x -q 2>&1 | tail -5
