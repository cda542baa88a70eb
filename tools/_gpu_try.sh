python tools/dev_rollout_time.py mt 4096 1023 2>&1 | tail -2
python tools/dev_rollout_time.py mt2 4096 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -2
