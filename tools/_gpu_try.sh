python tools/dev_rollout_time.py ub 4096 1023 2>&1 | tail -2
python tests/tools/dev_accuracy.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_readme_bowl.py -m gpu -x -q 2>&1 | tail -2
