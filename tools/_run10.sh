python tools/dev_fim_tc.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_fim_tc.py tests/test_active.py -m gpu -x -q 2>&1 | tail -5
for sp in 1 2 3 8; do echo split=$sp; SPI_B200_FIM_SPLIT=$sp python tools/dev_fim_tc.py 2>&1 | grep -E "^T=(64|256|1248) M"; done
