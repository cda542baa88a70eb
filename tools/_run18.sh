mkdir -p gpurun_out
SPI_B200_MINB=4 python tools/dev_rollout_time.py t5_4 4096 2>&1 | tail -1
SPI_B200_MINB=5 python tools/dev_rollout_time.py t5_5 4096 1023 2>&1 | tail -2
SPI_B200_MINB=3 python tools/dev_rollout_time.py t5_3 4096 2>&1 | tail -1
