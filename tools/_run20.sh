mkdir -p gpurun_out
python tests/tools/dev_accuracy.py 2>&1 | tail -1
SPI_B200_LIB=tools/_build/libspi_b200_dsc.so python tests/tools/dev_accuracy.py 2>&1 | tail -1
python tools/dev_rollout_time.py base 4096 2>&1 | tail -1
SPI_B200_LIB=tools/_build/libspi_b200_dsc.so python tools/dev_rollout_time.py dsc 4096 2>&1 | tail -1
python tools/dev_rollout_time.py base2 4096 2>&1 | tail -1
SPI_B200_LIB=tools/_build/libspi_b200_dsc.so python tools/dev_rollout_time.py dsc2 4096 2>&1 | tail -1
