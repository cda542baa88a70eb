mkdir -p gpurun_out
B=tools/_build
SPI_B200_LIB=$B/libspi_b200_m5.so SPI_B200_MINB=5 python tools/dev_rollout_time.py m5_5 4096 2>&1 | tail -1
SPI_B200_LIB=$B/libspi_b200_cold.so SPI_B200_MINB=4 python tools/dev_rollout_time.py cold_4 4096 2>&1 | tail -1
SPI_B200_LIB=$B/libspi_b200_cold.so SPI_B200_MINB=5 python tools/dev_rollout_time.py cold_5 4096 2>&1 | tail -1
SPI_B200_LIB=$B/libspi_b200_park.so SPI_B200_MINB=4 python tools/dev_rollout_time.py park_4 4096 2>&1 | tail -1
SPI_B200_LIB=$B/libspi_b200_park.so SPI_B200_MINB=5 python tools/dev_rollout_time.py park_5 4096 1023 2>&1 | tail -2
python - <<'PY'
import numpy as np
a=np.load("gpurun_out/cost_fold_4096.npy")
for t in ("m5_5","cold_4","cold_5","park_4","park_5"):
    b=np.load(f"gpurun_out/cost_{t}_4096.npy"); print(t, "max rel diff vs fold", np.abs(a-b).max()/np.abs(a).max())
PY
