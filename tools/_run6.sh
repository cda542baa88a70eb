mkdir -p gpurun_out
SPI_B200_LIB=$PWD/spi_active_b200/libspi_b200_direct.so python tools/dev_halves.py direct 4096 > gpurun_out/calf_direct.log 2>&1
python tools/dev_halves.py harm 4096 > gpurun_out/calf_harm.log 2>&1
cat gpurun_out/calf_direct.log gpurun_out/calf_harm.log
python - <<'PY'
import numpy as np
a=np.load("gpurun_out/cost_direct_4096.npy"); b=np.load("gpurun_out/cost_harm_4096.npy")
print("max rel diff", np.abs(a-b).max()/np.abs(a).max(), "max abs", np.abs(a-b).max())
PY
