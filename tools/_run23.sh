mkdir -p gpurun_out
python tools/dev_rollout_time.py base 4096 2>&1 | tail -1
SPI_B200_LIB=tools/_build/libspi_b200_harm2.so python tools/dev_rollout_time.py harm 4096 2>&1 | tail -1
python tools/dev_rollout_time.py base2 4096 2>&1 | tail -1
SPI_B200_LIB=tools/_build/libspi_b200_harm2.so python tools/dev_rollout_time.py harm2 4096 1023 2>&1 | tail -2
SPI_B200_LIB=tools/_build/libspi_b200_harm2.so python tests/tools/dev_accuracy.py 2>&1 | tail -1
python - <<'PY'
import numpy as np
a=np.load("gpurun_out/cost_base_4096.npy"); b=np.load("gpurun_out/cost_harm_4096.npy")
print("harm vs base: max rel diff", (np.abs(a-b)/np.abs(a)).max())
PY
SPI_B200_LIB=tools/_build/libspi_b200_harm2.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_readme_bowl.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
