set -x
mkdir -p gpurun_out
SPI_B200_WS_HALVES=1 python tools/dev_halves.py h1 4096 1023 > gpurun_out/h1.log 2>&1
SPI_B200_WS_HALVES=2 timeout 300 python tools/dev_halves.py h2 4096 1023 > gpurun_out/h2.log 2>&1
python - <<'PY' > gpurun_out/hcmp.log 2>&1
import numpy as np
for C in (4096, 1023):
    a=np.load(f"gpurun_out/cost_h1_{C}.npy"); b=np.load(f"gpurun_out/cost_h2_{C}.npy")
    print(C, "bit-identical:", np.array_equal(a,b), "maxdiff", np.abs(a-b).max())
PY
cat gpurun_out/h1.log gpurun_out/h2.log gpurun_out/hcmp.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests.log; cat gpurun_out/gpu_tests.log
