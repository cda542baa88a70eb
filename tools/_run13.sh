mkdir -p gpurun_out
python tools/dev_halves.py base 4096 1023 > gpurun_out/d_base.log 2>&1
SPI_B200_WS_DUAL=1 timeout 300 python tools/dev_halves.py dual 4096 1023 > gpurun_out/d_dual.log 2>&1
SPI_B200_WS_DUAL=1 SPI_B200_MINB=2 timeout 300 python tools/dev_halves.py dual2 4096 > gpurun_out/d_dual2.log 2>&1
cat gpurun_out/d_base.log gpurun_out/d_dual.log gpurun_out/d_dual2.log | tail -8
python - <<'PY'
import numpy as np
for C in (4096, 1023):
    a=np.load(f"gpurun_out/cost_base_{C}.npy"); b=np.load(f"gpurun_out/cost_dual_{C}.npy")
    print(C, "bit-identical:", np.array_equal(a,b), "maxdiff", np.abs(a-b).max())
PY
