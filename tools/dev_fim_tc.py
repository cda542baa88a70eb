"""dev helper: structured inputs through spi_b200_fim_contract, printed, to diagnose layout mistakes in one GPU run."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from spi_active_b200.engine import RolloutEngine
np.set_printoptions(linewidth=220, precision=4, suppress=True)
eng = RolloutEngine()
M, P, delta = 9, 10, 1.0
for T in (1, 2, 3, 5):
    hist = np.zeros((T, M, P + 1, 25), np.float32)
    for t in range(T):
        for m in range(M):
            for p in range(P):
                hist[t, m, p + 1, (p + t) % 25] = -(p + 1 + 0.001 * m)      # J[p][(p+t)%25] = p + 1 + m/1000
                hist[t, m, p + 1, 24] += -0.5                               # common column -> off-diagonals 0.25 per step
    J = hist[:, :, 0:1, :].astype(np.float64) - hist[:, :, 1:, :]
    ref = np.einsum("tmpd,tmqd->mpq", J, J)
    jtj, tr = eng.fim_contract(torch.from_numpy(hist), delta)
    torch.cuda.synchronize()
    jtj = jtj.cpu().numpy()
    print(f"T={T} max abs err {np.abs(jtj - ref).max():.3e}  trace err {np.abs(tr.cpu().numpy() - np.trace(ref, axis1=1, axis2=2)).max():.3e}")
    if np.abs(jtj - ref).max() > 1e-3:
        for m in (0, 1, 8):
            print("env", m, "got\n", jtj[m], "\nref\n", ref[m])
        break
import time
T, M, P = 1248, 1024, 10
h = torch.randn(T, M, P + 1, 25, device="cuda")
live = torch.ones(T, M, dtype=torch.uint8, device="cuda")
for _ in range(2): eng.fim_contract(h, 0.1, live=live)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.fim_contract(h, 0.1, live=live); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"T={T} M={M} P={P}: {ms:.3f} ms; hist {h.numel()*4/1e9:.2f} GB -> {h.numel()*4/ms/1e6:.0f} GB/s; useful {2*M*P*P*25*T/ms/1e9:.2f} TFLOP/s, issued {3*2*128*128*32*T*(M/8)/ms/1e9:.1f} TFLOP/s tf32")
# the chunk the rollout actually contracts: 64 steps
for Tc in (64, 256):
    hc, lc = h[:Tc].contiguous(), live[:Tc].contiguous()
    for _ in range(3): eng.fim_contract(hc, 0.1, live=lc)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): eng.fim_contract(hc, 0.1, live=lc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"T={Tc} M={M} P={P}: {ms*1e3:.1f} us; hist {hc.numel()*4/1e6:.1f} MB -> {hc.numel()*4/ms/1e6:.0f} GB/s")
