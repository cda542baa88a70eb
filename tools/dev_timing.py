"""dev helper (not a test): raw kernel timing on the GPU."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np, torch
from spi_active_b200.engine import RolloutEngine
from spi_active_b200 import recorders
from spi_active_b200.dataset import concat_windows, pack_segments, to_device, window_recording

eng = RolloutEngine()
fn = recorders.engine_rollout_fn(eng)
wins = [window_recording(recorders.record(n, fn, eng.model), 5) for n in recorders.CONFIG_FILES["all"]]
S, ds = concat_windows(wins)
segs = pack_segments(to_device(ds, eng.device))
print("S", S, "peak fp32 TFLOP/s", eng.fp32_peak(8192))
for C in (20, 256, 1024, 4096):
    params = torch.linspace(0.5, 2.0, C)[:, None] * 6.921
    params = params.cuda()
    for _ in range(2): eng.evaluate_candidates(params, ["mass"], segs)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); n = 3
    for _ in range(n): cost = eng.evaluate_candidates(params, ["mass"], segs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"C={C} ms={ms:.3f} env-steps/s={C*S*5/ms*1e3:.3e} substeps/s={C*S*40/ms*1e3:.3e} TFLOP/s(7.75k/substep)={C*S*40*7748/ms*1e3/1e12:.2f}")
w = cost.cpu().numpy() @ np.array([10., 5., 1.])
print("argmin mass", float(params[int(np.argmin(w)), 0]))
