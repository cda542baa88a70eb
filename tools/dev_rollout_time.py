"""dev helper (not a test): time the rollout kernel at the bench shape and dump the costs, so that two builds / launch
shapes (SPI_B200_MINB=2..5, SPI_B200_LIB=...) can be compared bit for bit:  python tools/dev_rollout_time.py TAG [C ...]"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from spi_active_b200 import cem, recorders
from spi_active_b200.dataset import pack_segments, to_device
from spi_active_b200.engine import RolloutEngine

tag = sys.argv[1]
eng = RolloutEngine()
S, ds = bench.build_dataset(recorders.engine_rollout_fn(eng), eng.model)
segs = pack_segments(to_device(ds, eng.device))
cfg = cem.default_full_config(eng.model)
peak = eng.fp32_peak(8192)[0]
out = Path("gpurun_out"); out.mkdir(exist_ok=True)
for C in [int(x) for x in (sys.argv[2:] or ["4096"])]:
    f = lambda a: torch.tensor(np.asarray(a, np.float32), device=eng.device)
    params = eng.cem_sample(f(cfg.mean), f(cfg.std), f(cfg.lo), f(cfg.hi), C, 0, cfg.seed, 0)
    for _ in range(2): eng.evaluate_candidates(params, cfg.names, segs, motor_model=cfg.motor_model)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); n = 5
    for _ in range(n): cost = eng.evaluate_candidates(params, cfg.names, segs, motor_model=cfg.motor_model)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = C * S * 311470.87 / ms * 1e3 / 1e12
    np.save(out / f"cost_{tag}_{C}.npy", cost.cpu().numpy())
    print(f"{tag} minb={os.environ.get("SPI_B200_MINB","auto")} C={C} ms={ms:.3f} env-steps/s={C*S*5/ms*1e3:.3e} "
          f"alg TFLOP/s={tf:.2f} frac={tf/peak:.3f} (peak {peak:.1f})", flush=True)
