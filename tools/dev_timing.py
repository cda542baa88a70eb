"""dev helper (not a test): raw kernel timing on the GPU at the bench shape."""
import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from spi_active_b200 import cem, recorders
from spi_active_b200.dataset import pack_segments, to_device
from spi_active_b200.engine import RolloutEngine

eng = RolloutEngine()
S, ds = bench.build_dataset(recorders.engine_rollout_fn(eng), eng.model)
segs = pack_segments(to_device(ds, eng.device))
cfg = cem.default_full_config(eng.model)
tag = f"kernel={os.environ.get('SPI_B200_KERNEL','ws')} minb={os.environ.get('SPI_B200_MINB','-')} lib={Path(os.environ.get('SPI_B200_LIB','default')).name}"
peak = eng.fp32_peak(8192)[0]
for C in [int(x) for x in (sys.argv[1:] or ["1024", "4096"])]:
    opt = cem.CemOptimizer(eng, segs, cfg, C)
    opt.iterate(); params = opt.params
    for _ in range(2): eng.evaluate_candidates(params, cfg.names, segs, motor_model=cfg.motor_model)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); n = 3
    for _ in range(n): cost = eng.evaluate_candidates(params, cfg.names, segs, motor_model=cfg.motor_model)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = C * S * 311470.87 / ms * 1e3 / 1e12
    print(f"{tag} C={C} ms={ms:.3f} env-steps/s={C*S*5/ms*1e3:.3e} alg TFLOP/s={tf:.2f} frac={tf/peak:.3f} (peak {peak:.1f})")
