"""dev helper: worst-case deviation of the CUDA path from the fp64 oracle over the whole `all` dataset
(per-sample errors and mean costs), to size the parity tolerances and compare sincos variants."""
import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import synth
from oracle import oracle as orc
from spi_active_b200 import cem, go2_model as gm
from spi_active_b200.dataset import pack_segments, to_device
from spi_active_b200.engine import RolloutEngine

eng = RolloutEngine()
S, ds = synth.dataset("all", 5)
segs = pack_segments(to_device(ds, eng.device))
init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
cfg = cem.default_full_config(eng.model)
ids = [gm.PARAM_IDS[n] for n in cfg.names]
params = np.clip(np.asarray(cfg.mean) + np.random.default_rng(0).standard_normal((16, 10)) * np.asarray(cfg.std), cfg.lo, cfg.hi).astype(np.float32)
cost, per = eng.evaluate_candidates(torch.from_numpy(params), cfg.names, segs, motor_model=cfg.motor_model, return_per_seg=True)
ref, _, ref_per = orc.eval_candidates(eng.blob, params, ids, init, act, tgt, gains, mask, motor_model=3, cost_denominator=denom, return_per_seg=True)
ref32, _, ref_per32 = orc.eval_candidates(eng.blob, params, ids, init, act, tgt, gains, mask, motor_model=3, cost_denominator=denom, return_per_seg=True, precision=32)
d = np.abs(per.cpu().numpy() - ref_per); d32 = np.abs(ref_per32 - ref_per)
print(f"lib={Path(os.environ.get('SPI_B200_LIB','default')).name} kernel={os.environ.get('SPI_B200_KERNEL','ws')}: "
      f"max per-sample |gpu-oracle64| pos/quat/joint = {d[...,0].max():.2e} {d[...,1].max():.2e} {d[...,2].max():.2e}; "
      f"oracle32 noise floor = {d32[...,0].max():.2e} {d32[...,1].max():.2e} {d32[...,2].max():.2e}; "
      f"max rel cost err = {(np.abs(cost.cpu().numpy()-ref)/ref).max():.2e}")
