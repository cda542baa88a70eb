#!/usr/bin/env python
"""Collect reference data on the B200 engine: the counterpart of the reference's scripts/data/{stand,sine,jump,walk}.py
(one trajectory of the named excitation at nominal URDF parameters, written in the exact .npz schema of
scripts/data/common.py:83-100 under the reference's file names go2_<name>_data.npz), so that the README's sim-to-sim
experiment — record at 6.921 kg, identify the base mass with scripts/mass_opt.py — runs without Isaac Gym.

    python scripts/record_data.py --name all --data-root /tmp/spi_data [--base-mass 6.921]
    python scripts/mass_landscape.py --config all --data-root /tmp/spi_data

(files land in <data-root>/spigym/data/sysid_bag/, where the reference's scripts/config/*.yaml point.)

`walk` is an open-loop trot-shaped law with the reference's 4 x 250-step phase structure (the reference drives an external
unitree_rl_gym policy, scripts/data/walk.py:15-21, which is not available).
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from spi_active_b200 import landscape, recorders  # noqa: E402
from spi_active_b200.dataset import save_recording  # noqa: E402


def record_to_file(name: str, rollout_fn, model, out_dir: Path) -> Path:
    rec = recorders.record(name, rollout_fn, model)
    path = out_dir / f"go2_{name}_data.npz"
    frames = {k: v for k, v in rec.items() if not k.startswith("pd_gain")}
    save_recording(path, frames, recorders.CONTROL_DT, rec["pd_gain_kp"], rec["pd_gain_kd"])
    print(f"Saved {frames['base_positions'].shape[0]} samples to {path}")
    print(f"PD gains - Kp: {rec['pd_gain_kp'][0]:.1f}, Kd: {rec['pd_gain_kd'][0]:.1f}")
    return path


def main() -> None:
    ap = argparse.ArgumentParser(description="Record Go2 reference trajectories on the B200 engine")
    ap.add_argument("--name", default="all", choices=["all", "stand", "sine", "jump", "walk"])
    ap.add_argument("--data-root", type=Path, default=Path.cwd(),
                    help="the files are written to <data-root>/spigym/data/sysid_bag/ (what --data-root of the SysID scripts reads)")
    ap.add_argument("--base-mass", type=float, default=None, help="base-link mass of the recorded robot (default: URDF, 6.921 kg)")
    args = ap.parse_args()
    import torch
    from spi_active_b200.engine import RolloutEngine
    eng = RolloutEngine()
    mass = float(args.base_mass) if args.base_mass is not None else float(eng.model.base.mass)

    def rollout_fn(init: np.ndarray, actions: np.ndarray) -> np.ndarray:
        p = torch.tensor([[mass]], dtype=torch.float32)
        st = eng.rollout_states(p, ["mass"], torch.from_numpy(init)[None], torch.from_numpy(actions)[None])
        return st[0, 0].cpu().numpy()

    out_dir = args.data_root / landscape.DATA_SUBDIR
    out_dir.mkdir(parents=True, exist_ok=True)
    names = recorders.CONFIG_FILES["all"] if args.name == "all" else [args.name]
    for n in names:
        record_to_file(n, rollout_fn, eng.model, out_dir)


if __name__ == "__main__":
    main()
