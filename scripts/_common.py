"""Shared set-up of the two SysID entrypoints: engine, dataset (from .npz recordings or recorded on the GPU)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from spi_active_b200 import landscape, recorders  # noqa: E402
from spi_active_b200.dataset import concat_windows, load_dataset, pack_segments, to_device, window_recording  # noqa: E402
from spi_active_b200.engine import RolloutEngine  # noqa: E402


def add_common_args(parser):
    parser.add_argument("--config", type=str, required=True,
                        help="Config name (e.g., 'walk', 'all') as in the reference's scripts/config/")
    parser.add_argument("--horizon", type=int, default=5, help="Rollout steps between comparisons")
    parser.add_argument("--env-batch", type=int, default=4096, help="Parallel env batch size (chunking of the reference)")
    parser.add_argument("--project-dir", type=Path, default=Path("logs"), help="Top-level directory where runs are written")
    parser.add_argument("--data-root", type=Path, default=None,
                        help="Directory holding spigym/data/sysid_bag/*.npz.  When absent, the trajectories are recorded "
                             "on the GPU with the nominal URDF parameters (sim-to-sim, as the reference's README does)")
    parser.add_argument("--strict-reference", action="store_true",
                        help="Reproduce scripts/eval.py literally: chunk-dependent eval_mask and row-0 PD gains")
    return parser


def build(args):
    """-> (engine, SegmentBatch, num_samples, reference_masses)"""
    np.random.seed(landscape.SEED)
    torch.manual_seed(landscape.SEED)
    engine = RolloutEngine()
    if args.data_root is not None:
        num_samples, dataset_np = load_dataset(landscape.load_config(args.config, args.data_root), args.horizon)
    else:
        fn = recorders.engine_rollout_fn(engine)
        wins = [window_recording(recorders.record(n, fn, engine.model), args.horizon) for n in recorders.CONFIG_FILES[args.config]]
        num_samples, dataset_np = concat_windows(wins)
        print(f"Recorded {len(wins)} trajectory(s) on the GPU with {num_samples} total samples")
    batch_size = min(args.env_batch, num_samples, landscape.MAX_SAFE_ENV_BATCH)   # mass_landscape.py:146
    dataset = to_device(dataset_np, engine.device)
    segs = pack_segments(dataset, env_batch=batch_size, strict_reference=args.strict_reference)
    ref_masses = engine.model.body_masses_isaac_order()                          # capture_reference_masses
    return engine, segs, num_samples, ref_masses
