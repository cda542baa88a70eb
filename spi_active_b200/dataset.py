"""Replay-dataset handling: the reference's on-disk format and windowing, and the packed HBM layout
the rollout kernel reads.

Mirrors (same names, argument meaning and results):
  * scripts/eval.py:21-35    DatasetBatch   (13 tensors; field names are API)
  * scripts/eval.py:38-42    MassEvalResult
  * scripts/eval.py:101-171  load_dataset   (init = row i, target = row i+H, actions[i : i+H],
                                             motion_ends[-1] = True per file, pd gains from .flat[0])
  * scripts/eval.py:174-182  to_device
  * scripts/eval.py:279-280, 304-309  eval_mask / total_valid  -> pack_segments(strict_reference=...)
  * scripts/data/common.py:83-100     format_data (the .npz schema)  -> save_recording

Packed layout (row-major fp32, one row per segment):
  seg_init[S,37] = pos3 quat_xyzw4 linvel3 angvel3 q12 qd12 | seg_actions[S,H,12] | seg_target[S,19] =
  pos3 quat4 q12 | seg_gains[S,24] = kp12 kd12 | seg_mask[S] u8.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np
import torch


@dataclass
class DatasetBatch:
    init_base_pos: torch.Tensor
    target_base_pos: torch.Tensor
    init_base_ori: torch.Tensor
    target_base_ori: torch.Tensor
    init_base_lin_vel: torch.Tensor
    init_base_ang_vel: torch.Tensor
    init_joint_pos: torch.Tensor
    target_joint_pos: torch.Tensor
    init_joint_vel: torch.Tensor
    action_sequences: torch.Tensor
    motion_ends: torch.Tensor
    pd_gain_kp: torch.Tensor  # (num_samples, num_joints)
    pd_gain_kd: torch.Tensor  # (num_samples, num_joints)


@dataclass
class MassEvalResult:
    base_pos: float
    base_quat: float
    joint_pos: float


def window_recording(data: Dict[str, np.ndarray], horizon: int) -> Tuple[Dict[str, np.ndarray], np.ndarray]:
    """One recording -> windowed samples (scripts/eval.py:112-149)."""
    base_positions = np.asarray(data["base_positions"])
    total_steps = base_positions.shape[0]
    num_samples = total_steps - horizon
    if num_samples <= 0:
        raise ValueError(f"recording has {total_steps} steps, need more than horizon={horizon}")
    init_idx = np.arange(num_samples)
    target_idx = init_idx + horizon
    actions = np.asarray(data["actions"]).astype(np.float32)
    joint_positions = np.asarray(data["joint_positions"])
    out = {
        "init_base_pos": base_positions[init_idx],
        "target_base_pos": base_positions[target_idx],
        "init_base_ori": np.asarray(data["base_orientations"])[init_idx],
        "target_base_ori": np.asarray(data["base_orientations"])[target_idx],
        "init_base_lin_vel": np.asarray(data["base_linear_velocities"])[init_idx],
        "init_base_ang_vel": np.asarray(data["base_angular_velocities"])[init_idx],
        "init_joint_pos": joint_positions[init_idx],
        "target_joint_pos": joint_positions[target_idx],
        "init_joint_vel": np.asarray(data["joint_velocities"])[init_idx],
        "action_sequences": np.stack([actions[init_idx + off] for off in range(horizon)], axis=1),
    }
    motion_ends = np.zeros(num_samples, dtype=bool)
    motion_ends[-1] = True
    kp = float(np.asarray(data["pd_gain_kp"]).flat[0])
    kd = float(np.asarray(data["pd_gain_kd"]).flat[0])
    nj = joint_positions.shape[1]
    out["pd_gain_kp"] = np.full((num_samples, nj), kp, dtype=np.float32)
    out["pd_gain_kd"] = np.full((num_samples, nj), kd, dtype=np.float32)
    return out, motion_ends


def concat_windows(windows) -> Tuple[int, Dict[str, np.ndarray]]:
    """Concatenate per-file windows in list order (scripts/eval.py:158-171)."""
    datasets = [w[0] for w in windows]
    out = {k: np.concatenate([d[k] for d in datasets], axis=0) for k in datasets[0].keys()}
    out["motion_ends"] = np.concatenate([w[1] for w in windows], axis=0)
    return out["init_base_pos"].shape[0], out


def load_dataset(paths, horizon: int) -> Tuple[int, Dict[str, np.ndarray]]:
    """Load and concatenate .npz recordings (scripts/eval.py:101-171).  Returns (num_samples, dict)."""
    paths = [paths] if isinstance(paths, (str, Path)) else list(paths)
    windows = []
    for p in paths:
        with np.load(p) as data:
            windows.append(window_recording({k: data[k] for k in data.files}, horizon))
    total, out = concat_windows(windows)
    print(f"Loaded {len(paths)} trajectory(s) with {total} total samples")
    return total, out


def to_device(dataset: Dict[str, np.ndarray], device) -> DatasetBatch:
    """numpy dataset -> torch tensors on device (scripts/eval.py:174-182)."""
    tensors = {}
    for key, value in dataset.items():
        if key == "motion_ends":
            tensors[key] = torch.from_numpy(np.ascontiguousarray(value)).to(device=device, dtype=torch.bool)
        else:
            tensors[key] = torch.from_numpy(np.ascontiguousarray(value)).to(device=device, dtype=torch.float32)
    return DatasetBatch(**tensors)


def save_recording(path, frames: Dict[str, np.ndarray], dt: float, kp, kd) -> None:
    """Write one recording in the reference's .npz schema (scripts/data/common.py:83-100)."""
    T = frames["base_positions"].shape[0]
    out = dict(frames)
    out["timestamps"] = np.arange(T) * dt
    out["sim_duration"] = T * dt
    out["data_frequency"] = int(round(1.0 / dt))
    out["robot_type"] = "go2"
    out["pd_gain_kp"] = np.asarray(kp, dtype=np.float32).reshape(-1)
    out["pd_gain_kd"] = np.asarray(kd, dtype=np.float32).reshape(-1)
    np.savez(path, **out)


# -----------------------------------------------------------------------------------------------
# packed layout for the kernel
# -----------------------------------------------------------------------------------------------
@dataclass
class SegmentBatch:
    seg_init: torch.Tensor      # [S,37]
    seg_actions: torch.Tensor   # [S,H,12]
    seg_target: torch.Tensor    # [S,19]
    seg_gains: torch.Tensor     # [S,24]
    seg_mask: torch.Tensor      # [S] uint8
    cost_denominator: float     # eval.py:304 total_valid

    @property
    def num_segments(self) -> int:
        return int(self.seg_init.shape[0])

    @property
    def horizon(self) -> int:
        return int(self.seg_actions.shape[1])

    def to(self, device) -> "SegmentBatch":
        return SegmentBatch(*(getattr(self, f.name).to(device) if isinstance(getattr(self, f.name), torch.Tensor)
                              else getattr(self, f.name) for f in fields(self)))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.seg_init, self.seg_actions, self.seg_target, self.seg_gains, self.seg_mask))


def reference_eval_mask(motion_ends: torch.Tensor, env_batch: Optional[int], strict_reference: bool) -> torch.Tensor:
    """Which samples count.

    strict_reference=True reproduces scripts/eval.py:279-280 literally: within each chunk of
    `env_batch` samples, every sample at or after the first file boundary of that chunk is masked
    (quirk D2 — chunking dependent).  Default: only the boundary samples themselves are masked.
    """
    me = motion_ends.to(torch.bool)
    if not strict_reference:
        return ~me
    S = me.shape[0]
    B = S if not env_batch else int(env_batch)
    mask = torch.empty_like(me)
    for start in range(0, S, B):
        chunk = me[start:start + B]
        mask[start:start + B] = ~(torch.cumsum(chunk.float(), dim=0) > 0)
    return mask


def pack_segments(batch: DatasetBatch, env_batch: Optional[int] = None, strict_reference: bool = False) -> SegmentBatch:
    """DatasetBatch (reference layout) -> SegmentBatch (kernel layout), on the batch's device.

    strict_reference also reproduces quirk D3 (scripts/eval.py:192-201, 245-250): the PD gains of the
    first sample of each chunk are applied to the whole chunk.
    """
    f32 = torch.float32
    seg_init = torch.cat([batch.init_base_pos, batch.init_base_ori, batch.init_base_lin_vel,
                          batch.init_base_ang_vel, batch.init_joint_pos, batch.init_joint_vel], dim=1).to(f32)
    seg_target = torch.cat([batch.target_base_pos, batch.target_base_ori, batch.target_joint_pos], dim=1).to(f32)
    kp, kd = batch.pd_gain_kp.to(f32), batch.pd_gain_kd.to(f32)
    S = seg_init.shape[0]
    if strict_reference:
        B = S if not env_batch else int(env_batch)
        idx = (torch.arange(S, device=kp.device) // B) * B
        kp, kd = kp[idx], kd[idx]
    seg_gains = torch.cat([kp, kd], dim=1)
    mask = reference_eval_mask(batch.motion_ends, env_batch, strict_reference)
    denom = float((~batch.motion_ends.to(torch.bool)).sum().item())
    return SegmentBatch(seg_init.contiguous(), batch.action_sequences.to(f32).contiguous(), seg_target.contiguous(),
                        seg_gains.contiguous(), mask.to(torch.uint8).contiguous(), denom)
