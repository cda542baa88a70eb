"""Synthetic replay datasets for the tests: trajectories recorded with the CPU oracle at nominal
URDF parameters (test infrastructure; the bench records with the CUDA engine instead)."""
from __future__ import annotations

import functools

import numpy as np

from oracle import oracle as orc
from spi_active_b200 import go2_model as gm
from spi_active_b200 import recorders
from spi_active_b200.dataset import concat_windows, window_recording


def oracle_rollout_fn(blob, model, precision=64):
    def fn(init, actions):
        st = orc.rollout_states(blob, np.array([[model.base.mass]], np.float32), [gm.PARAM_IDS["mass"]],
                                init[None], actions[None], precision=precision)
        return st[0, 0]
    return fn


@functools.lru_cache(maxsize=None)
def recording(name: str, steps: int | None = None):
    model = gm.go2_nominal()
    blob = gm.build_model_blob(model)
    return recorders.record(name, oracle_rollout_fn(blob, model), model, steps)


def dataset(config: str = "all", horizon: int = 5, steps: int | None = None):
    """-> (S, dict of numpy arrays) in the reference's load_dataset layout."""
    windows = [window_recording(recording(n, steps), horizon) for n in recorders.CONFIG_FILES[config]]
    return concat_windows(windows)


def pack_numpy(ds: dict):
    """numpy dataset dict -> (seg_init, seg_actions, seg_target, seg_gains, mask_u8, denom)"""
    init = np.concatenate([ds["init_base_pos"], ds["init_base_ori"], ds["init_base_lin_vel"],
                           ds["init_base_ang_vel"], ds["init_joint_pos"], ds["init_joint_vel"]], axis=1)
    tgt = np.concatenate([ds["target_base_pos"], ds["target_base_ori"], ds["target_joint_pos"]], axis=1)
    gains = np.concatenate([ds["pd_gain_kp"], ds["pd_gain_kd"]], axis=1)
    mask = (~ds["motion_ends"]).astype(np.uint8)
    return (init.astype(np.float32), ds["action_sequences"].astype(np.float32), tgt.astype(np.float32),
            gains.astype(np.float32), mask, float(mask.sum()))
