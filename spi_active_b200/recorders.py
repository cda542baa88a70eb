"""Dataset recorders on the B200 engine: the excitation laws of the reference's data scripts and the
exact .npz schema they write.

Mirrors:
  * scripts/data/stand.py:17-57   zero action, 250 steps @ 50 Hz
  * scripts/data/sine.py:14-45    thigh = +0.8 sin(2 pi 1.0 t), calf = -same, base height 0.32
  * scripts/data/jump.py:14-45    amplitude 1.2, 1.5 Hz
  * scripts/data/walk.py:15-21    1000 steps, 4 command phases of 250 steps.  The reference drives a
                                  unitree_rl_gym policy that is not available; `walk` here is an
                                  open-loop trot-shaped law with the same 4 x 250 phase structure.
  * scripts/data/common.py:83-100 format_data (schema)

Row t of a recording holds (state after env.step(action_t), action_t), exactly like the reference
recorders (scripts/data/stand.py:25-45).
"""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np

from . import go2_model as gm

CONTROL_DT = 0.02  # 200 Hz physics x decimation 4
THIGH_IDX = [1, 4, 7, 10]
CALF_IDX = [2, 5, 8, 11]
BASE_HEIGHT = 0.32
DURATION_STEPS = dict(stand=250, sine=250, jump=250, walk=1000)


def action_law(name: str, steps: int | None = None) -> np.ndarray:
    """actions[T,12] of the named excitation."""
    T = steps or DURATION_STEPS[name]
    t = np.arange(T) * CONTROL_DT
    a = np.zeros((T, 12), dtype=np.float32)
    if name == "stand":
        return a
    if name in ("sine", "jump"):
        amp, freq = (0.8, 1.0) if name == "sine" else (1.2, 1.5)
        phase = amp * np.sin(2 * np.pi * freq * t)
        a[:, THIGH_IDX] = phase[:, None]
        a[:, CALF_IDX] = -phase[:, None]
        return a
    if name == "walk":
        # 4 phases x 250 steps: (frequency Hz, thigh amplitude, calf amplitude, hip sway)
        phases = [(1.5, 0.6, 0.9, 0.0), (2.0, 0.8, 1.2, 0.0), (2.0, 0.6, 0.9, 0.4), (2.5, 0.5, 1.0, -0.4)]
        leg_phase = np.array([0.0, np.pi, np.pi, 0.0])  # trot: FL+RR / FR+RL
        for k in range(T):
            f, at, ac, hip = phases[min(k // 250, 3)]
            for leg in range(4):
                ph = 2 * np.pi * f * t[k] + leg_phase[leg]
                a[k, 3 * leg + 0] = hip * np.sin(ph) * (1.0 if leg % 2 == 0 else -1.0)
                a[k, 3 * leg + 1] = at * np.sin(ph)
                a[k, 3 * leg + 2] = -ac * max(np.sin(ph + 0.5 * np.pi), 0.0)
        return a
    raise KeyError(name)


def initial_state(model: gm.Go2Model | None = None, base_height: float = BASE_HEIGHT) -> np.ndarray:
    m = model or gm.go2_nominal()
    s = np.zeros(gm.STATE_DIM, dtype=np.float32)
    s[2] = base_height
    s[6] = 1.0
    s[13:25] = m.q_default
    return s


def frames_from_states(states: np.ndarray, actions: np.ndarray) -> Dict[str, np.ndarray]:
    """states[T,37] (after each step) + actions[T,12] -> the per-step arrays of the .npz schema."""
    states = np.asarray(states)
    return {
        "joint_positions": states[:, 13:25].astype(np.float32),
        "joint_velocities": states[:, 25:37].astype(np.float32),
        "joint_torques": np.zeros((states.shape[0], 12), dtype=np.float32),
        "actions": np.asarray(actions, dtype=np.float32),
        "base_positions": states[:, 0:3].astype(np.float32),
        "base_orientations": states[:, 3:7].astype(np.float32),
        "base_linear_velocities": states[:, 7:10].astype(np.float32),
        "base_angular_velocities": states[:, 10:13].astype(np.float32),
    }


def record(name: str, rollout_fn: Callable[[np.ndarray, np.ndarray], np.ndarray], model: gm.Go2Model | None = None,
           steps: int | None = None) -> Dict[str, np.ndarray]:
    """Record one trajectory.  `rollout_fn(init[37], actions[T,12]) -> states[T,37]` is the backend:
    RolloutEngine.rollout_states on the GPU (product path) or the oracle in tests."""
    m = model or gm.go2_nominal()
    actions = action_law(name, steps)
    states = np.asarray(rollout_fn(initial_state(m), actions))
    out = frames_from_states(states, actions)
    out["pd_gain_kp"] = np.asarray(m.kp, dtype=np.float32)
    out["pd_gain_kd"] = np.asarray(m.kd, dtype=np.float32)
    return out


def engine_rollout_fn(engine) -> Callable[[np.ndarray, np.ndarray], np.ndarray]:
    """Backend for `record` that runs on the CUDA engine (nominal URDF parameters)."""
    import torch

    def fn(init: np.ndarray, actions: np.ndarray) -> np.ndarray:
        nominal = torch.tensor([[engine.model.base.mass]], dtype=torch.float32)
        st = engine.rollout_states(nominal, ["mass"], torch.from_numpy(init)[None], torch.from_numpy(actions)[None])
        return st[0, 0].cpu().numpy()

    return fn


# file order of scripts/config/all.yaml:2-6
CONFIG_FILES = dict(all=["jump", "sine", "stand", "walk"], jump=["jump"], sine=["sine"], stand=["stand"],
                    walk=["walk"])
