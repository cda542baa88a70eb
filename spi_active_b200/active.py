"""Active exploration (BASELINE config 5): finite-difference Fisher-information rollouts of command trajectories
under a trained locomotion policy, on the B200 engine.

Mirrors (same names, argument meaning, results):
  * spigym/envs/sysid/active_sysid_openloop.py:29-80     (1 main + P aux) env groups, +delta on one parameter each,
                                                         `params_dict`, main_idx / aux_idx layout
  * :116-131, 196-201                                    reset_all(main_commands) / open-loop command playback
  * :174-187 + go2_omni.py:423-465                        torque law: hip x0.5 -> PD -> clip -> act2tau_* (fused in
                                                         spi_b200_env_step)
  * :203-252                                             _post_physics_step order: commands <- step_idx, termination,
                                                         reward, observations, history push, k-step aux <- main sync
  * :259-272                                             group-wise termination OR
  * :402-426                                             _reward_fisher_information_matrix (spi_b200_fim_reward)
  * spigym/envs/locomotion/go2_omni.py:348-377           _step_contact_targets (gait clock)
  * :627-642                                             eval-mode command_body_height oscillator
  * spigym/envs/legged_base_task/legged_robot_base.py:261-269, 336-339, 511-527, 819-829
                                                         projected gravity / base ang vel, gravity termination,
                                                         sorted-key observation concat, short_history layout
  * spigym/envs/env_utils/history_handler.py:36-44       newest-first history push
  * spigym/utils/helpers.py:77-94                        obs = getter * scale (+ noise, scales are 0 here)
  * spigym/config/obs/loco/go2_omni.yaml                 keys, dims, scales, 14-frame history  -> 900-dim actor input
  * spigym/agents/modules/modules.py:47-63, config/algo/ppo.yaml:32-40   actor MLP 900-512-256-128-12, ELU
  * spigym/agents/sysid/active_sysid.py:528-600          evaluate_policy (1248 steps, zero actions / zero reward for
                                                         terminated groups, mean over steps)
  * :259-400                                             command samplers (constant / polynomial / bezier)
  * :165-242                                             optimize (ask M trials -> evaluate -> tell -reward)

The layer is device-agnostic torch around two engine calls (`env_step`, `fim_reward`), so the host logic runs on a CPU
box against the oracle in tests; the product path is CUDA: physics = spi_b200_env_step, policy = the tcgen05 fp16-pair actor
(spi_b200_policy_forward_ring; torch fp32 GEMMs only for actor shapes it does not support), bookkeeping =
spi_b200_active_post_step, FIM = spi_b200_fim_contract, one CUDA graph per control step.

Reference quirks: D10 absolute delta (default) — `relative_delta=True` for the documented
"10 %"; D11 the reference's k-step sync never reaches the simulator — here it does (the intent), `ksync_steps=0`
disables it; D8 `inertiay` is applied.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import go2_model as gm

# ---- observation layout (go2_omni.yaml; keys concatenated in SORTED order, legged_robot_base.py:521-527) -------------
OBS_DIMS = dict(actions=12, base_ang_vel=3, clock_inputs=4, command_ang_vel=1, command_body_attitude=2,
                command_body_height=1, command_footswing_height=1, command_gait_freq=1, command_gait_phase=4,
                command_lin_vel=2, command_stance=2, dof_pos=12, dof_vel=12, projected_gravity=3)
OBS_SCALES = dict(actions=1.0, base_ang_vel=0.25, clock_inputs=1.0, command_ang_vel=1.0, command_body_attitude=0.3,
                  command_body_height=2.0, command_footswing_height=0.15, command_gait_freq=1.0, command_gait_phase=1.0,
                  command_lin_vel=1.0, command_stance=1.0, dof_pos=1.0, dof_vel=0.05, projected_gravity=1.0)
OBS_KEYS = sorted(OBS_DIMS)
FRAME_DIM = sum(OBS_DIMS.values())          # 60
HISTORY_LEN = 14
ACTOR_OBS_DIM = FRAME_DIM * (1 + HISTORY_LEN)   # 900
CLIP_OBSERVATIONS = 100.0                   # config/env/legged_base.yaml:19-21
BODY_HEIGHT_LIMIT = (-0.25, 0.15)           # obs.commands.limit_body_height
TERMINATION_GRAVITY = (0.8, 0.8)            # config/env/go2_omni.yaml:33-35

# default command vector and ranges: config/algo/active_sysid.yaml:24-45
DEFAULT_COMMAND = [0.5, 0.5, 0.5, 0.0, 3.0, 0.0, 0.5, 0.5, 0.5, 0.03, 0.0, 0.0, 0.3, 0.4]
COMMAND_RANGES = [[-1.0, 1.0], [-1.0, 1.0], [-1.0, 1.0], [-0.25, 0.15], [2.0, 4.0], [0.0, 1.0], [0.0, 1.0], [0.0, 1.0],
                  [0.5, 0.5], [0.03, 0.035], [-0.4, 0.4], [-0.0, 0.0], [0.1, 0.45], [0.35, 0.45]]
COMMAND_SAMPLING_IDXS = [0, 2, 5]
# default parameter set of the env: config/env/active_sysid_openloop.yaml:17-27
DEFAULT_PARAM = dict(mass=9.39, comx=0.0, comy=0.0, comz=0.0, inertiax=0.005, inertiay=0.005, inertiaz=0.005,
                     motor_model_hip_a=20.0, motor_model_thigh_a=20.0, motor_model_calf_a=20.0)


def _history_gather_index() -> torch.Tensor:
    """short_history = for key in sorted(keys): history[key][:, :14].reshape(N, 14 * dim) concatenated
    (legged_robot_base.py:819-829).  With the frames stored as [N, 14, 60] (newest first, sorted-key order inside a
    frame) that is a fixed gather of the flattened [14 * 60] axis."""
    idx, off = [], 0
    for key in OBS_KEYS:
        d = OBS_DIMS[key]
        for t in range(HISTORY_LEN):
            idx.extend(t * FRAME_DIM + off + j for j in range(d))
        off += d
    return torch.tensor(idx, dtype=torch.long)


RING_SLOTS = HISTORY_LEN + 1      # the current frame + the 14 history frames


def ring_col_map() -> np.ndarray:
    """col_map[h, slot * 60 + term] = the actor-input column that a ring of the last 15 frames feeds when its newest
    frame sits in slot h (the head moves DOWN by one slot per step, so slot (h + a) % 15 holds the frame of age a):
    age 0 is the frame block obs[:60], age a >= 1 is row a - 1 of the history, which _history_gather_index scatters over
    the per-key blocks of obs[60:].  Used to permute the first layer's weight columns once per head position instead of
    rebuilding the observation every step (spi_b200_policy_enable_ring)."""
    gather = _history_gather_index().numpy()                 # obs[60 + j] = hist_flat[gather[j]]
    where = np.empty(HISTORY_LEN * FRAME_DIM, dtype=np.int64)
    where[gather] = np.arange(HISTORY_LEN * FRAME_DIM)       # hist_flat[f] feeds obs[60 + where[f]]
    cm = np.full((RING_SLOTS, ACTOR_OBS_DIM), -1, dtype=np.int32)
    for h in range(RING_SLOTS):
        for s in range(RING_SLOTS):
            a = (s - h) % RING_SLOTS
            for i in range(FRAME_DIM):
                cm[h, s * FRAME_DIM + i] = i if a == 0 else FRAME_DIM + where[(a - 1) * FRAME_DIM + i]
    return cm


def quat_rotate_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """spigym/utils/torch_utils.py:83-92 (q = xyzw)."""
    q_w = q[:, 3:4]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0)
    b = torch.cross(q_vec, v, dim=-1) * q_w * 2.0
    c = q_vec * (q_vec * v).sum(dim=1, keepdim=True) * 2.0
    return a - b + c


def step_contact_targets(gait_indices: torch.Tensor, commands: torch.Tensor, dt: float):
    """go2_omni.py:348-377: advance the gait phase, warp the four foot phases by the stance duration, clock = sin.
    Returns (new gait_indices[N], clock_inputs[N,4])."""
    frequencies, phases, offsets, bounds, durations = (commands[:, 4], commands[:, 5], commands[:, 6], commands[:, 7],
                                                       commands[:, 8])
    gait_indices = torch.remainder(gait_indices + dt * frequencies, 1.0)
    foot = torch.stack([gait_indices + phases + offsets + bounds, gait_indices + offsets, gait_indices + bounds,
                        gait_indices + phases], dim=1)
    r = torch.remainder(foot, 1.0)
    d = durations[:, None]
    stance, swing = r < d, r > d
    warped = torch.where(stance, r * (0.5 / d), foot)
    warped = torch.where(swing, 0.5 + (r - d) * (0.5 / (1 - d)), warped)
    return gait_indices, torch.sin(2 * math.pi * warped)


def build_frame(state: torch.Tensor, actions: torch.Tensor, commands: torch.Tensor, clock: torch.Tensor,
                gait_indices: torch.Tensor, q_default: torch.Tensor) -> torch.Tensor:
    """The 60 scaled per-step observation terms in sorted-key order (what the policy sees first and what is pushed
    into the history)."""
    quat = state[:, 3:7]
    ang = quat_rotate_inverse(quat, state[:, 10:13])
    grav = torch.zeros_like(ang)
    grav[:, 2] = -1.0
    pg = quat_rotate_inverse(quat, grav)
    height = torch.clamp(commands[:, 3:4] + 0.20 * torch.sin(2 * math.pi * gait_indices[:, None]),
                         min=BODY_HEIGHT_LIMIT[0], max=BODY_HEIGHT_LIMIT[1])        # go2_omni.py:627-642 (eval mode)
    terms = dict(actions=actions, base_ang_vel=ang, clock_inputs=clock, command_ang_vel=commands[:, 2:3],
                 command_body_attitude=commands[:, 10:12], command_body_height=height,
                 command_footswing_height=commands[:, 9:10], command_gait_freq=commands[:, 4:5],
                 command_gait_phase=commands[:, 5:9], command_lin_vel=commands[:, 0:2], command_stance=commands[:, 12:14],
                 dof_pos=state[:, 13:25] - q_default, dof_vel=state[:, 25:37], projected_gravity=pg)
    return torch.cat([terms[k] * OBS_SCALES[k] for k in OBS_KEYS], dim=1)


class PolicyMLP:
    """actor: Linear(900,512) ELU Linear(512,256) ELU Linear(256,128) ELU Linear(128,12); inference = mean action
    (agents/modules/ppo_modules.py:73-75)."""

    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]):
        self.weights = [w.contiguous() for w in weights]
        self.biases = [b.contiguous() for b in biases]

    @classmethod
    def random(cls, device, seed: int = 0, dims=(ACTOR_OBS_DIM, 512, 256, 128, 12), gain: float = 0.5):
        g = torch.Generator().manual_seed(seed)
        ws, bs = [], []
        for i in range(len(dims) - 1):
            ws.append((torch.randn(dims[i + 1], dims[i], generator=g) * gain / math.sqrt(dims[i])).to(device))
            bs.append(torch.zeros(dims[i + 1]).to(device))
        return cls(ws, bs)

    @classmethod
    def from_checkpoint(cls, path, device):
        """PPO checkpoint: `actor_model_state_dict` (agents/ppo/ppo.py:138-143), Sequential keys `...module.{0,2,4,6}`."""
        sd = torch.load(path, map_location="cpu")["actor_model_state_dict"]
        w = sorted((k for k in sd if k.endswith(".weight")), key=lambda k: int(k.split(".")[-2]))
        b = sorted((k for k in sd if k.endswith(".bias")), key=lambda k: int(k.split(".")[-2]))
        return cls([sd[k].float().to(device) for k in w], [sd[k].float().to(device) for k in b])

    def __call__(self, obs: torch.Tensor) -> torch.Tensor:
        x = obs
        n = len(self.weights)
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            x = torch.addmm(b, x, w.t())
            if i < n - 1:
                x = torch.nn.functional.elu(x)
        return x


@dataclass
class ActiveConfig:
    exploration_params: List[str] = field(default_factory=lambda: ["mass"])
    default_param: Dict[str, float] = field(default_factory=lambda: dict(DEFAULT_PARAM))
    delta_param: float = 0.1
    relative_delta: bool = False          # quirk D10
    ksync_steps: int = 5
    motor_model: str = "act2tau_vec3_tanh"
    rollout_length: float = 25.0          # s  -> total_steps = 1250
    action_clip: float = 20.0
    termination_rew: float = 0.0
    randomize_reset: bool = True          # legged_robot_base.py:737-784 reset distribution
    seed: int = 0
    # "step": J J^T per control step on CUDA cores (spi_b200_fim_reward); "tensor": the states of `fim_chunk` steps are
    # kept on device and contracted on the tensor cores (spi_b200_fim_contract); "fused": accumulated inside the post-step
    # kernel from the rows it already holds (no state history); "auto": tensor on a CUDA backend
    fim_mode: str = "auto"
    fim_chunk: int = 64
    # "torch": the post-physics bookkeeping of a step as ~130 torch ops (device-agnostic statement, runs on the CPU
    # oracle in tests); "fused": one CUDA kernel (spi_b200_active_post_step, needs fim_mode tensor); "auto": fused on CUDA
    step_impl: str = "auto"
    # actor evaluation in the fused step: "tensor" = spi_b200_policy_forward (tcgen05, 3xTF32: fp32-grade accuracy),
    # "cublas" = torch fp32 GEMMs; "auto" = tensor when the actor has the supported shape (the reference's does)
    policy_impl: str = "auto"
    # with the tensor-core actor: keep the observation as a ring of 15 pre-split frames and permute the first layer's weight
    # columns per head position instead of rebuilding the 900-dim observation every step (spi_b200_policy_forward_ring);
    # False = the post-step kernel materialises obs / history like the reference does
    obs_ring: bool = True


class ActiveExploration:
    """(1 main + P aux) env groups on the engine; evaluate_policy(commands[M,T,14]) -> FIM rewards."""

    PARAM_ORDER = ["mass", "comx", "comy", "comz", "inertiax", "inertiay", "inertiaz", "motor_model_hip_a",
                   "motor_model_thigh_a", "motor_model_calf_a"]

    def __init__(self, backend, policy: PolicyMLP, num_main_envs: int, cfg: Optional[ActiveConfig] = None,
                 device=None, model: Optional[gm.Go2Model] = None):
        self.backend, self.policy, self.cfg = backend, policy, cfg or ActiveConfig()
        self.model = model or getattr(backend, "model", None) or gm.go2_nominal()
        self.device = torch.device(device if device is not None else getattr(backend, "device", "cpu"))
        c = self.cfg
        self.param_dim = len(c.exploration_params)
        self.num_main_envs = int(num_main_envs)
        self.num_envs = self.num_main_envs * (self.param_dim + 1)
        self.dt = self.model.dt * self.model.control_decimation
        self.total_steps = int(c.rollout_length / self.dt)
        # per-env parameter table (active_sysid_openloop.py:51-68): env (i + 1) mod (P + 1) gets +delta on parameter i
        names = [n for n in self.PARAM_ORDER if n in c.default_param]
        self.param_names = names
        table = np.tile(np.array([c.default_param[n] for n in names], dtype=np.float64)[None], (self.num_envs, 1))
        slot = np.arange(self.num_envs) % (self.param_dim + 1)
        self.deltas = []
        for i, p in enumerate(c.exploration_params):
            col = names.index(p)
            d = c.delta_param * abs(c.default_param[p]) if c.relative_delta else c.delta_param
            table[slot == i + 1, col] += d
            self.deltas.append(d)
        self.params = torch.tensor(table, dtype=torch.float32, device=self.device)
        self.params_dict = {n: {"value": table[:, j].tolist()} for j, n in enumerate(names)}   # reference layout
        idx = torch.arange(self.num_envs, device=self.device).view(self.num_main_envs, self.param_dim + 1)
        self.main_idx, self.aux_idx = idx[:, 0:1], idx[:, 1:]
        self.q_default = torch.tensor(self.model.q_default, dtype=torch.float32, device=self.device)
        self.hist_index = _history_gather_index().to(self.device)
        N = self.num_envs
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.state, self.actions, self.obs = z(N, gm.STATE_DIM), z(N, 12), z(N, ACTOR_OBS_DIM)
        self.history, self.gait_indices, self.clock = z(N, HISTORY_LEN, FRAME_DIM), z(N), z(N, 4)
        self.commands, self.next_commands = z(N, 14), z(N, 14)
        self.done = torch.zeros(N, dtype=torch.bool, device=self.device)
        self.total_reward, self.step_reward = z(N), z(N)
        self.jtj = z(self.num_main_envs, self.param_dim, self.param_dim)
        self.sync_flag = torch.zeros((), dtype=torch.bool, device=self.device)
        self._graph = None
        mode = c.fim_mode
        if mode == "auto":
            # "fused" where the post-step kernel exists: the contraction's useful work is 2 P^2 25 flop per group-step (6.4 GFLOP
            # per 1 024-trial rollout) against 1.41 GB of recorded states — memory-bound by two orders of magnitude, so the best
            # kernel is the one that never writes the states out; "tensor" (tcgen05) stays available and tested
            if self.device.type == "cuda" and hasattr(backend, "active_post_step") and c.step_impl in ("auto", "fused"):
                mode = "fused"
            else:
                mode = "tensor" if (self.device.type == "cuda" and hasattr(backend, "fim_contract")) else "step"
        if mode not in ("step", "tensor", "fused"):
            raise ValueError(f"fim_mode must be 'auto', 'step', 'tensor' or 'fused', not {c.fim_mode!r}")
        if mode == "fused" and not hasattr(backend, "active_post_step"):
            raise ValueError("fim_mode='fused' needs a backend with active_post_step")
        if mode == "tensor" and not hasattr(backend, "fim_contract"):
            raise ValueError("fim_mode='tensor' needs a backend with fim_contract")
        if mode == "tensor" and self.param_dim > 16:
            raise ValueError("fim_mode='tensor' supports at most 16 exploration parameters")
        self.fim_mode = mode
        if mode == "fused":       # J J^T accumulated inside spi_b200_active_post_step: no state history at all
            self.hist = self.live_hist = None
            self.trace_acc, self.dead_steps = z(self.num_main_envs), z(N)
        if mode == "tensor":
            K = max(1, int(c.fim_chunk))
            self.hist = z(K, self.num_main_envs, self.param_dim + 1, 25)
            self.live_hist = torch.zeros(K, self.num_main_envs, dtype=torch.uint8, device=self.device)
            self.slot = torch.zeros(1, dtype=torch.long, device=self.device)
            self.trace_acc, self.dead_steps = z(self.num_main_envs), z(N)
            self._hist_count = 0
        impl = c.step_impl
        if impl == "auto":
            impl = "fused" if (mode in ("tensor", "fused") and hasattr(backend, "active_post_step")) else "torch"
        if impl not in ("torch", "fused"):
            raise ValueError(f"step_impl must be 'auto', 'torch' or 'fused', not {c.step_impl!r}")
        if impl == "fused" and not (mode in ("tensor", "fused") and hasattr(backend, "active_post_step")):
            raise ValueError("step_impl='fused' needs fim_mode='tensor' or 'fused' and a backend with active_post_step")
        if mode == "fused" and impl != "fused":
            raise ValueError("fim_mode='fused' needs step_impl='fused'")
        self.step_impl = impl
        if impl == "fused":
            self.hist_index_i32 = self.hist_index.to(torch.int32).contiguous()
            self.counter = torch.zeros(2, dtype=torch.int32, device=self.device)   # step counter, block ticket of the post-step kernel
            self.ctrl = torch.zeros(4, dtype=torch.int32, device=self.device)
            self.zero_actions = z(N, 12)
            self.tc_policy = self.obs_hi = self.obs_lo = None
            if c.policy_impl not in ("auto", "tensor", "cublas"):
                raise ValueError(f"policy_impl must be 'auto', 'tensor' or 'cublas', not {c.policy_impl!r}")
            if c.policy_impl != "cublas" and hasattr(backend, "tensor_core_policy"):
                try:
                    self.tc_policy = backend.tensor_core_policy(policy.weights, policy.biases)
                except Exception as exc:
                    if c.policy_impl == "tensor":
                        raise
                    import warnings
                    warnings.warn(f"tensor-core actor unavailable for this policy ({exc}); using torch fp32 GEMMs", RuntimeWarning)
            elif c.policy_impl == "tensor":
                raise ValueError("policy_impl='tensor' needs a backend with tensor_core_policy")
            self.ring = False
            if self.tc_policy is not None:
                self.obs_hi, self.obs_lo = self.tc_policy.alloc_input(N)
                self.raw_actions = z(N, 12)
                if c.obs_ring and self.tc_policy.dims[0] == ACTOR_OBS_DIM:
                    self.tc_policy.enable_ring(ring_col_map())
                    self.ring = True

    # ---- physics / reward through the engine -------------------------------------------------------------------------
    def _physics(self):
        self.backend.env_step(self.state, self.actions, params=self.params, param_names=self.param_names,
                              motor_model=self.cfg.motor_model, flags=gm.FLAG_HIP_HALF)

    def _fim(self):
        """rew[N] = trace(J^T J) of the group, repeated to its P + 1 envs (active_sysid_openloop.py:402-426).  The
        finite-difference divisor is the single configured delta, as in the reference (:412, :417)."""
        M, P1 = self.num_main_envs, self.param_dim + 1
        states = self.state[:, :25].reshape(M, P1, 25).contiguous()      # root13 (origins are 0) + q12
        jtj, trace = self.backend.fim_reward(states, float(self.cfg.delta_param))
        return trace.repeat_interleave(P1), jtj

    # ---- one control step (everything between two policy evaluations) -------------------------------------------------------
    def _env_step(self, actions: torch.Tensor):
        c = self.cfg
        # go2_omni.step: gait clock first, with the commands of the current step_idx
        self.gait_indices, self.clock = step_contact_targets(self.gait_indices, self.commands, self.dt)
        self.actions.copy_(torch.clip(actions, -c.action_clip, c.action_clip))     # _pre_physics_step
        self._physics()                                                           # _physics_step
        # _post_physics_step
        self.commands.copy_(self.next_commands)                                   # _update_tasks_callback
        grav = torch.zeros_like(self.state[:, 0:3]); grav[:, 2] = -1.0
        pg = quat_rotate_inverse(self.state[:, 3:7], grav)
        reset = (pg[:, 0].abs() > TERMINATION_GRAVITY[0]) | (pg[:, 1].abs() > TERMINATION_GRAVITY[1])
        reset = reset | ~torch.isfinite(self.state).all(dim=1)
        group = reset.view(self.num_main_envs, -1).any(dim=1, keepdim=True)        # _update_reset_buf group OR
        self.done.copy_(group.expand(-1, self.param_dim + 1).reshape(-1))
        if self.fim_mode == "tensor":                                            # _compute_reward, deferred:
            self._record_fim_inputs(group)                                        # contracted every fim_chunk steps
            rew = jtj = None
        else:
            rew, jtj = self._fim()
        frame = build_frame(self.state, self.actions, self.commands, self.clock, self.gait_indices, self.q_default)
        hist_flat = self.history.reshape(self.num_envs, HISTORY_LEN * FRAME_DIM)
        obs = torch.cat([frame, hist_flat[:, self.hist_index]], dim=1)             # _compute_observations
        self.obs.copy_(torch.clip(obs, -CLIP_OBSERVATIONS, CLIP_OBSERVATIONS))
        self.history.copy_(torch.cat([frame[:, None, :], self.history[:, :-1, :]], dim=1))   # history_handler.add
        # k-step synchronisation aux <- main (the intent of :247-252, 316-330; quirk D11)
        grp = self.state.view(self.num_main_envs, self.param_dim + 1, gm.STATE_DIM)
        synced = grp[:, 0:1, :].expand(-1, self.param_dim + 1, -1).reshape(self.num_envs, gm.STATE_DIM)
        self.state.copy_(torch.where(self.sync_flag, synced, self.state))
        if self.fim_mode == "tensor":
            return
        # active_sysid.py:567-577: terminated groups score termination_rew
        rew = torch.where(self.done, torch.full_like(rew, c.termination_rew), torch.nan_to_num(rew, nan=0.0, posinf=0.0))
        self.step_reward.copy_(rew)
        self.total_reward.add_(rew)
        live = (~self.done.view(self.num_main_envs, -1)[:, 0]).to(jtj.dtype)
        self.jtj.add_(torch.nan_to_num(jtj) * live[:, None, None])

    # ---- tensor-core FIM: record now, contract later ---------------------------------------------------------------------
    def _record_fim_inputs(self, group_done: torch.Tensor):
        """Slot `self.slot` of the history ring <- this step's (root13, q12) of every env and the group's live flag.
        Runs inside the captured step; the slot index is a device scalar the host advances between replays."""
        M, P1 = self.num_main_envs, self.param_dim + 1
        self.hist.index_copy_(0, self.slot, self.state[:, :25].reshape(1, M, P1, 25))
        self.live_hist.index_copy_(0, self.slot, (~group_done).reshape(1, M).to(torch.uint8))
        self.dead_steps.add_(self.done.to(self.dead_steps.dtype))

    def _flush_fim(self):
        """sum over the recorded steps of live * J J^T -> self.jtj, trace -> self.trace_acc (spi_b200_fim_contract)."""
        n = self._hist_count
        if n == 0:
            return
        self.backend.fim_contract(self.hist[:n], float(self.cfg.delta_param), live=self.live_hist[:n],
                                  out_JtJ=self.jtj, out_trace=self.trace_acc, accumulate=True)
        self._hist_count = 0

    def _policy_step(self):
        if self.step_impl == "fused":
            if self.tc_policy is not None and self.ring:
                return self._fused_step(self.tc_policy.forward_ring(self.obs_hi, self.obs_lo, self.num_envs,
                                                                    self.ctrl[3:4], out=self.raw_actions))
            if self.tc_policy is not None:
                return self._fused_step(self.tc_policy.forward_split(self.obs_hi, self.obs_lo, self.num_envs,
                                                                     out=self.raw_actions))
            return self._fused_step(self.policy(self.obs))
        actions = self.policy(self.obs)
        actions = torch.where(self.done[:, None], torch.zeros_like(actions), actions)   # active_sysid.py:559-562
        self._env_step(actions)

    # ---- fused CUDA step: policy output -> physics (terminated envs run with a zero action) -> one post-step kernel -----------
    def _fused_step(self, raw_actions: torch.Tensor):
        c = self.cfg
        self.backend.env_step(self.state, raw_actions, params=self.params, param_names=self.param_names,
                              motor_model=c.motor_model, flags=gm.FLAG_HIP_HALF, zero_action_mask=self.done)
        ring = getattr(self, "ring", False)
        self.backend.active_post_step(self.state, raw_actions, self.done, self.main_commands, self.commands,
                                      self.actions, self.gait_indices, self.clock, None if ring else self.history,
                                      None if ring else self.obs, None if ring else self.hist_index_i32, self.hist,
                                      self.live_hist, self.dead_steps, self.schedule,
                                      self.counter, self.ctrl, self.dt, c.action_clip, CLIP_OBSERVATIONS,
                                      TERMINATION_GRAVITY, self.model.q_default, obs_hi=self.obs_hi, obs_lo=self.obs_lo,
                                      ring_slots=RING_SLOTS if ring else 0,
                                      fim_jtj=self.jtj if self.fim_mode == "fused" else None,
                                      fim_trace=self.trace_acc if self.fim_mode == "fused" else None,
                                      fim_delta=float(c.delta_param))

    # ---- ring mode: the plain observation / history exist only on request ------------------------------------------------
    def materialize_observation(self):
        """Ring mode: self.obs [N,900] and self.history [N,14,60] <- what the ring holds (clipped frames; the reference
        clips when it builds the observation, so obs is identical and history differs only beyond +-clip)."""
        if not getattr(self, "ring", False):
            return self.obs
        N = self.num_envs
        R = self.tc_policy.unsplit_input(self.obs_hi, self.obs_lo, N).view(N, -1)[:, :RING_SLOTS * FRAME_DIM]
        R = R.reshape(N, RING_SLOTS, FRAME_DIM)
        h = int(self.ctrl[3].item())
        order = [(h + a) % RING_SLOTS for a in range(RING_SLOTS)]
        frames = R[:, order, :]                                      # age 0 .. 14
        # the observation reads the history BEFORE this step's frame was pushed (ages 1 .. 14); the stored history
        # already holds it (ages 0 .. 13)
        self.history.copy_(frames[:, :HISTORY_LEN, :])
        before = frames[:, 1:, :].reshape(N, HISTORY_LEN * FRAME_DIM)
        self.obs.copy_(torch.cat([frames[:, 0, :], before[:, self.hist_index]], dim=1))
        return self.obs

    def load_observation(self, obs: torch.Tensor, history: torch.Tensor):
        """Ring mode: ring <- the 15 frames behind (obs [N,900], history [N,14,60] AFTER the push of obs' frame) at the
        current head position: ages 0 .. 13 are the history rows, age 14 is the oldest row of the pre-push history,
        which only the observation still carries."""
        assert getattr(self, "ring", False)
        N = self.num_envs
        h = int(self.ctrl[3].item())
        before = torch.zeros(N, HISTORY_LEN * FRAME_DIM, device=self.device)
        before[:, self.hist_index] = obs[:, FRAME_DIM:]
        oldest = before.view(N, HISTORY_LEN, FRAME_DIM)[:, HISTORY_LEN - 1:, :]
        frames = torch.cat([history, oldest], dim=1).clamp(-CLIP_OBSERVATIONS, CLIP_OBSERVATIONS)
        R = torch.zeros(N, RING_SLOTS, FRAME_DIM, device=self.device)
        R[:, [(h + a) % RING_SLOTS for a in range(RING_SLOTS)], :] = frames
        self.tc_policy.split_input(R.reshape(N, RING_SLOTS * FRAME_DIM).contiguous(), self.obs_hi, self.obs_lo)

    def _build_schedule(self, n_calls: int, T: int):
        """Row i = the host inputs of the (i + 1)-th env step after a reset: (command row, k-sync flag, FIM ring slot)
        — what _advance_inputs computes step by step, uploaded once."""
        k, K = self.cfg.ksync_steps, (self.hist.shape[0] if self.hist is not None else 1)
        n_rows = n_calls + 4              # + the rows the graph warm-up / capture steps read past the end
        key = (n_rows, T, k, K)
        if getattr(self, "_schedule_key", None) != key:      # same rollout shape as last time: the rows are already on the device
            i = np.arange(n_rows)
            idx = i + 1                                                      # step_idx after the increment
            rows = np.zeros((n_rows, 4), dtype=np.int32)
            rows[:, 0] = np.minimum(idx, T - 1)
            rows[:, 1] = 1 if k == 1 else ((idx % k == 1) if k > 1 else 0)
            rows[:, 2] = np.where(i == 0, 0, (i - 1) % K)                    # call 0 is the reset step (dropped)
            rows[:, 3] = (-i) % RING_SLOTS                                   # observation ring: the head moves down one slot per step
            if getattr(self, "schedule", None) is None or self.schedule.shape[0] < n_rows:
                self.schedule = torch.zeros((n_rows, 4), dtype=torch.int32, device=self.device)
                self._graph = None        # a captured step holds the old buffer's address
            self.schedule[:n_rows].copy_(torch.from_numpy(rows))
            self._schedule_key = key
        self.counter.zero_()

    # ---- reset ---------------------------------------------------------------------------------------------------------------
    @staticmethod
    def initial_main_states(M: int, model, cfg: "ActiveConfig") -> torch.Tensor:
        """[M, 37] reset states of the main envs (go2.yaml:52-55 + the reset distribution of
        legged_robot_base.py:737-784), seeded: trial j of a population gets the same draw however the population is cut."""
        g = torch.Generator().manual_seed(cfg.seed)
        s = torch.zeros(M, gm.STATE_DIM)
        s[:, 2], s[:, 6] = 0.34, 1.0                                       # go2.yaml:52-55
        s[:, 13:25] = torch.tensor(model.q_default)
        if cfg.randomize_reset:                                            # legged_robot_base.py:737-784
            s[:, 7:13] = torch.rand(M, 6, generator=g) - 0.5
            s[:, 13:25] *= 0.5 + torch.rand(M, 12, generator=g)
        return s

    def reset_all(self, main_commands: torch.Tensor, total_steps: Optional[int] = None,
                  initial_main_states: Optional[torch.Tensor] = None, materialize: bool = True):
        """active_sysid_openloop.py:116-131 + base_task.py:90-100: reset, then ONE env step with zero actions.
        `materialize=False` skips the return value (in ring mode it costs a host sync and a rebuild of the 900-dim rows)."""
        c, N, P1 = self.cfg, self.num_envs, self.param_dim + 1
        if getattr(self, "main_commands", None) is None or self.main_commands.shape != main_commands.shape:
            self.main_commands = torch.empty(tuple(main_commands.shape), dtype=torch.float32, device=self.device)
            self._graph = None            # persistent buffer: a captured step reads the command rows from it
        # straight into the persistent buffer; asynchronous on this stream when the caller's tensor is pinned host memory
        self.main_commands.copy_(main_commands, non_blocking=True)
        fused = self.step_impl == "fused"
        if not fused:
            self.expanded_main_commands = self.main_commands.repeat_interleave(P1, dim=0)
        self.step_idx = 0
        self.commands.copy_(self.main_commands[:, 0, :].repeat_interleave(P1, dim=0))
        M = self.num_main_envs
        s = initial_main_states if initial_main_states is not None else self.initial_main_states(M, self.model, c)
        assert tuple(s.shape) == (M, gm.STATE_DIM)
        self.state.copy_(s.repeat_interleave(P1, dim=0).to(self.device))    # every env of a group starts from the main's draw
        for t in (self.actions, self.history, self.gait_indices, self.clock, self.total_reward, self.jtj, self.step_reward):
            t.zero_()
        self.done.zero_()
        if fused and self.tc_policy is not None:
            self.obs_hi.zero_(); self.obs_lo.zero_()
            self.ctrl.zero_()
        if fused:
            self._build_schedule(int(total_steps or self.total_steps), self.main_commands.shape[1])
            # the gait clock of the first step (go2_omni.step); afterwards the post-step kernel advances it
            g, clk = step_contact_targets(self.gait_indices, self.commands, self.dt)
            self.gait_indices.copy_(g); self.clock.copy_(clk)
            self.step_idx += 1
            self._fused_step(self.zero_actions)
            return self.materialize_observation() if materialize else None
        self._advance_inputs()
        self._env_step(torch.zeros(N, 12, device=self.device))
        return self.obs

    def _advance_inputs(self):
        """Host-side per-step inputs of the captured step: next command row and the k-sync flag."""
        self.step_idx += 1                                                  # _pre_physics_step increments step_idx
        t = min(self.step_idx, self.expanded_main_commands.shape[1] - 1)
        self.next_commands.copy_(self.expanded_main_commands[:, t, :])
        k = self.cfg.ksync_steps
        sync = (k == 1) or (k > 1 and self.step_idx % k == 1)
        self.sync_flag.fill_(bool(sync))
        if self.fim_mode == "tensor":
            self.slot.fill_(self._hist_count)

    # ---- evaluate_policy -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def evaluate_policy(self, commands: torch.Tensor, total_steps: Optional[int] = None, use_cuda_graph: Optional[bool] = None):
        """commands[M, T, 14] -> {"total_reward": [N] mean FIM reward per step (main env j at j * (P + 1)),
        "fim": [M, P, P] accumulated J J^T / steps}  (active_sysid.py:528-600)."""
        n = self.begin_rollout(commands, total_steps, use_cuda_graph)
        for _ in range(n):
            self.advance_rollout()
        return self.finish_rollout()

    # the three pieces of evaluate_policy, exposed so that PipelinedExploration can interleave several explorers
    @torch.no_grad()
    def begin_rollout(self, commands: torch.Tensor, total_steps: Optional[int] = None, use_cuda_graph: Optional[bool] = None,
                      initial_main_states: Optional[torch.Tensor] = None) -> int:
        """Reset + the reset step (+ capture of the step graph); returns the number of advance_rollout() calls of the rollout."""
        total_steps = int(total_steps or self.total_steps)
        assert commands.shape[0] == self.num_main_envs and commands.shape[2] == 14
        self.reset_all(commands, total_steps, initial_main_states, materialize=False)
        self.total_reward.zero_(); self.jtj.zero_()
        if self.fim_mode in ("tensor", "fused"):      # the reset step's record / contribution is dropped, like its reward (:545-548)
            self.trace_acc.zero_(); self.dead_steps.zero_()
            self._hist_count = 0
        self._graph_ok = self.device.type == "cuda" if use_cuda_graph is None else use_cuda_graph
        self._steps_done = 1
        self._steps_scheduled = total_steps
        if self._graph_ok and self._graph is None:
            self._capture()
        return max(0, total_steps - 2)                # the reference's loop is range(1, total_steps - 1)

    @torch.no_grad()
    def advance_rollout(self):
        assert self._steps_done < self._steps_scheduled, "advance_rollout() called more often than begin_rollout() scheduled"
        if self.step_impl == "fused":
            self.step_idx += 1
        else:
            self._advance_inputs()
        if self._graph_ok:
            self._graph.replay()
        else:
            self._policy_step()
        self._steps_done += 1
        if self.fim_mode == "tensor":
            self._hist_count += 1
            if self._hist_count == self.hist.shape[0]:
                self._flush_fim()

    @torch.no_grad()
    def finish_rollout(self):
        step = self._steps_done
        if self.fim_mode in ("tensor", "fused"):
            if self.fim_mode == "tensor":
                self._flush_fim()
            self.total_reward.copy_(self.trace_acc.repeat_interleave(self.param_dim + 1)
                                    + self.cfg.termination_rew * self.dead_steps)
        return {"total_reward": (self.total_reward / step).cpu().numpy(), "fim": (self.jtj / step).cpu().numpy(),
                "steps": step}

    def _capture(self):
        """One control step (policy + clock + physics + reward + observation + sync) as a CUDA graph."""
        live = [self.state, self.actions, self.obs, self.history, self.gait_indices, self.clock, self.commands,
                self.total_reward, self.jtj, self.step_reward]
        if self.fim_mode == "tensor":
            live += [self.dead_steps]                 # hist / live_hist slots are rewritten before they are read
        if self.fim_mode == "fused":
            live += [self.dead_steps, self.trace_acc]
        if self.step_impl == "fused":
            live += [self.counter, self.ctrl]         # the warm-up / capture steps must not consume schedule rows
            if self.tc_policy is not None:
                live += [self.obs_hi, self.obs_lo]    # what the tensor-core actor actually reads
        saved = [t.clone() for t in live]
        saved_done = self.done.clone()
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._policy_step()
        torch.cuda.current_stream(self.device).wait_stream(s)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._policy_step()
        for t, v in zip(live, saved):
            t.copy_(v)
        self.done.copy_(saved_done)


# ---- command samplers (active_sysid.py:259-400) ------------------------------------------------------------------------------
class PipelinedExploration:
    """The M trials of an evaluate_policy call cut into `n_pipelines` independent explorers whose captured control steps
    are replayed round-robin on their own CUDA streams.  A control step is a serial chain (actor -> physics -> post-step)
    whose physics / bookkeeping part is latency-bound (0.3 - 0.6 waves of the GPU); two chains that are free to drift
    against each other fill those gaps with the other chain's actor GEMMs.  Trials are independent, the initial states
    are drawn once for the whole population and sliced, and every kernel is deterministic per row, so the result is the
    same as ActiveExploration's on the uncut population."""

    def __init__(self, backend, policy: "PolicyMLP", num_main_envs: int, cfg: Optional["ActiveConfig"] = None,
                 n_pipelines: int = 2, **kw):
        n = max(1, min(int(n_pipelines), int(num_main_envs)))
        cuts = [int(num_main_envs) * i // n for i in range(n + 1)]
        self.slices = [slice(cuts[i], cuts[i + 1]) for i in range(n)]
        self.subs = [ActiveExploration(backend, policy, sl.stop - sl.start, cfg, **kw) for sl in self.slices]
        e = self.subs[0]
        self.cfg, self.model, self.device, self.dt = e.cfg, e.model, e.device, e.dt
        self.param_dim, self.param_names, self.total_steps = e.param_dim, e.param_names, e.total_steps
        self.num_main_envs = int(num_main_envs)
        self.num_envs = self.num_main_envs * (self.param_dim + 1)
        self.fim_mode, self.step_impl = e.fim_mode, e.step_impl
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.subs] if self.device.type == "cuda" else None

    @torch.no_grad()
    def evaluate_policy(self, commands: torch.Tensor, total_steps: Optional[int] = None, use_cuda_graph: Optional[bool] = None):
        assert commands.shape[0] == self.num_main_envs
        init = ActiveExploration.initial_main_states(self.num_main_envs, self.model, self.cfg)
        import contextlib
        if self.streams is not None:
            cur = torch.cuda.current_stream(self.device)
            for st in self.streams:
                st.wait_stream(cur)
        ctx = (lambda i: torch.cuda.stream(self.streams[i])) if self.streams is not None else (lambda i: contextlib.nullcontext())
        n = 0
        for i, (sub, sl) in enumerate(zip(self.subs, self.slices)):
            with ctx(i):
                n = sub.begin_rollout(commands[sl], total_steps, use_cuda_graph, init[sl])
        for _ in range(n):
            for i, sub in enumerate(self.subs):
                with ctx(i):
                    sub.advance_rollout()
        outs = []
        for i, sub in enumerate(self.subs):
            with ctx(i):
                outs.append(sub.finish_rollout())
        if self.streams is not None:
            for st in self.streams:
                torch.cuda.current_stream(self.device).wait_stream(st)
        return {"total_reward": np.concatenate([o["total_reward"] for o in outs]),
                "fim": np.concatenate([o["fim"] for o in outs]), "steps": outs[0]["steps"]}


def expand_commands(sampled: np.ndarray, sampling_idxs=COMMAND_SAMPLING_IDXS, default_command=DEFAULT_COMMAND) -> np.ndarray:
    """sampled[T, len(idxs)] -> full[T, 14] with the defaults elsewhere (:284-293)."""
    full = np.zeros((sampled.shape[0], len(default_command)), dtype=np.float32)
    full[:, :] = np.asarray(default_command, dtype=np.float32)
    for i, idx in enumerate(sampling_idxs):
        full[:, idx] = sampled[:, i]
    return full


def commands_constant(values: np.ndarray, num_steps_per_update: int) -> np.ndarray:
    """values[num_updates, dims] -> [num_updates * steps, dims]: one scalar per update window (:295-314)."""
    return np.repeat(np.asarray(values, dtype=np.float32), num_steps_per_update, axis=0)


def commands_polynomial(coeffs: np.ndarray, ranges: np.ndarray, num_steps_per_update: int) -> np.ndarray:
    """coeffs[num_updates, dims, degree + 1] in [-2, 2]: per window a polynomial in t in [0, 1], squashed by tanh to
    [-1, 1] and mapped linearly onto the range (:316-360; float32 accumulation like the reference)."""
    t = np.linspace(0.0, 1.0, num_steps_per_update, dtype=np.float32)
    out = []
    for u in range(coeffs.shape[0]):
        seg = np.zeros((num_steps_per_update, coeffs.shape[1]), dtype=np.float32)
        for d in range(coeffs.shape[1]):
            poly = np.zeros_like(t)
            for i in range(coeffs.shape[2]):
                poly += coeffs[u, d, i] * (t ** i)
            poly = np.tanh(poly)
            vmin, vmax = ranges[d]
            seg[:, d] = vmin + (poly + 1.0) * 0.5 * (vmax - vmin)
        out.append(seg)
    return np.concatenate(out, axis=0).astype(np.float32)


def _binomial_coefficient(n: int, k: int) -> int:
    """active_sysid.py:402-410 (float product truncated to int, like the reference)."""
    if k > n - k:
        k = n - k
    result = 1
    for i in range(k):
        result *= (n - i) / (i + 1)
    return int(result)


def commands_bezier(points: np.ndarray, ranges: np.ndarray, num_steps_per_update: int) -> np.ndarray:
    """points[num_updates, num_points, dims] control points: per update WINDOW a Bernstein curve over t in [0, 1], every sample
    clipped to the dimension's range (:362-400) -> [num_updates * steps, dims]."""
    points = np.asarray(points, dtype=np.float32)
    n = points.shape[1] - 1
    t = np.linspace(0, 1, num_steps_per_update, dtype=np.float32)
    out = []
    for u in range(points.shape[0]):
        seg = np.zeros((num_steps_per_update, points.shape[2]), dtype=np.float64)
        for j in range(n + 1):
            bern = _binomial_coefficient(n, j) * (t ** j) * ((1 - t) ** (n - j))          # float32 powers, like the reference
            seg += bern.astype(np.float64)[:, None] * points[u, j].astype(np.float64)[None]
        out.append(np.clip(seg, ranges[:, 0][None], ranges[:, 1][None]).astype(np.float32))
    return np.concatenate(out, axis=0).astype(np.float32)


def sample_commands(suggest, mode: str, command_ranges=COMMAND_RANGES, num_command_updates: int = 5,
                    num_steps_per_update: int = 250, sampling_idxs=COMMAND_SAMPLING_IDXS, default_command=DEFAULT_COMMAND,
                    poly_degree: int = 3, num_bezier_points: int = 4) -> np.ndarray:
    """ActiveSysId.sample_commands (active_sysid.py:259-293) for the three sampling modes.  `suggest(name, low, high)` plays
    trial.suggest_float: the parameter names and the ORDER of the calls are the reference's, so a study (or a recorded
    sequence of suggestions) drives both implementations identically.  -> [num_updates * steps, 14]."""
    ranges = np.asarray(command_ranges, dtype=np.float32)
    dims = list(sampling_idxs)
    sub = ranges[dims]
    if mode == "constant":                                    # :295-314 — dimension-major suggestion order
        vals = np.zeros((num_command_updates, len(dims)), dtype=np.float32)
        for d, actual in enumerate(dims):
            for u in range(num_command_updates):
                vals[u, d] = suggest(f"dim_{actual}_update_{u}", float(ranges[actual][0]), float(ranges[actual][1]))
        sampled = commands_constant(vals, num_steps_per_update)
    elif mode == "polynomial":                                # :316-360 — update-major, coefficients in [-2, 2]
        coeffs = np.zeros((num_command_updates, len(dims), poly_degree + 1), dtype=np.float64)
        for u in range(num_command_updates):
            for d, actual in enumerate(dims):
                for i in range(poly_degree + 1):
                    coeffs[u, d, i] = suggest(f"dim_{actual}_update_{u}_coeff_{i}", -2.0, 2.0)
        sampled = commands_polynomial(coeffs, sub, num_steps_per_update)
    elif mode == "bezier":                                    # :362-400 — update-major, control points inside the ranges
        pts = np.zeros((num_command_updates, num_bezier_points, len(dims)), dtype=np.float32)
        for u in range(num_command_updates):
            for d, actual in enumerate(dims):
                for k in range(num_bezier_points):
                    pts[u, k, d] = suggest(f"dim_{actual}_update_{u}_bezier_{k}", float(ranges[actual][0]), float(ranges[actual][1]))
        sampled = commands_bezier(pts, sub, num_steps_per_update)
    else:
        raise ValueError(f"Unknown command_sampling_mode: {mode}")
    return expand_commands(sampled, sampling_idxs, default_command)


def search_space(mode: str, **kw):
    """(names, low, high) of the suggestions one trial makes in `mode`, in call order — the box a vector optimiser
    (CMA-ES here, optuna's CmaEsSampler in the reference) searches."""
    names, lo, hi = [], [], []

    def rec(name, low, high):
        names.append(name); lo.append(low); hi.append(high)
        return 0.5 * (low + high)
    sample_commands(rec, mode, **kw)
    return names, np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)


def commands_from_vector(x: np.ndarray, mode: str, **kw) -> np.ndarray:
    """One trial's parameter vector (in search_space order) -> commands [T, 14]."""
    it = iter(np.asarray(x, dtype=np.float64).tolist())
    return sample_commands(lambda name, low, high: next(it), mode, **kw)


class CmaEs:
    """Minimal (mu/mu_w, lambda)-CMA-ES with box clipping: stand-in for optuna.samplers.CmaEsSampler
    (config/algo/active_sysid.yaml:19-21), same ask/tell shape as the study loop of active_sysid.py:196-217."""

    def __init__(self, lo, hi, seed: int = 0, sigma0: float = 0.3):
        self.lo, self.hi = np.asarray(lo, float), np.asarray(hi, float)
        self.n = self.lo.size
        self.rng = np.random.default_rng(seed)
        self.mean = 0.5 * (self.lo + self.hi)
        self.sigma = sigma0
        self.C = np.eye(self.n)
        self.pc, self.ps = np.zeros(self.n), np.zeros(self.n)
        self.gen = 0
        self.best = (None, np.inf)

    def _scale(self):
        return np.maximum(self.hi - self.lo, 1e-12)

    def ask(self, lam: int) -> np.ndarray:
        A = np.linalg.cholesky(self.C + 1e-12 * np.eye(self.n))
        self._z = self.rng.standard_normal((lam, self.n)) @ A.T
        x = self.mean[None] + self.sigma * self._z * self._scale()[None]
        self._x = np.clip(x, self.lo, self.hi)
        return self._x

    def tell(self, values: np.ndarray):
        lam, n = self._x.shape[0], self.n
        order = np.argsort(values, kind="stable")
        if values[order[0]] < self.best[1]:
            self.best = (self._x[order[0]].copy(), float(values[order[0]]))
        mu = max(1, lam // 2)
        w = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1)); w /= w.sum()
        mueff = 1.0 / (w ** 2).sum()
        cc, cs = (4 + mueff / n) / (n + 4 + 2 * mueff / n), (mueff + 2) / (n + mueff + 5)
        c1 = 2 / ((n + 1.3) ** 2 + mueff)
        cmu = min(1 - c1, 2 * (mueff - 2 + 1 / mueff) / ((n + 2) ** 2 + mueff))
        damps = 1 + 2 * max(0.0, math.sqrt((mueff - 1) / (n + 1)) - 1) + cs
        y = (self._x[order[:mu]] - self.mean[None]) / (self.sigma * self._scale()[None])
        yw = (w[:, None] * y).sum(axis=0)
        self.mean = np.clip(self.mean + self.sigma * self._scale() * yw, self.lo, self.hi)
        Cinv_sqrt = np.linalg.inv(np.linalg.cholesky(self.C + 1e-12 * np.eye(n)))
        self.ps = (1 - cs) * self.ps + math.sqrt(cs * (2 - cs) * mueff) * (Cinv_sqrt @ yw)
        self.gen += 1
        chi = math.sqrt(n) * (1 - 1 / (4 * n) + 1 / (21 * n * n))
        hs = float(np.linalg.norm(self.ps) / math.sqrt(1 - (1 - cs) ** (2 * self.gen)) < (1.4 + 2 / (n + 1)) * chi)
        self.pc = (1 - cc) * self.pc + hs * math.sqrt(cc * (2 - cc) * mueff) * yw
        self.C = ((1 - c1 - cmu) * self.C + c1 * (np.outer(self.pc, self.pc) + (1 - hs) * cc * (2 - cc) * self.C)
                  + cmu * (y.T * w) @ y)
        self.C = 0.5 * (self.C + self.C.T)
        self.sigma *= math.exp((cs / damps) * (np.linalg.norm(self.ps) / chi - 1))
        self.sigma = float(np.clip(self.sigma, 1e-4, 1.0))


def gather_main_results(main_reward: torch.Tensor, fim: Optional[torch.Tensor], world: int, group=None):
    """All-gather the per-trial rewards [M_local] (and Fisher blocks [M_local, P, P]) of every rank's shard of command
    trajectories into rank order (one small collective per iteration; NCCL on CUDA tensors, gloo on CPU)."""
    if world == 1:
        return main_reward, fim
    import torch.distributed as dist

    def gather(x):
        x = x.contiguous()
        out = torch.empty((world,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        if x.is_cuda:
            dist.all_gather_into_tensor(out, x, group=group)
        else:
            dist.all_gather(list(out.unbind(0)), x, group=group)
        return out.reshape((-1,) + tuple(x.shape[1:]))
    return gather(main_reward), (None if fim is None else gather(fim))


def optimize_commands(explorer: ActiveExploration, iterations: int = 5, rollout_length: float = 25.0,
                      horizon_length: float = 5.0, seed: int = 0, total_steps: Optional[int] = None,
                      rank: int = 0, world: int = 1, group=None, mode: str = "constant", poly_degree: int = 3,
                      num_bezier_points: int = 4):
    """The study loop of active_sysid.py:165-242: M trials per iteration, each trial's commands sampled in `mode`
    (`constant` | `polynomial` | `bezier`, config/algo/active_sysid.yaml:44-50) over 3 command dims x (rollout / horizon)
    update windows; objective = -total_reward of the main env.

    Multi-GPU (world > 1): the M = world * explorer.num_main_envs trials of an iteration are sharded contiguously over
    the ranks; every rank runs the same seeded sampler, rolls out its own slice and all-gathers the [M_local] rewards
    and [M_local, P, P] Fisher blocks, so `tell` sees the full population on every rank (no broadcast)."""
    dt = explorer.dt
    kw = dict(num_command_updates=int(rollout_length // horizon_length), num_steps_per_update=int(horizon_length / dt),
              poly_degree=poly_degree, num_bezier_points=num_bezier_points)
    names, lo, hi = search_space(mode, **kw)
    es = CmaEs(lo, hi, seed=seed)
    M_local, P1 = explorer.num_main_envs, explorer.param_dim + 1
    M = M_local * world
    history, best_fim = [], None
    for it in range(iterations):
        x = es.ask(M)                                                     # [M, len(names)], identical on every rank
        mine = x[rank * M_local:(rank + 1) * M_local]
        cmds = np.stack([commands_from_vector(xi, mode, **kw) for xi in mine])
        out = explorer.evaluate_policy(torch.from_numpy(cmds), total_steps=total_steps)
        dev = explorer.device
        local_r = torch.from_numpy(np.ascontiguousarray(out["total_reward"][::P1])).to(dev)
        local_f = torch.from_numpy(out["fim"]).to(dev)
        main_reward, fim = gather_main_results(local_r, local_f, world, group)
        main_reward, fim = main_reward.cpu().numpy(), fim.cpu().numpy()
        prev_best = es.best[1]
        es.tell(-main_reward)
        if es.best[1] < prev_best:
            best_fim = fim[int(np.argmax(main_reward))]
        history.append(float(main_reward.max()))
    best_x, best_v = es.best
    return {"best_commands": commands_from_vector(best_x, mode, **kw), "best_value": -best_v, "history": history,
            "best_fim": best_fim, "param_names": names}
