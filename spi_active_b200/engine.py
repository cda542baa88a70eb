"""Python face of the C-ABI (include/spi_b200.h): torch tensors in, torch tensors out, raw pointers
across the boundary.  PyTorch is only the owner of device memory and streams here.

`RolloutEngine.evaluate_candidates` is the fused replacement of the reference's candidate loop
(scripts/mass_landscape.py:123-126: apply_base_mass + evaluate_batch per candidate).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import go2_model as gm
from .dataset import SegmentBatch


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _ids(names_or_ids: Sequence) -> np.ndarray:
    out = []
    for x in names_or_ids:
        if isinstance(x, str):
            if x not in gm.PARAM_IDS:
                raise KeyError(f"unknown parameter name {x!r}; known: {sorted(gm.PARAM_IDS)}")
            out.append(gm.PARAM_IDS[x])
        else:
            out.append(int(x))
    return np.asarray(out, dtype=np.int32)


def motor_model_id(name_or_id) -> int:
    if isinstance(name_or_id, str):
        return gm.MOTOR_MODELS[name_or_id]
    return int(name_or_id)


class TensorCorePolicy:
    """The actor MLP (Linear-ELU x3 + Linear) on the tensor cores with fp32-grade accuracy (spi_b200_policy_*,
    fp16 pairs: 3 x FP16).  `weights[l]` is the torch Linear weight [out, in]; the reference's 900-512-256-128-12 actor qualifies
    (h1, h2 multiples of 128, h3 = 128, <= 16 outputs) — anything else raises and the caller keeps its cuBLAS path."""

    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], device: torch.device):
        self.lib = _lib.lib()
        if len(weights) != 4 or len(biases) != 4:
            raise _lib.SpiB200Error("tensor-core policy needs exactly 4 Linear layers")
        ws = [np.ascontiguousarray(w.detach().cpu().numpy(), dtype=np.float32) for w in weights]
        bs = [np.ascontiguousarray(b.detach().cpu().numpy(), dtype=np.float32) for b in biases]
        dims = np.asarray([ws[0].shape[1]] + [w.shape[0] for w in ws], dtype=np.int32)
        for l in range(4):
            if ws[l].shape != (dims[l + 1], dims[l]) or bs[l].shape != (dims[l + 1],):
                raise _lib.SpiB200Error(f"layer {l}: inconsistent shapes {ws[l].shape} / {bs[l].shape}")
        self.device = torch.device(device)
        self.dims = dims.tolist()
        fp = C.POINTER(C.c_float)
        wp = (fp * 4)(*[w.ctypes.data_as(fp) for w in ws])
        bp = (fp * 4)(*[b.ctypes.data_as(fp) for b in bs])
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_policy_create(dims.ctypes.data_as(C.POINTER(C.c_int)), wp, bp, C.byref(self._handle))
        _lib.check(rc, "spi_b200_policy_create")

    def close(self):
        if getattr(self, "_handle", None) and self._handle.value:
            self.lib.spi_b200_policy_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def input_layout(self, M: int):
        """-> (rows, stride): the pre-split input buffers for M envs hold rows * stride floats each, in an opaque tiled
        order (csrc/tiled_layout.cuh) — write them with split_input or spi_b200_active_post_step only."""
        rows, stride = C.c_int(), C.c_int()
        _lib.check(self.lib.spi_b200_policy_input_layout(self._handle, int(M), C.byref(rows), C.byref(stride)),
                   "spi_b200_policy_input_layout")
        return rows.value, stride.value

    def alloc_input(self, M: int):
        rows, stride = self.input_layout(M)
        return (torch.zeros(rows, stride, device=self.device, dtype=torch.float16),
                torch.zeros(rows, stride, device=self.device, dtype=torch.float16))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def split_input(self, x: torch.Tensor, x_hi: torch.Tensor, x_lo: torch.Tensor):
        assert x.is_cuda and x.is_contiguous() and x.dtype == torch.float32 and x.shape[1] == self.dims[0]
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_policy_split_input(self._handle, _ptr(x), int(x.shape[0]), _ptr(x_hi), _ptr(x_lo),
                                                      self._stream())
        _lib.check(rc, "spi_b200_policy_split_input")

    def unsplit_input(self, x_hi: torch.Tensor, x_lo: torch.Tensor, M: int) -> torch.Tensor:
        """The plain [M, in] matrix a pair of split buffers holds (x_hi + x_lo, exact)."""
        x = torch.empty((M, self.dims[0]), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_policy_unsplit_input(self._handle, _ptr(x_hi), _ptr(x_lo), int(M), _ptr(x), self._stream())
        _lib.check(rc, "spi_b200_policy_unsplit_input")
        return x

    def forward_split(self, x_hi: torch.Tensor, x_lo: torch.Tensor, M: int, out: Optional[torch.Tensor] = None):
        rows, stride = self.input_layout(M)
        assert tuple(x_hi.shape) == (rows, stride) == tuple(x_lo.shape) and x_hi.is_contiguous() and x_lo.is_contiguous()
        if out is None:
            out = torch.empty((M, self.dims[4]), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_policy_forward(self._handle, _ptr(x_hi), _ptr(x_lo), int(M), _ptr(out), self._stream())
        _lib.check(rc, "spi_b200_policy_forward")
        return out

    def enable_ring(self, col_map: np.ndarray):
        """col_map[n_rot, in] (int): ring element k multiplies the weight of original input column col_map[r, k] when the
        head is at position r (-1 = unused) — spi_b200_policy_enable_ring."""
        cm = np.ascontiguousarray(col_map, dtype=np.int32)
        assert cm.ndim == 2 and cm.shape[1] == self.dims[0]
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_policy_enable_ring(self._handle, cm.ctypes.data_as(C.POINTER(C.c_int)), int(cm.shape[0]))
        _lib.check(rc, "spi_b200_policy_enable_ring")
        self.n_rot = int(cm.shape[0])

    def forward_ring(self, x_hi: torch.Tensor, x_lo: torch.Tensor, M: int, rot: torch.Tensor, out: Optional[torch.Tensor] = None):
        """forward_split on a ring-ordered input; rot = DEVICE int32 tensor whose element 0 is the head position."""
        rows, stride = self.input_layout(M)
        assert tuple(x_hi.shape) == (rows, stride) == tuple(x_lo.shape) and x_hi.is_contiguous() and x_lo.is_contiguous()
        assert rot.is_cuda and rot.dtype == torch.int32
        if out is None:
            out = torch.empty((M, self.dims[4]), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_policy_forward_ring(self._handle, _ptr(x_hi), _ptr(x_lo), int(M), _ptr(rot), _ptr(out),
                                                       self._stream())
        _lib.check(rc, "spi_b200_policy_forward_ring")
        return out

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(self.device, torch.float32).contiguous()
        x_hi, x_lo = self.alloc_input(x.shape[0])
        self.split_input(x, x_hi, x_lo)
        return self.forward_split(x_hi, x_lo, x.shape[0])


class RolloutEngine:
    """Device-resident Go2 model + workspaces; one instance per (process, GPU, stream)."""

    def __init__(self, model: Optional[gm.Go2Model] = None, device: Optional[torch.device] = None):
        self.lib = _lib.lib()  # raises if libspi_b200.so is missing: no CPU fallback
        if not torch.cuda.is_available():
            raise _lib.SpiB200Error("RolloutEngine needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.index is None:
            self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        self._status_buf = None
        self.model = model or gm.go2_nominal()
        self.blob = gm.build_model_blob(self.model)
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_model_create(self.blob.ctypes.data_as(C.POINTER(C.c_float)), int(self.blob.size),
                                                C.byref(self._handle))
        _lib.check(rc, "spi_b200_model_create")
        self.kernel = "auto"

    def close(self):
        if getattr(self, "_handle", None) and self._handle.value:
            self.lib.spi_b200_model_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers --------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f32(self, t, shape=None) -> Optional[torch.Tensor]:
        if t is None:
            return None
        t = torch.as_tensor(t, device=self.device, dtype=torch.float32).contiguous()
        if shape is not None:
            t = t.reshape(shape)
        return t

    # ---- fused hot path -------------------------------------------------------------------
    def evaluate_candidates(self, params, param_names, segs: SegmentBatch, decimation: Optional[int] = None,
                            motor_model="none", flags: int = 0, cost_denominator: Optional[float] = None,
                            return_per_seg: bool = False, return_status: bool = False, out: torch.Tensor = None):
        """params[C,P] -> cost[C,3] (mean base_pos, base_quat, joint_pos L2 errors)."""
        params = self._f32(params)
        if params.dim() == 1:
            params = params[:, None]
        Cn, P = params.shape
        ids = _ids(param_names)
        assert ids.size == P, "one name/id per parameter column"
        S, H = segs.num_segments, segs.horizon
        # raw pointers cross the C-ABI below: a wrong device / dtype / stride must be an error here, not an illegal access there
        for name, t, dt, shape in (("seg_init", segs.seg_init, torch.float32, (S, 37)),
                                   ("seg_actions", segs.seg_actions, torch.float32, (S, H, 12)),
                                   ("seg_target", segs.seg_target, torch.float32, (S, 19)),
                                   ("seg_gains", segs.seg_gains, torch.float32, (S, 24)),
                                   ("seg_mask", segs.seg_mask, torch.uint8, (S,)),
                                   ("out", out, torch.float32, (Cn, 3))):
            if t is None:
                if name in ("seg_gains", "seg_mask", "out"):
                    continue
                raise ValueError(f"{name} is None")
            if not (t.is_cuda and t.device == self.device and t.dtype == dt and t.is_contiguous() and tuple(t.shape) == shape):
                raise ValueError(f"{name}: need a contiguous {dt} tensor of shape {shape} on {self.device}, got "
                                 f"{t.dtype} {tuple(t.shape)} on {t.device} (contiguous={t.is_contiguous()})")
        cost = out if out is not None else torch.empty((Cn, 3), device=self.device, dtype=torch.float32)
        per = torch.empty((Cn, S, 3), device=self.device, dtype=torch.float32) if return_per_seg else None
        if return_status:
            status = torch.empty((Cn,), device=self.device, dtype=torch.int32)     # handed to the caller: its own buffer
        else:                                                                      # scratch: grown once, then reused
            if self._status_buf is None or self._status_buf.shape[0] < Cn:
                self._status_buf = torch.empty((Cn,), device=self.device, dtype=torch.int32)
            status = self._status_buf[:Cn]
        denom = segs.cost_denominator if cost_denominator is None else cost_denominator
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_eval_candidates(
                self._handle, _ptr(params), Cn, P, ids.ctypes.data_as(C.POINTER(C.c_int)), _ptr(segs.seg_init),
                _ptr(segs.seg_actions), _ptr(segs.seg_target), _ptr(segs.seg_gains), _ptr(segs.seg_mask), S, H,
                int(decimation or self.model.control_decimation), motor_model_id(motor_model), int(flags),
                float(denom), _ptr(cost), _ptr(per), _ptr(status), self._stream())
        _lib.check(rc, "spi_b200_eval_candidates")
        res = (cost,)
        if return_per_seg:
            res += (per,)
        if return_status:
            res += (status,)
        return res if len(res) > 1 else cost

    def evaluate_candidates_host(self, params: np.ndarray, param_names, seg_init: np.ndarray, seg_actions: np.ndarray,
                                 seg_target: np.ndarray, seg_gains: Optional[np.ndarray] = None,
                                 seg_mask: Optional[np.ndarray] = None, decimation: Optional[int] = None,
                                 motor_model="none", flags: int = 0, cost_denominator: float = 0.0):
        """numpy in / numpy out through spi_b200_eval_candidates_host (H2D + D2H inside the call)."""
        params = np.ascontiguousarray(params, dtype=np.float32).reshape(len(params), -1)
        seg_init = np.ascontiguousarray(seg_init, dtype=np.float32)
        seg_actions = np.ascontiguousarray(seg_actions, dtype=np.float32)
        seg_target = np.ascontiguousarray(seg_target, dtype=np.float32)
        seg_gains = None if seg_gains is None else np.ascontiguousarray(seg_gains, dtype=np.float32)
        seg_mask = None if seg_mask is None else np.ascontiguousarray(seg_mask, dtype=np.uint8)
        Cn, P = params.shape
        ids = _ids(param_names)
        S, H = seg_actions.shape[0], seg_actions.shape[1]
        cost = np.empty((Cn, 3), dtype=np.float32)
        status = np.empty((Cn,), dtype=np.int32)
        fp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_eval_candidates_host(
                self._handle, fp(params), Cn, P, ids.ctypes.data_as(C.POINTER(C.c_int)), fp(seg_init), fp(seg_actions),
                fp(seg_target), fp(seg_gains),
                None if seg_mask is None else seg_mask.ctypes.data_as(C.POINTER(C.c_ubyte)), S, H,
                int(decimation or self.model.control_decimation), motor_model_id(motor_model), int(flags),
                float(cost_denominator), fp(cost), status.ctypes.data_as(C.POINTER(C.c_int)), self._stream())
        _lib.check(rc, "spi_b200_eval_candidates_host")
        return cost, status

    def rollout_states(self, params, param_names, seg_init, seg_actions, seg_gains=None, decimation=None,
                       motor_model="none", flags: int = 0) -> torch.Tensor:
        """-> states[C,S,H,37] after each control step (parity hook, dataset recorder)."""
        params = self._f32(params)
        if params.dim() == 1:
            params = params[:, None]
        Cn, P = params.shape
        ids = _ids(param_names)
        seg_init = self._f32(seg_init).reshape(-1, gm.STATE_DIM)
        S = seg_init.shape[0]
        seg_actions = self._f32(seg_actions).reshape(S, -1, 12)
        H = seg_actions.shape[1]
        seg_gains = self._f32(seg_gains)
        out = torch.empty((Cn, S, H, gm.STATE_DIM), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_rollout_states(
                self._handle, _ptr(params), Cn, P, ids.ctypes.data_as(C.POINTER(C.c_int)), _ptr(seg_init),
                _ptr(seg_actions), _ptr(seg_gains), S, H, int(decimation or self.model.control_decimation),
                motor_model_id(motor_model), int(flags), _ptr(out), self._stream())
        _lib.check(rc, "spi_b200_rollout_states")
        return out

    # ---- stepwise boundary ------------------------------------------------------------------
    def sim_step(self, state: torch.Tensor, torques: torch.Tensor, n_steps: int = 1, params=None, param_names=(),
                 flags: int = 0, foot_force: Optional[torch.Tensor] = None,
                 ext_wrench: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Advance state[N,37] IN PLACE by n_steps physics steps under constant torques[N,12].  ext_wrench[N,13,6]: external
        [torque; force] on the 13 moving bodies in their link frames (spi_b200_sim_step_ext)."""
        assert state.is_cuda and state.dtype == torch.float32 and state.is_contiguous()
        torques = self._f32(torques)
        N = state.shape[0]
        P = 0
        ids = np.zeros(1, dtype=np.int32)
        if params is not None:
            params = self._f32(params).reshape(N, -1)
            P = params.shape[1]
            ids = _ids(param_names)
        with torch.cuda.device(self.device):
            if ext_wrench is not None:
                ext_wrench = self._f32(ext_wrench).reshape(N, 13, 6).contiguous()
            rc = self.lib.spi_b200_sim_step_ext(self._handle, _ptr(params), P, ids.ctypes.data_as(C.POINTER(C.c_int)),
                                                int(flags), _ptr(state), _ptr(torques), _ptr(ext_wrench), N, int(n_steps),
                                                _ptr(foot_force), self._stream())
        _lib.check(rc, "spi_b200_sim_step_ext")
        return state

    def body_states(self, state: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """state[N,37] -> [N,19,13] rigid-body states of the 19 Isaac Gym bodies (forward kinematics, world frame)."""
        assert state.is_cuda and state.dtype == torch.float32 and state.is_contiguous()
        N = state.shape[0]
        if out is None:
            out = torch.empty((N, 19, 13), device=self.device, dtype=torch.float32)
        assert out.is_contiguous() and tuple(out.shape) == (N, 19, 13)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_body_states(self._handle, _ptr(state), N, _ptr(out), self._stream())
        _lib.check(rc, "spi_b200_body_states")
        return out

    def env_step(self, state: torch.Tensor, actions: torch.Tensor, params=None, param_names=(), gains=None,
                 decimation: Optional[int] = None, motor_model="none", flags: int = 0,
                 zero_action_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One control step of N independent envs IN PLACE: state[N,37], actions[N,12], per-env params[N,P];
        zero_action_mask[N] (bool / uint8): envs whose action is replaced by 0."""
        assert state.is_cuda and state.dtype == torch.float32 and state.is_contiguous()
        actions = self._f32(actions).reshape(-1, 12)
        N = state.shape[0]
        P = 0
        ids = np.zeros(1, dtype=np.int32)
        if params is not None:
            params = self._f32(params).reshape(N, -1)
            P = params.shape[1]
            ids = _ids(param_names)
        gains = self._f32(gains)
        if zero_action_mask is not None:
            assert zero_action_mask.is_cuda and zero_action_mask.element_size() == 1 and zero_action_mask.numel() == N
            zero_action_mask = zero_action_mask.contiguous()
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_env_step(self._handle, _ptr(params), P, ids.ctypes.data_as(C.POINTER(C.c_int)),
                                            _ptr(state), _ptr(actions), _ptr(zero_action_mask), _ptr(gains), N,
                                            int(decimation or self.model.control_decimation),
                                            motor_model_id(motor_model), int(flags), self._stream())
        _lib.check(rc, "spi_b200_env_step")
        return state

    def compute_torques(self, actions, q, qd, gains=None, motor_params=None, motor_model="none", flags: int = 0):
        actions = self._f32(actions).reshape(-1, 12)
        q = self._f32(q).reshape(-1, 12)
        qd = self._f32(qd).reshape(-1, 12)
        gains = self._f32(gains)
        motor_params = self._f32(motor_params)
        N = actions.shape[0]
        out = torch.empty((N, 12), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_compute_torques(self._handle, _ptr(actions), _ptr(q), _ptr(qd), _ptr(gains),
                                                   _ptr(motor_params), N, motor_model_id(motor_model), int(flags),
                                                   _ptr(out), self._stream())
        _lib.check(rc, "spi_b200_compute_torques")
        return out

    # ---- Fisher information ---------------------------------------------------------------------
    def fim_reward(self, states: torch.Tensor, delta: float, out_JtJ: Optional[torch.Tensor] = None,
                   out_trace: Optional[torch.Tensor] = None, accumulate: bool = False):
        """states[M,P+1,25] (main + P aux; root13 + q12) -> (JJt[M,P,P], trace[M])."""
        states = self._f32(states)
        Mn, P1, D = states.shape
        assert D == 25 and P1 >= 2
        P = P1 - 1
        if out_JtJ is None:
            out_JtJ = torch.zeros((Mn, P, P), device=self.device, dtype=torch.float32)
        if out_trace is None:
            out_trace = torch.zeros((Mn,), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_fim_reward(self._handle, _ptr(states), Mn, P, float(delta), int(accumulate),
                                              _ptr(out_JtJ), _ptr(out_trace), self._stream())
        _lib.check(rc, "spi_b200_fim_reward")
        return out_JtJ, out_trace

    def active_post_step(self, state, raw_actions, done, main_commands, commands, actions, gait, clock, history, obs,
                         hist_index, fim_hist, fim_live, dead_steps, schedule, counter, ctrl, dt: float,
                         action_clip: float, clip_obs: float, grav_xy, q_default, obs_hi=None, obs_lo=None,
                         ring_slots: int = 0, fim_jtj=None, fim_trace=None, fim_delta: float = 0.0):
        """The fused post-physics step of the active-exploration rollout (spi_b200_active_post_step); every tensor is
        updated in place.  state[N,37], main_commands[M,T,14], N = M * P1.  ring_slots = 15: obs_hi / obs_lo are the ring
        of frames the actor reads (history / obs / hist_index unused, may be None).  fim_jtj [M,P,P] / fim_trace [M]: fused
        Fisher accumulation inside the kernel (fim_hist / fim_live then None)."""
        if fim_jtj is not None:
            assert fim_jtj.is_cuda and fim_jtj.is_contiguous() and fim_jtj.dtype == torch.float32 and fim_delta != 0.0
            assert fim_trace is None or (fim_trace.is_cuda and fim_trace.is_contiguous() and fim_trace.dtype == torch.float32)
        Mn, T = int(main_commands.shape[0]), int(main_commands.shape[1])
        N = int(state.shape[0])
        P1 = N // Mn
        assert N == Mn * P1 and main_commands.shape[2] == 14 and done.element_size() == 1
        for t in (state, raw_actions, done, main_commands, commands, actions, gait, clock, history, obs, hist_index,
                  schedule, counter, ctrl):
            assert t is None or (t.is_cuda and t.is_contiguous())
        assert (hist_index is None or hist_index.dtype == torch.int32) and schedule.dtype == torch.int32 and counter.dtype == torch.int32
        qd = np.ascontiguousarray(q_default, dtype=np.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_active_post_step(
                self._handle, _ptr(state), _ptr(raw_actions), _ptr(done), _ptr(main_commands), T, _ptr(commands),
                _ptr(actions), _ptr(gait), _ptr(clock), _ptr(history), _ptr(obs), _ptr(obs_hi), _ptr(obs_lo),
                0 if obs_hi is None else int(obs_hi.shape[1]), int(ring_slots), _ptr(hist_index), _ptr(fim_hist),
                _ptr(fim_live), _ptr(dead_steps), _ptr(fim_jtj), _ptr(fim_trace), float(fim_delta), _ptr(schedule),
                int(schedule.shape[0]), _ptr(counter), _ptr(ctrl), Mn, P1, float(dt),
                float(action_clip), float(clip_obs), float(grav_xy[0]), float(grav_xy[1]),
                qd.ctypes.data_as(C.POINTER(C.c_float)), self._stream())
        _lib.check(rc, "spi_b200_active_post_step")

    def tensor_core_policy(self, weights, biases) -> "TensorCorePolicy":
        """The actor MLP as a tensor-core operator on this engine's device (raises SpiB200Error for unsupported shapes)."""
        return TensorCorePolicy(weights, biases, self.device)

    def fim_contract(self, hist: torch.Tensor, delta: float, live: Optional[torch.Tensor] = None,
                     out_JtJ: Optional[torch.Tensor] = None, out_trace: Optional[torch.Tensor] = None,
                     accumulate: bool = False):
        """hist[T,M,P+1,25] (the per-step `states` of fim_reward stacked over T control steps), live[T,M] (bool/u8) ->
        (sum_t live * J_t J_t^T [M,P,P], its trace [M]) on the tensor cores (spi_b200_fim_contract, 3xTF32)."""
        hist = self._f32(hist)
        T, Mn, P1, D = hist.shape
        assert D == 25 and P1 >= 2
        P = P1 - 1
        if live is not None:
            live = torch.as_tensor(live, device=self.device).to(torch.uint8).contiguous()
            assert tuple(live.shape) == (T, Mn)
        if out_JtJ is None:
            out_JtJ = torch.zeros((Mn, P, P), device=self.device, dtype=torch.float32)
        if out_trace is None:
            out_trace = torch.zeros((Mn,), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_fim_contract(self._handle, _ptr(hist), _ptr(live), T, Mn, P, float(delta),
                                                int(accumulate), _ptr(out_JtJ), _ptr(out_trace), self._stream())
        _lib.check(rc, "spi_b200_fim_contract")
        return out_JtJ, out_trace

    # ---- optimiser pieces -------------------------------------------------------------------------
    def weighted_cost(self, cost3: torch.Tensor, w=(10.0, 5.0, 1.0), out: Optional[torch.Tensor] = None):
        cost3 = self._f32(cost3)
        Cn = cost3.shape[0]
        out = out if out is not None else torch.empty((Cn,), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_weighted_cost(self._handle, _ptr(cost3), Cn, float(w[0]), float(w[1]), float(w[2]),
                                                 _ptr(out), self._stream())
        _lib.check(rc, "spi_b200_weighted_cost")
        return out

    def cem_sample(self, mean, std, lo, hi, C_local: int, c0: int, seed: int, iteration: int,
                   out: Optional[torch.Tensor] = None):
        mean = self._f32(mean); std = self._f32(std); lo = self._f32(lo); hi = self._f32(hi)
        P = mean.numel()
        out = out if out is not None else torch.empty((C_local, P), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_cem_sample(self._handle, _ptr(mean), _ptr(std), _ptr(lo), _ptr(hi), int(C_local), P,
                                              int(c0), C.c_ulonglong(seed), int(iteration), _ptr(out), self._stream())
        _lib.check(rc, "spi_b200_cem_sample")
        return out

    def cem_refit(self, params, cost, n_elite: int, alpha: float, mean, std, std_floor=None, out_best=None):
        """mean/std updated IN PLACE; returns out_best[P+1] (best params, best cost)."""
        params = self._f32(params); cost = self._f32(cost)
        Cn, P = params.shape
        assert mean.is_cuda and std.is_cuda and mean.dtype == torch.float32 and std.dtype == torch.float32
        std_floor = self._f32(std_floor)
        out_best = out_best if out_best is not None else torch.empty((P + 1,), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_cem_refit(self._handle, _ptr(params), _ptr(cost), Cn, P, int(n_elite), float(alpha),
                                             _ptr(std_floor), _ptr(mean), _ptr(std), _ptr(out_best), self._stream())
        _lib.check(rc, "spi_b200_cem_refit")
        return out_best

    # ---- measurement ----------------------------------------------------------------------------------
    def fp32_peak(self, iters: int = 4096):
        tf, ms = C.c_float(), C.c_float()
        with torch.cuda.device(self.device):
            rc = self.lib.spi_b200_fp32_peak(int(iters), C.byref(tf), C.byref(ms), self._stream())
        _lib.check(rc, "spi_b200_fp32_peak")
        return float(tf.value), float(ms.value)

    def set_kernel(self, kernel: str = "auto") -> None:
        """'auto' | 'lane' (generic leg-per-lane kernel) | 'ws' (warp-specialised Go2-family fast path) | 'ws-padded' (the fast
        path with every candidate padded to whole CTAs instead of the dense packing: same bits, for comparison)."""
        kid = {"auto": 0, "lane": 1, "ws": 2, "ws-padded": 3}[kernel]
        _lib.check(self.lib.spi_b200_model_set_kernel(self._handle, kid), "spi_b200_model_set_kernel")
        self.kernel = kernel

    def timing_enable(self, on: bool = True) -> None:
        """Bracket every rollout-kernel launch with CUDA events on its stream (roofline instrumentation)."""
        _lib.check(self.lib.spi_b200_timing_enable(self._handle, int(bool(on))), "spi_b200_timing_enable")

    def timing_read(self, reset: bool = True):
        """-> (summed rollout-kernel ms, launches) since the last reset; synchronises on the events."""
        ms, n = C.c_double(), C.c_longlong()
        _lib.check(self.lib.spi_b200_timing_read(self._handle, C.byref(ms), C.byref(n), int(reset)),
                   "spi_b200_timing_read")
        return float(ms.value), int(n.value)

    def launch_count(self) -> int:
        return int(self.lib.spi_b200_launch_count())
