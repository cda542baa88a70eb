"""ctypes front-end of the CPU oracle (oracle/spi_oracle.hpp).

TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs import this module.  The product package (spi_active_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libspi_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src_m = max((_HERE / f).stat().st_mtime for f in ("spi_oracle_capi.cpp", "spi_oracle.hpp"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src_m:
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.run(["make", "-C", str(_HERE), "CXX=g++"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(_LIB_PATH))
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, ty=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


def eval_candidates(blob, params, param_ids, seg_init, seg_actions, seg_target, seg_gains=None, seg_mask=None,
                    decimation=4, motor_model=0, flags=0, cost_denominator=0.0, precision=64, n_threads=0,
                    return_per_seg=False):
    """-> cost[C,3] float64, status[C] int32 (, per_seg[C,S,3])"""
    blob = _f32(blob); params = _f32(params); seg_init = _f32(seg_init)
    seg_actions = _f32(seg_actions); seg_target = _f32(seg_target); seg_gains = _f32(seg_gains)
    ids = np.ascontiguousarray(param_ids, dtype=np.int32)
    Cn, P = params.shape
    S, H = seg_actions.shape[0], seg_actions.shape[1]
    mask = None if seg_mask is None else np.ascontiguousarray(seg_mask, dtype=np.uint8)
    cost = np.zeros((Cn, 3), dtype=np.float64)
    per = np.zeros((Cn, S, 3), dtype=np.float64) if return_per_seg else None
    status = np.zeros(Cn, dtype=np.int32)
    rc = lib().spi_oracle_eval_candidates(
        C.c_int(precision), _ptr(blob), C.c_int(blob.size), _ptr(params), C.c_int(Cn), C.c_int(P),
        _ptr(ids, C.c_int), _ptr(seg_init), _ptr(seg_actions), _ptr(seg_target), _ptr(seg_gains),
        _ptr(mask, C.c_ubyte), C.c_int(S), C.c_int(H), C.c_int(decimation), C.c_int(motor_model),
        C.c_uint(flags), C.c_float(cost_denominator), _ptr(cost, C.c_double), _ptr(per, C.c_double),
        _ptr(status, C.c_int), C.c_int(n_threads))
    if rc != 0:
        raise RuntimeError(f"spi_oracle_eval_candidates failed: {rc}")
    return (cost, status, per) if return_per_seg else (cost, status)


def rollout_states(blob, params, param_ids, seg_init, seg_actions, seg_gains=None, decimation=4, motor_model=0,
                   flags=0, precision=64):
    """-> states[C,S,H,37] float64"""
    blob = _f32(blob); params = _f32(params); seg_init = _f32(seg_init)
    seg_actions = _f32(seg_actions); seg_gains = _f32(seg_gains)
    ids = np.ascontiguousarray(param_ids, dtype=np.int32)
    Cn, P = params.shape
    S, H = seg_actions.shape[0], seg_actions.shape[1]
    out = np.zeros((Cn, S, H, 37), dtype=np.float64)
    rc = lib().spi_oracle_rollout_states(
        C.c_int(precision), _ptr(blob), C.c_int(blob.size), _ptr(params), C.c_int(Cn), C.c_int(P),
        _ptr(ids, C.c_int), _ptr(seg_init), _ptr(seg_actions), _ptr(seg_gains), C.c_int(S), C.c_int(H),
        C.c_int(decimation), C.c_int(motor_model), C.c_uint(flags), _ptr(out, C.c_double))
    if rc != 0:
        raise RuntimeError(f"spi_oracle_rollout_states failed: {rc}")
    return out


def sim_step(blob, state, torques, n_steps=1, params=None, param_ids=None, flags=0, return_foot_force=False,
             ext_wrench=None):
    """state[N,37] float64 (copied), torques[N,12] -> new state (, foot_force[N,4,3]).  ext_wrench[N,13,6]: external
    [torque; force] on the 13 moving bodies (base, then hip / thigh / calf per leg) in their link frames, held for the call."""
    blob = _f32(blob)
    state = np.array(state, dtype=np.float64, order="C").reshape(-1, 37)
    torques = np.ascontiguousarray(torques, dtype=np.float64).reshape(-1, 12)
    N = state.shape[0]
    P = 0
    ids = np.zeros(1, dtype=np.int32)
    if params is not None:
        params = _f32(params).reshape(N, -1)
        P = params.shape[1]
        ids = np.ascontiguousarray(param_ids, dtype=np.int32)
    ff = np.zeros((N, 4, 3), dtype=np.float64)
    ext = None if ext_wrench is None else np.ascontiguousarray(ext_wrench, dtype=np.float64).reshape(N, 13, 6)
    rc = lib().spi_oracle_sim_step_ext(_ptr(blob), C.c_int(blob.size), _ptr(params), C.c_int(P), _ptr(ids, C.c_int),
                                       C.c_uint(flags), _ptr(state, C.c_double), _ptr(torques, C.c_double),
                                       C.c_int(N), C.c_int(n_steps), _ptr(ff, C.c_double), _ptr(ext, C.c_double))
    if rc != 0:
        raise RuntimeError(f"spi_oracle_sim_step failed: {rc}")
    return (state, ff) if return_foot_force else state


def env_step(blob, state, actions, params=None, param_ids=None, gains=None, decimation=4, motor_model=0, flags=0):
    """One control step of N independent envs (own parameter row each): state[N,37] float64 (copied) -> new state."""
    blob = _f32(blob)
    state = np.array(state, dtype=np.float64, order="C").reshape(-1, 37)
    actions = _f32(actions).reshape(-1, 12)
    N = state.shape[0]
    P = 0
    ids = np.zeros(1, dtype=np.int32)
    if params is not None:
        params = _f32(params).reshape(N, -1)
        P = params.shape[1]
        ids = np.ascontiguousarray(param_ids, dtype=np.int32)
    gains = _f32(gains)
    rc = lib().spi_oracle_env_step(_ptr(blob), C.c_int(blob.size), _ptr(params), C.c_int(P), _ptr(ids, C.c_int),
                                   _ptr(state, C.c_double), _ptr(actions), _ptr(gains), C.c_int(N), C.c_int(decimation),
                                   C.c_int(motor_model), C.c_uint(flags))
    if rc != 0:
        raise RuntimeError(f"spi_oracle_env_step failed: {rc}")
    return state


def forward_dynamics(blob, state, tau, params=None, param_ids=None, flags=0, with_contact=True, with_gravity=True):
    """-> base spatial accel in base coords [ang3, lin3], qdd[12], foot_force[4,3]"""
    blob = _f32(blob)
    state = np.ascontiguousarray(state, dtype=np.float64).reshape(37)
    tau = np.ascontiguousarray(tau, dtype=np.float64).reshape(12)
    P = 0
    ids = np.zeros(1, dtype=np.int32)
    if params is not None:
        params = _f32(params).reshape(-1)
        P = params.size
        ids = np.ascontiguousarray(param_ids, dtype=np.int32)
    acc = np.zeros(6); qdd = np.zeros(12); foot = np.zeros((4, 3))
    rc = lib().spi_oracle_forward_dynamics(_ptr(blob), C.c_int(blob.size), _ptr(params), C.c_int(P),
                                           _ptr(ids, C.c_int), C.c_uint(flags), _ptr(state, C.c_double),
                                           _ptr(tau, C.c_double), C.c_int(int(with_contact)),
                                           C.c_int(int(with_gravity)), _ptr(acc, C.c_double),
                                           _ptr(qdd, C.c_double), _ptr(foot, C.c_double))
    if rc != 0:
        raise RuntimeError(f"spi_oracle_forward_dynamics failed: {rc}")
    return acc, qdd, foot


def compute_torques(blob, actions, q, qd, gains=None, motor_params=None, motor_model=0, flags=0, precision=64):
    blob = _f32(blob); actions = _f32(actions).reshape(-1, 12); q = _f32(q).reshape(-1, 12)
    qd = _f32(qd).reshape(-1, 12); gains = _f32(gains); motor_params = _f32(motor_params)
    N = actions.shape[0]
    out = np.zeros((N, 12), dtype=np.float64)
    rc = lib().spi_oracle_compute_torques(C.c_int(precision), _ptr(blob), C.c_int(blob.size), _ptr(actions), _ptr(q),
                                          _ptr(qd), _ptr(gains), _ptr(motor_params), C.c_int(N),
                                          C.c_int(motor_model), C.c_uint(flags), _ptr(out, C.c_double))
    if rc != 0:
        raise RuntimeError(f"spi_oracle_compute_torques failed: {rc}")
    return out


def count_flops(blob, params, param_ids, init, actions, target, decimation=4, motor_model=0, flags=0):
    """op counts of one (candidate, segment) rollout -> dict"""
    blob = _f32(blob); params = _f32(params).reshape(-1); init = _f32(init).reshape(37)
    actions = _f32(actions).reshape(-1, 12); target = _f32(target).reshape(19)
    ids = np.ascontiguousarray(param_ids, dtype=np.int32)
    out = np.zeros(6, dtype=np.uint64)
    rc = lib().spi_oracle_count_flops(_ptr(blob), C.c_int(blob.size), _ptr(params), C.c_int(params.size),
                                      _ptr(ids, C.c_int), _ptr(init), _ptr(actions), _ptr(target),
                                      C.c_int(actions.shape[0]), C.c_int(decimation), C.c_int(motor_model),
                                      C.c_uint(flags), _ptr(out, C.c_ulonglong))
    if rc != 0:
        raise RuntimeError(f"spi_oracle_count_flops failed: {rc}")
    names = ("add", "mul", "div", "sqrt", "transcendental", "compare")
    d = {k: int(v) for k, v in zip(names, out)}
    d["total"] = int(out.sum())
    return d


def num_threads() -> int:
    return int(lib().spi_oracle_num_threads())


# ---- numpy restatements of the reference's small torch reductions (pinned by tests/golden) -----------
def fim_reward(root_states, dof_pos, env_origins, num_main, param_dim, delta):
    """ActiveSysId_OpenLoop._reward_fisher_information_matrix
    (spigym/envs/sysid/active_sysid_openloop.py:402-426): envs interleaved [main, aux_1..aux_P] per group,
    J = [(root13_main - root13_aux) / delta | (q_main - q_aux) / delta] with origin-compensated positions,
    reward = trace(J^T J) repeated to all P+1 envs of the group.  Returns (reward[N], JJt[M,P,P])."""
    root = np.asarray(root_states, dtype=np.float64).reshape(num_main, param_dim + 1, 13).copy()
    dof = np.asarray(dof_pos, dtype=np.float64).reshape(num_main, param_dim + 1, 12)
    org = np.asarray(env_origins, dtype=np.float64).reshape(num_main, param_dim + 1, 3)
    root[..., :3] -= org
    x = np.concatenate([root, dof], axis=2)              # [M, P+1, 25]
    J = (x[:, 0:1, :] - x[:, 1:, :]) / float(delta)      # [M, P, 25]
    trace = (J * J).sum(axis=(1, 2))
    JJt = np.einsum("mpd,mqd->mpq", J, J)
    return np.repeat(trace, param_dim + 1), JJt


def fim_states(root_states, dof_pos, env_origins, num_main, param_dim):
    """Pack (root13 origin-compensated, q12) -> states[M, P+1, 25], the input layout of spi_b200_fim_reward."""
    root = np.asarray(root_states, dtype=np.float32).reshape(num_main, param_dim + 1, 13).copy()
    root[..., :3] -= np.asarray(env_origins, dtype=np.float32).reshape(num_main, param_dim + 1, 3)
    dof = np.asarray(dof_pos, dtype=np.float32).reshape(num_main, param_dim + 1, 12)
    return np.concatenate([root, dof], axis=2)


def weighted_cost(cost3, weights=(10.0, 5.0, 1.0)):
    """scripts/mass_landscape.py:162-164 / scripts/mass_opt.py:62-76."""
    c = np.asarray(cost3, dtype=np.float64)
    return c[..., 0] * weights[0] + c[..., 1] * weights[1] + c[..., 2] * weights[2]


def quat_rotate_inverse(q_xyzw, v):
    """spigym/utils/torch_utils.py:83-92 — rotate v by the inverse of the unit quaternion q (x,y,z,w)."""
    q = np.asarray(q_xyzw, dtype=np.float64); v = np.asarray(v, dtype=np.float64)
    w = q[:, 3:4]; u = q[:, :3]
    a = v * (2.0 * w ** 2 - 1.0)
    b = np.cross(u, v) * w * 2.0
    c = u * (u * v).sum(axis=1, keepdims=True) * 2.0
    return a - b + c


# ---- rigid-body state tensor (forward kinematics) -----------------------------------------------------------------
def _quat_mul(a, b):
    ax, ay, az, aw = np.moveaxis(a, -1, 0); bx, by, bz, bw = np.moveaxis(b, -1, 0)
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def _quat_rot(q, v):
    t = 2.0 * np.cross(q[..., :3], v)
    return v + q[..., 3:4] * t + np.cross(q[..., :3], t)


def body_states(blob, state):
    """state[N,37] -> [N,19,13] (pos, quat xyzw, linear velocity of the link origin, angular velocity; world frame) of
    the 19 Isaac Gym bodies in the order of spigym/config/robot/go2/go2.yaml:44 — what
    spigym/simulator/isaacgym/isaacgym.py:541-577 exposes as `_rigid_body_pos/_rot/_vel/_ang_vel`.  fp64 numpy;
    joint origins / axes from the model blob (go2.urdf: hip about x, thigh and calf about y, feet and heads fixed)."""
    b = np.asarray(blob, np.float64)
    s = np.asarray(state, np.float64)
    N = s.shape[0]
    out = np.zeros((N, 19, 13))
    LEG_BODIES, STRIDE, FOOT_OFFSET, FOOT_SPHERE, LUMPS = 46, 14, 214, 12, 26     # include/spi_b200.h

    def child(pp, pq, pv, pw, r, axis, ang=None, rate=None):
        rw = _quat_rot(pq, np.broadcast_to(r, pp.shape))
        cp, cv = pp + rw, pv + np.cross(pw, rw)
        if axis < 0:
            return cp, pq, cv, pw
        jq = np.zeros((N, 4)); jq[:, axis] = np.sin(0.5 * ang); jq[:, 3] = np.cos(0.5 * ang)
        cq = _quat_mul(pq, jq)
        ax = np.zeros((N, 3)); ax[:, axis] = rate
        return cp, cq, cv, pw + _quat_rot(cq, ax)

    base = (s[:, 0:3], s[:, 3:7], s[:, 7:10], s[:, 10:13])
    out[:, 0] = np.concatenate(base, axis=1)
    for h in range(2):
        out[:, 9 + h] = np.concatenate(child(*base, b[LUMPS + 10 * h + 1:LUMPS + 10 * h + 4], -1), axis=1)
    for leg in range(4):
        first = 1 + 4 * leg if leg < 2 else 11 + 4 * (leg - 2)
        cur = base
        for j in range(3):
            r = b[LEG_BODIES + STRIDE * (3 * leg + j) + 10:LEG_BODIES + STRIDE * (3 * leg + j) + 13]
            cur = child(*cur, r, 0 if j == 0 else 1, s[:, 13 + 3 * leg + j], s[:, 25 + 3 * leg + j])
            out[:, first + j] = np.concatenate(cur, axis=1)
        foot = b[FOOT_OFFSET + 3 * leg:FOOT_OFFSET + 3 * leg + 3] - b[FOOT_SPHERE:FOOT_SPHERE + 3]
        out[:, first + 3] = np.concatenate(child(*cur, foot, -1), axis=1)
    return out
