"""the per-step Python / torch glue the reference pays ON TOP of the physics — the literal
`_compute_torques` + clip (legged_robot_base.py:545, 557), the tanh motor model rebuilt from Python lists every sub-step
(active_sysid_openloop.py:381-400) and the final cost reduction (scripts/eval.py:287-296) — restated in eager torch on the
CPU and timed at the reference's env batches.  No physics: this is the dispatch floor, not a simulator.

    python tools/python_glue_floor.py            # prints candidate-env steps/s of the glue alone at B = 64 and B = 1730
"""
import time

import torch

H, DECIMATION = 5, 4


def glue_step(actions, q, qd, q_default, kp, kd, tlim, hip, thigh, calf, a_list):
    a = torch.clip(actions, -20.0, 20.0)                                         # _pre_physics_step
    for _ in range(DECIMATION):
        tau = kp * (0.25 * a + q_default - q) - kd * qd                          # legged_robot_base.py:545
        tau = torch.clip(tau, -tlim, tlim)                                       # :557
        # act2tau_vec3_tanh: the per-group gains are rebuilt from Python lists every sub-step (:385-393)
        ah = torch.tensor(a_list[0], dtype=torch.float32); at = torch.tensor(a_list[1], dtype=torch.float32)
        ac = torch.tensor(a_list[2], dtype=torch.float32)
        out = tau.clone()
        out[:, hip] = ah[:, None] * torch.tanh(tau[:, hip] / ah[:, None])
        out[:, thigh] = at[:, None] * torch.tanh(tau[:, thigh] / at[:, None])
        out[:, calf] = ac[:, None] * torch.tanh(tau[:, calf] / ac[:, None])
        q = q + 0.0 * out                                                        # stands in for the physics step's refresh
    return q


def run(B, reps=20):
    g = torch.Generator().manual_seed(0)
    q = torch.randn(B, 12, generator=g); qd = torch.randn(B, 12, generator=g); act = torch.randn(B, 12, generator=g)
    qdef = torch.zeros(12); kp = torch.full((12,), 25.0); kd = torch.full((12,), 0.6); tlim = torch.full((12,), 23.7)
    hip, thigh, calf = [0, 3, 6, 9], [1, 4, 7, 10], [2, 5, 8, 11]
    a_list = [[20.0] * B, [20.0] * B, [20.0] * B]
    tgt_p, tgt_q, tgt_j = torch.randn(B, 3, generator=g), torch.randn(B, 4, generator=g), torch.randn(B, 12, generator=g)
    mask = torch.ones(B, dtype=torch.bool)
    t0 = time.perf_counter()
    for _ in range(reps):
        x = q
        for _k in range(H):
            x = glue_step(act, x, qd, qdef, kp, kd, tlim, hip, thigh, calf, a_list)
        e = [torch.norm(x[:, :3] - tgt_p, dim=1), torch.norm(x[:, :4] - tgt_q, dim=1), torch.norm(x - tgt_j, dim=1)]   # eval.py:290-292
        _ = [float((v * mask).sum().item()) for v in e]                           # three .item() host syncs per chunk
    dt = (time.perf_counter() - t0) / reps
    return B * H / dt, dt


if __name__ == "__main__":
    torch.set_num_threads(1)
    for B in (64, 1730):
        rate, dt = run(B)
        print(f"B={B}: glue alone {dt * 1e3:.2f} ms per candidate ({H} control steps x {DECIMATION} sub-steps) "
              f"-> at most {rate:.3e} candidate-env steps/s before any physics")
