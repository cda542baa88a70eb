"""Active-exploration path (BASELINE config 5).

* observation / gait-clock / history layout against golden vectors produced by the reference's own code
  (tests/golden/active_obs.npz, see make_golden.py);
* the closed-loop evaluate_policy loop on CPU with the oracle as physics (host logic, FIM reward, k-sync,
  termination, command playback);
* on the GPU: spi_b200_env_step against the oracle, and the whole loop (CUDA graph) against the oracle-backed loop.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from spi_active_b200 import active as act
from spi_active_b200 import go2_model as gm

GOLD = Path(__file__).resolve().parent / "golden"


class OracleActiveBackend:
    """env_step / fim_reward with the same signatures as RolloutEngine, on the CPU oracle (test infrastructure)."""
    device = torch.device("cpu")

    def __init__(self, blob, model):
        self.blob, self.model = blob, model

    def env_step(self, state, actions, params=None, param_names=(), gains=None, decimation=None, motor_model="none", flags=0):
        from oracle import oracle as orc
        new = orc.env_step(self.blob, state.numpy().astype(np.float64), actions.numpy(),
                           None if params is None else params.numpy(), [gm.PARAM_IDS[n] for n in param_names], gains,
                           decimation or 4, gm.MOTOR_MODELS[motor_model], flags)
        state.copy_(torch.from_numpy(new.astype(np.float32)))
        return state

    def fim_reward(self, states, delta):
        s = states.numpy().astype(np.float64)
        J = (s[:, 0:1, :] - s[:, 1:, :]) / float(delta)
        return (torch.from_numpy(np.einsum("mpd,mqd->mpq", J, J).astype(np.float32)),
                torch.from_numpy((J * J).sum(axis=(1, 2)).astype(np.float32)))


def test_observation_layout_matches_reference():
    g = np.load(GOLD / "active_obs.npz")
    T, N = g["states"].shape[:2]
    assert act.ACTOR_OBS_DIM == g["actor_obs"].shape[2] == 900 and act.FRAME_DIM == 60
    gait = torch.zeros(N)
    hist = torch.zeros(N, act.HISTORY_LEN, act.FRAME_DIM)
    idx = act._history_gather_index()
    qdef = torch.tensor(gm.go2_nominal().q_default, dtype=torch.float32)
    for t in range(T):
        gait, clock = act.step_contact_targets(gait, torch.from_numpy(g["commands_clock"][t]), 0.02)
        np.testing.assert_allclose(gait.numpy(), g["gait"][t], atol=1e-6)
        np.testing.assert_allclose(clock.numpy(), g["clock"][t], atol=2e-5)
        frame = act.build_frame(torch.from_numpy(g["states"][t]), torch.from_numpy(g["actions"][t]),
                                torch.from_numpy(g["commands_obs"][t]), clock, gait, qdef)
        obs = torch.cat([frame, hist.reshape(N, -1)[:, idx]], dim=1).clip(-100.0, 100.0)
        np.testing.assert_allclose(obs.numpy(), g["actor_obs"][t], atol=3e-5, rtol=1e-6, err_msg=f"step {t}")
        hist = torch.cat([frame[:, None], hist[:, :-1]], dim=1)
    assert np.abs(g["actor_obs"][5]).max() == 100.0        # the clip was exercised


def test_param_table_and_group_layout(blob, nominal_model):
    cfg = act.ActiveConfig(exploration_params=["mass", "comx", "motor_model_calf_a"])
    ex = act.ActiveExploration(OracleActiveBackend(blob, nominal_model), act.PolicyMLP.random("cpu"), 3, cfg)
    assert ex.num_envs == 12 and ex.params.shape == (12, 10)
    tab = ex.params.numpy().reshape(3, 4, 10)
    base = np.array([act.DEFAULT_PARAM[n] for n in ex.param_names], np.float32)
    np.testing.assert_allclose(tab[:, 0], np.tile(base, (3, 1)))                       # main envs: defaults
    for i, col in enumerate([0, 1, 9]):                                                # aux i: +0.1 ABSOLUTE on parameter i (D10)
        d = tab[:, i + 1] - base
        assert np.allclose(d[:, col], 0.1, atol=1e-6) and np.allclose(np.delete(d, col, axis=1), 0)
    assert ex.main_idx.flatten().tolist() == [0, 4, 8] and ex.aux_idx[1].tolist() == [5, 6, 7]
    assert ex.total_steps == 1250


def _commands(M, T, seed=0):
    rng = np.random.default_rng(seed)
    r = np.asarray(act.COMMAND_RANGES)
    vals = rng.uniform(r[act.COMMAND_SAMPLING_IDXS, 0], r[act.COMMAND_SAMPLING_IDXS, 1], (M, 1, 3)).astype(np.float32)
    return torch.from_numpy(np.stack([act.expand_commands(np.repeat(v, T, axis=0)) for v in vals]))


def test_evaluate_policy_on_oracle_backend(blob, nominal_model):
    cfg = act.ActiveConfig(exploration_params=["mass", "comx"], ksync_steps=5, seed=3)
    be = OracleActiveBackend(blob, nominal_model)
    ex = act.ActiveExploration(be, act.PolicyMLP.random("cpu", seed=1), 4, cfg)
    out = ex.evaluate_policy(_commands(4, 40), total_steps=30, use_cuda_graph=False)
    r = out["total_reward"].reshape(4, 3)
    assert out["steps"] == 29 and np.isfinite(r).all() and (r > 0).all()
    assert np.allclose(r, r[:, :1])                      # the reward is repeated over the group (:424)
    np.testing.assert_allclose(np.trace(out["fim"], axis1=1, axis2=2), r[:, 0], rtol=1e-4)
    # with a sync at EVERY step the aux envs never drift for more than one step: smaller information
    ex1 = act.ActiveExploration(be, act.PolicyMLP.random("cpu", seed=1), 4, act.ActiveConfig(
        exploration_params=["mass", "comx"], ksync_steps=1, seed=3))
    r1 = ex1.evaluate_policy(_commands(4, 40), total_steps=30, use_cuda_graph=False)["total_reward"].reshape(4, 3)
    assert (r1[:, 0] < r[:, 0]).all()
    # determinism
    r2 = ex.evaluate_policy(_commands(4, 40), total_steps=30, use_cuda_graph=False)["total_reward"]
    np.testing.assert_array_equal(r2, out["total_reward"])


def test_terminated_groups_score_zero(blob, nominal_model):
    be = OracleActiveBackend(blob, nominal_model)
    ex = act.ActiveExploration(be, act.PolicyMLP.random("cpu", seed=1), 2, act.ActiveConfig(seed=0, randomize_reset=False))
    ex.reset_all(_commands(2, 10))
    ex.state[1, 3:7] = torch.tensor([0.70710678, 0.0, 0.0, 0.70710678])   # aux env of group 0 rolled by 90 degrees
    ex.state[1, 2] = 1.0
    ex._advance_inputs()
    ex._policy_step()
    assert ex.done.tolist() == [True, True, False, False]          # group-wise OR (:259-272)
    assert ex.step_reward[:2].abs().sum() == 0 and ex.step_reward[2] > 0


def test_command_samplers():
    c = act.commands_constant(np.array([[1.0, 2.0], [3.0, 4.0]]), 3)
    assert c.shape == (6, 2) and c[:3, 0].tolist() == [1, 1, 1] and c[3:, 1].tolist() == [4, 4, 4]
    full = act.expand_commands(np.ones((5, 3), np.float32) * 0.7)
    assert full.shape == (5, 14) and np.allclose(full[:, [0, 2, 5]], 0.7) and np.allclose(full[:, 4], 3.0)
    rng = np.random.default_rng(0)
    r = np.asarray(act.COMMAND_RANGES, np.float32)[act.COMMAND_SAMPLING_IDXS]
    p = act.commands_polynomial(rng.uniform(-2, 2, (5, 3, 4)), r, 250)
    assert p.shape == (1250, 3) and (p >= r[:, 0] - 1e-6).all() and (p <= r[:, 1] + 1e-6).all()
    b = act.commands_bezier(np.array([[0.0], [1.0], [1.0], [0.0]]), 101)
    assert abs(b[0, 0]) < 1e-7 and abs(b[-1, 0]) < 1e-7 and abs(b[50, 0] - 0.75) < 1e-6


def test_cma_es_minimises():
    es = act.CmaEs([-1.0] * 6, [1.0] * 6, seed=0)
    target = np.array([0.3, -0.2, 0.5, 0.0, -0.7, 0.1])
    for _ in range(40):
        x = es.ask(16)
        es.tell(((x - target) ** 2).sum(axis=1))
    assert es.best[1] < 1e-3


@pytest.mark.gpu
def test_env_step_matches_oracle(engine, oracle_lib, blob, nominal_model):
    rng = np.random.default_rng(5)
    N = 77
    s = np.zeros((N, 37), np.float32)
    s[:, 2] = rng.uniform(0.25, 0.4, N); s[:, 6] = 1.0
    s[:, 7:13] = rng.uniform(-0.5, 0.5, (N, 6))
    s[:, 13:25] = np.array(nominal_model.q_default) * rng.uniform(0.7, 1.3, (N, 12))
    a = rng.uniform(-2, 2, (N, 12)).astype(np.float32)
    names = ["mass", "comx", "inertiay", "motor_model_hip_a", "motor_model_thigh_a", "motor_model_calf_a"]
    p = np.stack([rng.uniform(5, 11, N), rng.uniform(-0.03, 0.03, N), rng.uniform(0.05, 0.15, N), rng.uniform(15, 30, N),
                  rng.uniform(15, 30, N), rng.uniform(15, 30, N)], 1).astype(np.float32)
    st = torch.from_numpy(s.copy()).to(engine.device)
    ref = s.astype(np.float64)
    for k in range(3):
        engine.env_step(st, torch.from_numpy(a), params=torch.from_numpy(p), param_names=names,
                        motor_model="act2tau_vec3_tanh", flags=gm.FLAG_HIP_HALF)
        ref = oracle_lib.env_step(blob, ref, a, p, [gm.PARAM_IDS[n] for n in names], None, 4, 3, gm.FLAG_HIP_HALF)
    out = st.cpu().numpy()
    np.testing.assert_allclose(out[:, :7], ref[:, :7], atol=2e-5)
    np.testing.assert_allclose(out[:, 13:25], ref[:, 13:25], atol=2e-5)
    np.testing.assert_allclose(out[:, 7:13], ref[:, 7:13], atol=2e-3)


@pytest.mark.gpu
def test_evaluate_policy_gpu_matches_oracle_loop(engine, blob, nominal_model):
    """The CUDA loop (graph-captured step: cuBLAS policy + spi_b200_env_step + spi_b200_fim_reward) against the same loop
    on the oracle.  Closed loop amplifies fp32 differences, so the horizon is short and the tolerance relative."""
    cfg = act.ActiveConfig(exploration_params=["mass", "comx"], ksync_steps=5, seed=3)
    cmds = _commands(6, 40)
    ref = act.ActiveExploration(OracleActiveBackend(blob, nominal_model), act.PolicyMLP.random("cpu", seed=1), 6, cfg)
    r_ref = ref.evaluate_policy(cmds, total_steps=20, use_cuda_graph=False)
    for graph in (False, True):
        ex = act.ActiveExploration(engine, act.PolicyMLP.random(engine.device, seed=1), 6, cfg)
        r = ex.evaluate_policy(cmds, total_steps=20, use_cuda_graph=graph)
        np.testing.assert_allclose(r["total_reward"], r_ref["total_reward"], rtol=2e-2)
        np.testing.assert_allclose(r["fim"], r_ref["fim"], rtol=5e-2, atol=1e-3 * np.abs(r_ref["fim"]).max())
