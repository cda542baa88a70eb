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


class OracleTensorBackend(OracleActiveBackend):
    """+ fim_contract with the signature of RolloutEngine.fim_contract (fp64 einsum): host logic of the deferred FIM."""

    def fim_contract(self, hist, delta, live=None, out_JtJ=None, out_trace=None, accumulate=False):
        h = hist.numpy().astype(np.float64)
        J = (h[:, :, 0:1, :] - h[:, :, 1:, :]) / float(delta)
        if live is not None:
            J = np.where(live.numpy().astype(bool)[:, :, None, None], J, 0.0)
        jtj = torch.from_numpy(np.einsum("tmpd,tmqd->mpq", J, J).astype(np.float32))
        tr = torch.from_numpy((J * J).sum(axis=(0, 2, 3)).astype(np.float32))
        if accumulate:
            out_JtJ.add_(jtj); out_trace.add_(tr)
        else:
            out_JtJ.copy_(jtj); out_trace.copy_(tr)
        return out_JtJ, out_trace


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


def test_ring_col_map_reproduces_the_observation():
    """The ring formulation of the actor input (spi_b200_policy_forward_ring): pushing one frame per step into slot
    (-step) % 15 and reading through ring_col_map gives exactly the observation the reference builds from its history
    buffers (frame | per-key history blocks), for every head position."""
    rng = np.random.default_rng(0)
    gather = act._history_gather_index().numpy()
    cm = act.ring_col_map()
    assert cm.shape == (act.RING_SLOTS, 900)
    history = np.zeros((14, 60))
    ring = np.zeros((act.RING_SLOTS, 60))
    for step in range(40):
        frame = rng.standard_normal(60)
        obs = np.concatenate([frame, history.reshape(-1)[gather]])           # legged_robot_base.py:511-527, 819-829
        h = (-step) % act.RING_SLOTS
        ring[h] = frame
        flat = ring.reshape(-1)
        np.testing.assert_array_equal(obs[cm[h]], flat)                      # ring element k carries obs[cm[h, k]]
        assert sorted(cm[h].tolist()) == list(range(900))
        history = np.concatenate([frame[None], history[:-1]])                # history_handler.add


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


def test_pipelined_exploration_equals_single_explorer_on_oracle(blob, nominal_model):
    """PipelinedExploration cuts the trials into independent explorers; the seeded reset draw is made once for the whole
    population, so trial j scores the same as in the uncut explorer."""
    cfg = act.ActiveConfig(exploration_params=["mass", "comx"], ksync_steps=5, seed=3)
    cmds = _commands(5, 30, seed=4)
    be = OracleActiveBackend(blob, nominal_model)
    one = act.ActiveExploration(be, act.PolicyMLP.random("cpu", seed=1), 5, cfg).evaluate_policy(cmds, total_steps=12)
    pipe = act.PipelinedExploration(be, act.PolicyMLP.random("cpu", seed=1), 5, cfg, n_pipelines=2)
    assert [sl.stop - sl.start for sl in pipe.slices] == [2, 3] and pipe.num_envs == 15
    two = pipe.evaluate_policy(cmds, total_steps=12)
    assert two["steps"] == one["steps"]
    # (torch's CPU GEMM blocks a 6-row and a 15-row batch differently: last-bit differences in the actor output)
    np.testing.assert_allclose(two["total_reward"], one["total_reward"], rtol=2e-5)
    np.testing.assert_allclose(two["fim"], one["fim"], rtol=2e-5, atol=1e-6 * np.abs(one["fim"]).max())


def test_deferred_tensor_fim_equals_per_step_fim(blob, nominal_model):
    """fim_mode='tensor' (record states, contract every fim_chunk steps) gives the per-step accumulation's numbers,
    including a group that terminates mid-rollout (its later steps score termination_rew = 0)."""
    cmds = _commands(4, 40)

    def run(mode, backend_cls, tip=True, **kw):
        cfg = act.ActiveConfig(exploration_params=["mass", "comx", "motor_model_calf_a"], ksync_steps=5, seed=3,
                               fim_mode=mode, **kw)
        ex = act.ActiveExploration(backend_cls(blob, nominal_model), act.PolicyMLP.random("cpu", seed=1), 4, cfg)
        orig = ex._policy_step
        calls = {"n": 0}

        def tipping_step():            # roll group 1's main env over at step 12 -> terminated from then on
            calls["n"] += 1
            if tip and calls["n"] == 12:
                ex.state[4, 3:7] = torch.tensor([0.70710678, 0.0, 0.0, 0.70710678])
            orig()
        ex._policy_step = tipping_step
        return ex.evaluate_policy(cmds, total_steps=30, use_cuda_graph=False)

    a = run("step", OracleActiveBackend)
    for chunk in (7, 64):
        b = run("tensor", OracleTensorBackend, fim_chunk=chunk)
        assert b["steps"] == a["steps"]
        np.testing.assert_allclose(b["total_reward"], a["total_reward"], rtol=2e-5)
        np.testing.assert_allclose(b["fim"], a["fim"], rtol=2e-5, atol=1e-6 * np.abs(a["fim"]).max())
    r, r0 = a["total_reward"].reshape(4, 4), run("step", OracleActiveBackend, tip=False)["total_reward"].reshape(4, 4)
    assert r[1, 0] < 0.95 * r0[1, 0]                            # the tipped group stopped collecting reward
    np.testing.assert_array_equal(np.delete(r, 1, axis=0), np.delete(r0, 1, axis=0))
    with pytest.raises(ValueError):
        act.ActiveExploration(OracleActiveBackend(blob, nominal_model), act.PolicyMLP.random("cpu"), 2,
                              act.ActiveConfig(fim_mode="tensor"))


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


@pytest.mark.parametrize("k,K", [(5, 1), (1, 1), (0, 1), (5, 64), (3, 7)])
def test_step_schedule_rows_follow_the_per_step_definition(k, K):
    """The step schedule uploaded once per rollout shape (active._build_schedule, vectorised and cached) holds, row by row, what
    the host computes step by step in _advance_inputs: command row (clamped to the last one), k-sync flag
    (active_sysid_openloop.py:247-252), the slot of the Fisher history ring and the head of the observation ring."""
    from types import SimpleNamespace
    ns = SimpleNamespace(cfg=SimpleNamespace(ksync_steps=k), hist=(torch.zeros(K, 1) if K > 1 else None),
                         device=torch.device("cpu"), counter=torch.ones(1, dtype=torch.int32), _graph="captured")
    n_calls, T = 57, 50
    act.ActiveExploration._build_schedule(ns, n_calls, T)
    rows = ns.schedule.numpy()
    assert rows.shape == (n_calls + 4, 4) and int(ns.counter[0]) == 0 and ns._graph is None
    for i in range(n_calls + 4):
        idx = i + 1
        want = (min(idx, T - 1), int((k == 1) or (k > 1 and idx % k == 1)), 0 if i == 0 else (i - 1) % K,
                (-i) % act.RING_SLOTS)
        assert tuple(int(v) for v in rows[i]) == want, (i, rows[i], want)
    # the same shape again: nothing is rebuilt (a captured step keeps its buffer), only the step counter restarts
    buf, ns._graph = ns.schedule, "captured"
    ns.schedule[:, 0] = -7
    ns.counter.fill_(9)
    act.ActiveExploration._build_schedule(ns, n_calls, T)
    assert ns.schedule is buf and ns._graph == "captured" and int(ns.counter[0]) == 0 and int(ns.schedule[0, 0]) == -7
    # another command length: rebuilt in place (same buffer: it is large enough)
    act.ActiveExploration._build_schedule(ns, n_calls, T - 10)
    assert ns.schedule is buf and int(ns.schedule[n_calls - 1, 0]) == T - 11


def test_command_samplers():
    c = act.commands_constant(np.array([[1.0, 2.0], [3.0, 4.0]]), 3)
    assert c.shape == (6, 2) and c[:3, 0].tolist() == [1, 1, 1] and c[3:, 1].tolist() == [4, 4, 4]
    full = act.expand_commands(np.ones((5, 3), np.float32) * 0.7)
    assert full.shape == (5, 14) and np.allclose(full[:, [0, 2, 5]], 0.7) and np.allclose(full[:, 4], 3.0)
    rng = np.random.default_rng(0)
    r = np.asarray(act.COMMAND_RANGES, np.float32)[act.COMMAND_SAMPLING_IDXS]
    p = act.commands_polynomial(rng.uniform(-2, 2, (5, 3, 4)), r, 250)
    assert p.shape == (1250, 3) and (p >= r[:, 0] - 1e-6).all() and (p <= r[:, 1] + 1e-6).all()
    # bezier: one curve PER update window, clipped to the range (active_sysid.py:362-400)
    pts = np.array([[[0.0], [1.0], [1.0], [0.0]], [[0.5], [3.0], [3.0], [0.5]]], np.float32)
    b = act.commands_bezier(pts, np.array([[0.0, 1.0]], np.float32), 101)
    assert b.shape == (202, 1) and abs(b[0, 0]) < 1e-7 and abs(b[100, 0]) < 1e-7 and abs(b[50, 0] - 0.75) < 1e-6
    assert abs(b[101, 0] - 0.5) < 1e-7 and b[101:].max() == 1.0                   # second window restarts at t = 0, clipped


@pytest.mark.parametrize("mode", ["constant", "polynomial", "bezier"])
def test_sample_commands_matches_reference(mode):
    """tests/golden/commands.npz: the reference's real ActiveSysId.sample_commands (active_sysid.py:259-400) driven by a stub
    `trial.suggest_float` that replays a recorded sequence — same names, same call order, same commands."""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "commands.npz", allow_pickle=True)
    names, values = list(g[f"{mode}_names"]), g[f"{mode}_values"]
    seen = []
    it = iter(values.tolist())

    def suggest(name, low, high):
        seen.append(name)
        return next(it)
    kw = dict(num_command_updates=int(g["num_updates"]), num_steps_per_update=int(g["steps_per_update"]))
    cmds = act.sample_commands(suggest, mode, **kw)
    assert seen == names
    np.testing.assert_allclose(cmds, g[f"{mode}_commands"], rtol=0, atol=1e-6)
    sp_names, lo, hi = act.search_space(mode, **kw)
    assert sp_names == names
    np.testing.assert_allclose(lo, g[f"{mode}_low"]); np.testing.assert_allclose(hi, g[f"{mode}_high"])
    np.testing.assert_allclose(act.commands_from_vector(values, mode, **kw), g[f"{mode}_commands"], rtol=0, atol=1e-6)


def _active_gloo_worker(rank, world, port, out_dir):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spi_active_b200 import go2_model as gm2
    model = gm2.go2_nominal()
    be = OracleActiveBackend(gm2.build_model_blob(model), model)
    ex = act.ActiveExploration(be, act.PolicyMLP.random("cpu", seed=1), 4 // world,
                               act.ActiveConfig(exploration_params=["mass", "comx"], seed=3, randomize_reset=False))
    res = act.optimize_commands(ex, iterations=2, rollout_length=0.4, horizon_length=0.2, seed=0, total_steps=12,
                                rank=rank, world=world, mode="bezier" if out_dir.endswith("bezier") else "constant")
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), history=res["history"], best=res["best_commands"],
             fim=res["best_fim"])
    dist.destroy_process_group()


def test_sharded_command_search_over_gloo_matches_single_process(tmp_path):
    """Config 5 at world_size 2 (gloo, CPU oracle physics): trials sharded over ranks, rewards + Fisher blocks
    all-gathered; every rank ends with the single-process result."""
    import socket
    import torch.multiprocessing as mp
    outs = {}
    for world in (1, 2):
        d = tmp_path / f"w{world}_bezier"          # the per-window bezier sampler drives this search (worker reads the suffix)
        d.mkdir()
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
        mp.spawn(_active_gloo_worker, args=(world, port, str(d)), nprocs=world, join=True)
        outs[world] = [np.load(d / f"rank{r}.npz") for r in range(world)]
    a, b0, b1 = outs[1][0], outs[2][0], outs[2][1]
    for k in ("history", "best", "fim"):
        np.testing.assert_array_equal(b0[k], b1[k])
        # the policy GEMMs see a different batch size per rank: fp32 summation order differs, the closed loop amplifies it
        np.testing.assert_allclose(b0[k], a[k], rtol=5e-4)
    assert a["fim"].shape == (2, 2) and a["history"].shape == (2,) and (a["history"] > 0).all()


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
    import dataclasses
    for graph in (False, True):
        res = {}
        for mode in ("step", "tensor"):
            ex = act.ActiveExploration(engine, act.PolicyMLP.random(engine.device, seed=1), 6,
                                       dataclasses.replace(cfg, fim_mode=mode, fim_chunk=8, step_impl="torch"))
            assert ex.fim_mode == mode and ex.step_impl == "torch"
            r = res[mode] = ex.evaluate_policy(cmds, total_steps=20, use_cuda_graph=graph)
            np.testing.assert_allclose(r["total_reward"], r_ref["total_reward"], rtol=2e-2)
            np.testing.assert_allclose(r["fim"], r_ref["fim"], rtol=5e-2, atol=1e-3 * np.abs(r_ref["fim"]).max())
        # same physics, two FIM kernels (CUDA cores per step vs tcgen05 over the recorded states): fp32-tight
        np.testing.assert_allclose(res["tensor"]["total_reward"], res["step"]["total_reward"], rtol=1e-5)
        np.testing.assert_allclose(res["tensor"]["fim"], res["step"]["fim"], rtol=1e-4,
                                   atol=1e-5 * np.abs(res["step"]["fim"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("policy_impl", ["cublas", "tensor", "tensor-ring"])
def test_fused_post_step_kernel_matches_torch_statement(engine, policy_impl):
    """spi_b200_active_post_step (one kernel per step) against the torch statement of the same step (which the golden
    vectors pin to the reference's code): every piece of state after each of 12 closed-loop steps, with a k-sync in the
    window, one group tipped over, and the history ring wrapping."""
    import dataclasses
    base = act.ActiveConfig(exploration_params=["mass", "comx", "motor_model_calf_a"], ksync_steps=5, seed=3,
                            fim_mode="tensor", fim_chunk=4)
    cmds = _commands(5, 40, seed=2)
    cmds[:, 20:, 0] *= -1.0                      # the command rows change inside the window
    ring = policy_impl == "tensor-ring"          # the observation lives as a ring of pre-split frames (obs_ring)
    policy_impl = policy_impl.split("-")[0]
    base = dataclasses.replace(base, obs_ring=ring)
    exs = {}
    for impl in ("torch", "fused"):
        ex = exs[impl] = act.ActiveExploration(engine, act.PolicyMLP.random(engine.device, seed=1), 5,
                                               dataclasses.replace(base, step_impl=impl, policy_impl=policy_impl))
        assert ex.step_impl == impl
        ex.reset_all(cmds, total_steps=30)
    a, b = exs["torch"], exs["fused"]
    assert (b.tc_policy is not None) == (policy_impl == "tensor") and b.ring == ring
    # same actor arithmetic (cuBLAS on both sides): the post-step kernel must agree to fp32 rounding; with the 3xTF32
    # actor the actions differ by ~2e-6, which a stiff foot contact turns into ~1e-4 on a velocity within one step
    vel_tol = 2e-5 if policy_impl == "cublas" else 2e-3
    vel = np.zeros(37, bool); vel[7:13] = True; vel[25:37] = True
    frame_vel = np.zeros(60, bool); frame_vel[12:15] = True; frame_vel[45:57] = True      # base_ang_vel, dof_vel terms

    def compare(tag):
        b.materialize_observation()              # ring mode: obs / history <- the ring (a no-op otherwise)
        for name in ("state", "actions", "obs", "history", "commands", "done", "dead_steps"):
            x, y = getattr(a, name).float().cpu().numpy(), getattr(b, name).float().cpu().numpy()
            if name == "state":
                np.testing.assert_allclose(y[:, ~vel], x[:, ~vel], rtol=2e-5, atol=2e-5, err_msg=f"{tag}: {name}")
                np.testing.assert_allclose(y[:, vel], x[:, vel], rtol=vel_tol, atol=vel_tol, err_msg=f"{tag}: {name} (vel)")
            elif name in ("obs", "history"):
                m = np.tile(frame_vel, y.shape[-1] // 60) if name == "obs" else frame_vel
                if name == "obs":   # [frame | per-key history blocks]: compare everything at the velocity tolerance
                    np.testing.assert_allclose(y, x, rtol=vel_tol, atol=vel_tol, err_msg=f"{tag}: {name}")
                else:
                    np.testing.assert_allclose(y[..., ~m], x[..., ~m], rtol=2e-5, atol=2e-5, err_msg=f"{tag}: {name}")
                    np.testing.assert_allclose(y[..., m], x[..., m], rtol=vel_tol, atol=vel_tol, err_msg=f"{tag}: {name}")
            else:
                np.testing.assert_allclose(y, x, rtol=2e-5, atol=2e-5, err_msg=f"{tag}: {name}")
        # the fused kernel already holds the gait clock of the NEXT step (the torch path advances it at the start of it)
        g, clk = act.step_contact_targets(a.gait_indices, a.commands, a.dt)
        np.testing.assert_allclose(b.gait_indices.cpu().numpy(), g.cpu().numpy(), atol=2e-6, err_msg=f"{tag}: gait")
        np.testing.assert_allclose(b.clock.cpu().numpy(), clk.cpu().numpy(), atol=2e-5, err_msg=f"{tag}: clock")
    compare("reset")
    for k in range(12):
        if k == 6:                               # tip the main env of group 2 in both
            for ex in (a, b):
                ex.state[2 * 4, 3:7] = torch.tensor([0.70710678, 0.0, 0.0, 0.70710678], device=engine.device)
        a._advance_inputs(); a._policy_step(); a._hist_count = (a._hist_count + 1) % 4
        b.step_idx += 1; b._policy_step()
        compare(f"step {k}")
        if b.tc_policy is not None and not ring:     # hi + lo is the observation; the padding stays zero
            np.testing.assert_allclose(b.tc_policy.unsplit_input(b.obs_hi, b.obs_lo, b.num_envs).cpu().numpy(),
                                       b.obs.cpu().numpy(), rtol=1e-6, atol=1e-7)     # an fp16 pair carries 22 bits
            n_written = b.num_envs * 900                                          # the padding stays zero
            assert int((b.obs_hi != 0).sum()) <= n_written and int((b.obs_lo != 0).sum()) <= n_written
        # one-step check: restart the fused copy from the torch path's exact state (the closed loop amplifies the
        # ~1e-7 rounding differences of the observation arithmetic by an order of magnitude per few steps)
        for name in ("state", "obs", "history", "done", "dead_steps"):
            getattr(b, name).copy_(getattr(a, name))
        g, clk = act.step_contact_targets(a.gait_indices, a.commands, a.dt)
        b.gait_indices.copy_(g); b.clock.copy_(clk)
        if ring:
            b.load_observation(a.obs, a.history)
        elif b.tc_policy is not None:     # the tensor-core actor reads the pre-split observation
            b.tc_policy.split_input(b.obs, b.obs_hi, b.obs_lo)
        np.testing.assert_allclose(b.hist.cpu().numpy(), a.hist.cpu().numpy(), rtol=2e-5, atol=2e-5)
        np.testing.assert_array_equal(b.live_hist.cpu().numpy(), a.live_hist.cpu().numpy())
    assert a.done.view(5, 4)[2].all() and not a.done.view(5, 4)[[0, 1, 3, 4]].any()


@pytest.mark.gpu
def test_evaluate_policy_fused_matches_torch_path(engine):
    import dataclasses
    base = act.ActiveConfig(exploration_params=["mass", "comx"], ksync_steps=5, seed=3, fim_mode="tensor", fim_chunk=8)
    cmds = _commands(6, 40)
    out = {}
    for impl in ("torch", "fused"):
        for graph in (False, True):
            ex = act.ActiveExploration(engine, act.PolicyMLP.random(engine.device, seed=1), 6,
                                       dataclasses.replace(base, step_impl=impl))
            out[impl, graph] = ex.evaluate_policy(cmds, total_steps=25, use_cuda_graph=graph)
            if graph:   # a second rollout through the captured step: the schedule / counter restart cleanly
                again = ex.evaluate_policy(cmds, total_steps=25, use_cuda_graph=True)
                np.testing.assert_array_equal(again["total_reward"], out[impl, graph]["total_reward"])
    ref = out["torch", False]
    for key, r in out.items():
        assert r["steps"] == ref["steps"]
        # closed loop over 24 steps: cuBLAS inside / outside a graph already differs by 1 % on the torch path itself
        np.testing.assert_allclose(r["total_reward"], ref["total_reward"], rtol=3e-2, err_msg=str(key))
        np.testing.assert_allclose(r["fim"], ref["fim"], rtol=5e-2, atol=2e-3 * np.abs(ref["fim"]).max(), err_msg=str(key))


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_fused_fisher_accumulation_matches_tensor_core_contraction(engine, graph):
    """fim_mode='fused' (J J^T accumulated inside spi_b200_active_post_step from the rows it already holds, no state history)
    against fim_mode='tensor' (history + tcgen05 spi_b200_fim_contract): same physics, same actor, so rewards and Fisher
    blocks agree to fp32 summation order — incl. a group that terminates mid-rollout (its later steps score 0) and P = 10."""
    import dataclasses
    base = act.ActiveConfig(exploration_params=list(act.ActiveExploration.PARAM_ORDER), ksync_steps=5, seed=3, fim_chunk=8)
    cmds = _commands(9, 60, seed=4)
    res = {}
    for mode in ("tensor", "fused"):
        ex = act.ActiveExploration(engine, act.PolicyMLP.random(engine.device, seed=1), 9, dataclasses.replace(base, fim_mode=mode))
        assert ex.fim_mode == mode and ex.step_impl == "fused" and (mode == "tensor") == (ex.hist is not None)
        n = ex.begin_rollout(cmds, 40, graph)
        for k in range(n):
            if k == 11:
                ex.state[3 * 11, 3:7] = torch.tensor([0.70710678, 0.0, 0.0, 0.70710678], device=engine.device)   # tip group 3
            ex.advance_rollout()
        res[mode] = ex.finish_rollout()
        if graph:      # a second rollout through the captured step restarts cleanly
            again = ex.evaluate_policy(cmds, total_steps=40, use_cuda_graph=True)
            first = ex.evaluate_policy(cmds, total_steps=40, use_cuda_graph=True)
            np.testing.assert_array_equal(again["total_reward"], first["total_reward"])
    a, b = res["tensor"], res["fused"]
    assert a["steps"] == b["steps"]
    np.testing.assert_allclose(b["total_reward"], a["total_reward"], rtol=2e-5)
    np.testing.assert_allclose(b["fim"], a["fim"], rtol=1e-4, atol=2e-6 * np.abs(a["fim"]).max())
    assert b["total_reward"][3 * 11] != b["total_reward"][0]


@pytest.mark.gpu
def test_pipelined_exploration_matches_single_explorer_gpu(engine):
    """Two independent explorers replaying their captured steps on two streams give the same rewards / Fisher blocks as one
    explorer on the whole population (every kernel is deterministic per row; only the FIM chunking of the tensor-core
    contraction sees a different tile composition: 8 main envs per CTA)."""
    cfg = act.ActiveConfig(exploration_params=["mass", "comx", "motor_model_calf_a"], ksync_steps=5, seed=3, fim_chunk=8)
    cmds = _commands(24, 40, seed=6)
    pol = act.PolicyMLP.random(engine.device, seed=1)
    one = act.ActiveExploration(engine, pol, 24, cfg).evaluate_policy(cmds, total_steps=30)
    pipe = act.PipelinedExploration(engine, pol, 24, cfg, n_pipelines=2)
    two = pipe.evaluate_policy(cmds, total_steps=30)
    again = pipe.evaluate_policy(cmds, total_steps=30)
    assert two["steps"] == one["steps"] == 29
    np.testing.assert_array_equal(again["total_reward"], two["total_reward"])
    np.testing.assert_allclose(two["total_reward"], one["total_reward"], rtol=1e-5)
    np.testing.assert_allclose(two["fim"], one["fim"], rtol=1e-5, atol=1e-6 * np.abs(one["fim"]).max())
