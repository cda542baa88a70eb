"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances (fp32 kernel vs fp64 oracle, both integrating the same 40 sub-steps per H=5 rollout):
  per-step states   |d| <= 2e-5 absolute (positions m, quaternion, joint rad), 2e-3 on velocities
  per-sample errors |d| <= 2e-5 absolute
  mean costs        rel 5e-5 + abs 2e-6
Measured on B200 over the whole `all` dataset x 16 ten-parameter candidates (tools/dev_accuracy.py): the kernel
deviates from the fp64 oracle by at most 1.5e-7 / 2.5e-7 / 2.7e-6 (pos / quat / joint per-sample errors), the
oracle's OWN float instantiation by 1.4e-7 / 1.7e-7 / 1.7e-6 — the kernel sits on the fp32 noise floor; the
tolerances are ~10x that floor, so they are a statement about fp32, not about the kernel.
"""
import numpy as np
import pytest
import torch

from spi_active_b200 import go2_model as gm
from spi_active_b200.dataset import pack_segments, to_device

import synth

pytestmark = pytest.mark.gpu


def _segs(config, H, device, steps=None):
    S, ds = synth.dataset(config, H, steps)
    return S, ds, pack_segments(to_device(ds, device))


def test_library_loaded(engine):
    assert engine.lib.spi_b200_version() == 105


@pytest.mark.parametrize("config,H", [("stand", 5), ("sine", 5), ("jump", 3), ("all", 5)])
def test_eval_candidates_matches_oracle(engine, oracle_lib, blob, config, H):
    S, ds, segs = _segs(config, H, engine.device)
    scales = np.linspace(0.5, 2.0, 20 if config != "all" else 6).astype(np.float32)
    params = (scales * 6.921)[:, None].astype(np.float32)
    cost, per, status = engine.evaluate_candidates(torch.from_numpy(params), ["mass"], segs, return_per_seg=True,
                                                   return_status=True)
    torch.cuda.synchronize()
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    ref_cost, ref_status, ref_per = oracle_lib.eval_candidates(blob, params, [0], init, act, tgt, gains, mask,
                                                               cost_denominator=denom, return_per_seg=True)
    assert status.cpu().numpy().sum() == 0 and ref_status.sum() == 0
    np.testing.assert_allclose(per.cpu().numpy(), ref_per, atol=2e-5, rtol=0)
    np.testing.assert_allclose(cost.cpu().numpy(), ref_cost, rtol=5e-5, atol=2e-6)
    # same argmin of the weighted landscape
    w = np.array([10.0, 5.0, 1.0])
    assert int(np.argmin(cost.cpu().numpy() @ w)) == int(np.argmin(ref_cost @ w))


def test_fp32_noise_floor_of_oracle(oracle_lib, blob):
    S, ds = synth.dataset("sine", 5)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    params = np.array([[6.921], [9.0]], np.float32)
    c64, _, p64 = oracle_lib.eval_candidates(blob, params, [0], init, act, tgt, gains, mask, cost_denominator=denom,
                                             return_per_seg=True)
    c32, _, p32 = oracle_lib.eval_candidates(blob, params, [0], init, act, tgt, gains, mask, cost_denominator=denom,
                                             precision=32, return_per_seg=True)
    assert np.abs(p64 - p32).max() < 2e-5


def test_rollout_states_per_step(engine, oracle_lib, blob):
    S, ds = synth.dataset("jump", 5)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    sel = np.arange(0, S, 7)
    params = np.array([[6.921, 0.02, 0.0, -0.005], [10.0, -0.03, 0.01, 0.02]], np.float32)
    names = ["mass", "comx", "comy", "comz"]
    st = engine.rollout_states(torch.from_numpy(params), names, torch.from_numpy(init[sel]),
                               torch.from_numpy(act[sel]), torch.from_numpy(gains[sel])).cpu().numpy()
    ref = oracle_lib.rollout_states(blob, params, [gm.PARAM_IDS[n] for n in names], init[sel], act[sel], gains[sel])
    assert st.shape == ref.shape
    np.testing.assert_allclose(st[..., :7], ref[..., :7], atol=2e-5, rtol=0)       # pos, quat
    np.testing.assert_allclose(st[..., 13:25], ref[..., 13:25], atol=2e-5, rtol=0)  # joint pos
    np.testing.assert_allclose(st[..., 7:13], ref[..., 7:13], atol=2e-3, rtol=0)    # base vel
    np.testing.assert_allclose(st[..., 25:37], ref[..., 25:37], atol=1e-2, rtol=0)  # joint vel


@pytest.mark.parametrize("motor,flags", [("act2tau_scalar", 0), ("act2tau_vec3", 0), ("act2tau_vec3_tanh", 0),
                                         ("act2tau_vec3_tanh", gm.FLAG_HIP_HALF),
                                         ("act2tau_vec3_tanh", gm.FLAG_TANH_BEFORE_CLIP)])
def test_motor_models_in_rollout(engine, oracle_lib, blob, motor, flags):
    S, ds, segs = _segs("sine", 5, engine.device)
    rng = np.random.default_rng(0)
    Cn = 8
    names = ["mass", "motor_model_hip_a", "motor_model_thigh_a", "motor_model_calf_a"]
    params = np.stack([rng.uniform(5, 9, Cn), rng.uniform(0.7, 1.2, Cn) if "tanh" not in motor else rng.uniform(10, 30, Cn),
                       rng.uniform(0.7, 1.2, Cn) if "tanh" not in motor else rng.uniform(10, 30, Cn),
                       rng.uniform(0.7, 1.2, Cn) if "tanh" not in motor else rng.uniform(10, 30, Cn)], 1).astype(np.float32)
    cost = engine.evaluate_candidates(torch.from_numpy(params), names, segs, motor_model=motor, flags=flags)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    ref, _ = oracle_lib.eval_candidates(blob, params, [gm.PARAM_IDS[n] for n in names], init, act, tgt, gains, mask,
                                        motor_model=gm.MOTOR_MODELS[motor], flags=flags, cost_denominator=denom)
    np.testing.assert_allclose(cost.cpu().numpy(), ref, rtol=5e-5, atol=2e-6)


def test_full_parameter_vector(engine, oracle_lib, blob):
    """10-parameter candidates of BASELINE config 4: mass, com xyz, Ixx Iyy Izz, motor a x3."""
    S, ds, segs = _segs("all", 5, engine.device, steps=None)
    rng = np.random.default_rng(1)
    Cn = 12
    nominal = gm.default_param_vector()
    names = ["mass", "comx", "comy", "comz", "inertiax", "inertiay", "inertiaz", "motor_model_hip_a",
             "motor_model_thigh_a", "motor_model_calf_a"]
    ids = [gm.PARAM_IDS[n] for n in names]
    sigma = np.array([1.5, 0.03, 0.03, 0.03, 0.3 * nominal[4], 0.3 * nominal[5], 0.3 * nominal[6], 3, 3, 3])
    params = (nominal[ids] + rng.standard_normal((Cn, 10)) * sigma).astype(np.float32)
    params[:, 0] = np.clip(params[:, 0], 2.0, None)
    params[:, 4:7] = np.clip(params[:, 4:7], 1e-3, None)
    cost, status = engine.evaluate_candidates(torch.from_numpy(params), names, segs, motor_model="act2tau_vec3_tanh",
                                              return_status=True)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    ref, ref_status = oracle_lib.eval_candidates(blob, params, ids, init, act, tgt, gains, mask, motor_model=3,
                                                 cost_denominator=denom)
    assert status.cpu().numpy().sum() == 0
    np.testing.assert_allclose(cost.cpu().numpy(), ref, rtol=5e-5, atol=2e-6)


def test_host_entry_point_matches_device(engine):
    S, ds, segs = _segs("stand", 5, engine.device)
    params = (np.linspace(0.5, 2.0, 20) * 6.921).astype(np.float32)[:, None]
    dev = engine.evaluate_candidates(torch.from_numpy(params), ["mass"], segs).cpu().numpy()
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    host, status = engine.evaluate_candidates_host(params, ["mass"], init, act, tgt, gains, mask,
                                                   cost_denominator=denom)
    assert np.array_equal(host, dev)  # same kernel, same reduction order: bit-identical
    assert status.sum() == 0


def test_determinism(engine):
    S, ds, segs = _segs("jump", 5, engine.device)
    params = torch.linspace(4.0, 12.0, 16)[:, None]
    a = engine.evaluate_candidates(params, ["mass"], segs).clone()
    b = engine.evaluate_candidates(params, ["mass"], segs).clone()
    assert torch.equal(a, b)


def test_mask_and_ragged_sizes(engine, oracle_lib, blob):
    """S not a multiple of 8/32, masks with holes, S = 1."""
    S, ds = synth.dataset("sine", 5)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    for n in (1, 7, 33, 100):
        m = np.ones(n, np.uint8); m[::3] = 0
        if m.sum() == 0: m[0] = 1
        from spi_active_b200.dataset import SegmentBatch
        dev = engine.device
        segs = SegmentBatch(torch.from_numpy(init[:n]).to(dev), torch.from_numpy(act[:n]).to(dev),
                            torch.from_numpy(tgt[:n]).to(dev), torch.from_numpy(gains[:n]).to(dev),
                            torch.from_numpy(m).to(dev), 0.0)
        params = np.array([[6.0], [8.0], [7.0]], np.float32)
        cost = engine.evaluate_candidates(torch.from_numpy(params), ["mass"], segs).cpu().numpy()
        ref, _ = oracle_lib.eval_candidates(blob, params, [0], init[:n], act[:n], tgt[:n], gains[:n], m)
        np.testing.assert_allclose(cost, ref, rtol=5e-5, atol=2e-6)


def test_dense_packing_is_bit_identical_to_the_padded_launch():
    """The fast path gives a candidate only its S // 32 full CTAs and packs the S % 32 left-over segments of several candidates
    into shared tail CTAs (aligned groups of L' lanes, L' the power of two >= S % 32), whose partial sums use the tail of the padded
    launch's butterfly: costs, per-segment errors and status flags are bit-identical to the padded launch, for ragged S (left-overs
    of 1, 7, 4, 2 -> groups of 1, 8, 4, 2 lanes; S < 32: tail CTAs only; a left-over of 21: no packing), masks with holes, many candidates,
    a NaN row."""
    from spi_active_b200.dataset import SegmentBatch
    from spi_active_b200.engine import RolloutEngine
    from spi_active_b200 import cem
    dense, padded = RolloutEngine(), RolloutEngine()
    dense.set_kernel("ws"); padded.set_kernel("ws-padded")
    S, ds = synth.dataset("all", 5)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    cfg = cem.default_full_config(dense.model)
    rng = np.random.default_rng(11)
    dev = dense.device
    for n, C in ((1, 5), (7, 64), (33, 37), (100, 3), (245, 129), (S, 40), (66, 17)):
        m = mask[:n].copy(); m[::5] = 0
        if m.sum() == 0: m[0] = 1
        segs = SegmentBatch(torch.from_numpy(init[:n]).to(dev), torch.from_numpy(act[:n]).to(dev),
                            torch.from_numpy(tgt[:n]).to(dev), torch.from_numpy(gains[:n]).to(dev),
                            torch.from_numpy(m).to(dev), 0.0)
        params = (np.asarray(cfg.mean) + 0.3 * np.asarray(cfg.std) * rng.standard_normal((C, len(cfg.names)))).astype(np.float32)
        if C > 4:
            params[C // 2, 0] = np.nan
        out = []
        for eng in (dense, padded):
            cost, per, status = eng.evaluate_candidates(torch.from_numpy(params), cfg.names, segs, motor_model=cfg.motor_model,
                                                        return_per_seg=True, return_status=True)
            out.append((cost.cpu().numpy(), per.cpu().numpy(), status.cpu().numpy()))
        for a, b in zip(*out):
            assert np.array_equal(a, b, equal_nan=True), (n, C)
        assert out[0][2].sum() == (1 if C > 4 else 0)


def test_nonfinite_candidate_flagged(engine):
    S, ds, segs = _segs("stand", 5, engine.device)
    params = torch.tensor([[6.921], [float("nan")], [7.5]])
    cost, status = engine.evaluate_candidates(params, ["mass"], segs, return_status=True)
    status = status.cpu().numpy(); cost = cost.cpu().numpy()
    assert status.tolist() == [0, 1, 0]
    assert np.isinf(cost[1]).all() and np.isfinite(cost[[0, 2]]).all()


def test_sim_step_matches_oracle(engine, oracle_lib, blob, nominal_model):
    rng = np.random.default_rng(3)
    N = 37
    s = np.zeros((N, 37), np.float32)
    s[:, 2] = rng.uniform(0.2, 0.4, N); s[:, 6] = 1.0
    qn = rng.standard_normal((N, 4)) * 0.1 + np.array([0, 0, 0, 1.0]); s[:, 3:7] = qn / np.linalg.norm(qn, axis=1, keepdims=True)
    s[:, 7:13] = rng.uniform(-0.5, 0.5, (N, 6))
    s[:, 13:25] = np.array(nominal_model.q_default) * rng.uniform(0.8, 1.2, (N, 12))
    s[:, 25:37] = rng.uniform(-1, 1, (N, 12))
    tau = rng.uniform(-5, 5, (N, 12)).astype(np.float32)
    st = torch.from_numpy(s.copy()).to(engine.device)
    ff = torch.zeros((N, 4, 3), device=engine.device)
    engine.sim_step(st, torch.from_numpy(tau), n_steps=4, foot_force=ff)
    ref, ref_ff = oracle_lib.sim_step(blob, s.astype(np.float64), tau.astype(np.float64), 4, return_foot_force=True)
    np.testing.assert_allclose(st.cpu().numpy()[:, :7], ref[:, :7], atol=1e-4)
    np.testing.assert_allclose(st.cpu().numpy()[:, 13:25], ref[:, 13:25], atol=1e-4)
    np.testing.assert_allclose(st.cpu().numpy()[:, 7:13], ref[:, 7:13], atol=3e-3)
    np.testing.assert_allclose(ff.cpu().numpy(), ref_ff, atol=0.5, rtol=2e-2)


@pytest.mark.parametrize("motor", ["none", "act2tau_scalar", "act2tau_vec3", "act2tau_vec3_tanh"])
def test_compute_torques_matches_oracle(engine, oracle_lib, blob, motor):
    rng = np.random.default_rng(4)
    N = 257
    a = rng.uniform(-30, 30, (N, 12)).astype(np.float32)
    q = rng.uniform(-2, 2, (N, 12)).astype(np.float32)
    qd = rng.uniform(-20, 20, (N, 12)).astype(np.float32)
    gains = np.concatenate([rng.uniform(10, 40, (N, 12)), rng.uniform(0.2, 1.5, (N, 12))], 1).astype(np.float32)
    mp = (rng.uniform(0.5, 1.5, (N, 3)) if "tanh" not in motor else rng.uniform(10, 30, (N, 3))).astype(np.float32)
    for flags in (0, gm.FLAG_HIP_HALF):
        out = engine.compute_torques(torch.from_numpy(a), torch.from_numpy(q), torch.from_numpy(qd),
                                     torch.from_numpy(gains), torch.from_numpy(mp), motor, flags).cpu().numpy()
        ref = oracle_lib.compute_torques(blob, a, q, qd, gains, mp, gm.MOTOR_MODELS[motor], flags, precision=32)
        np.testing.assert_allclose(out, ref, rtol=2e-6, atol=2e-5)


def test_argument_errors_and_degenerate_sizes(engine, oracle_lib, blob):
    """The C-ABI's error behaviour (negative code + message, no exception across the boundary, nothing launched) and the
    smallest legal problem: C = 1, S = 1, H = 1, no parameters at all (the nominal model)."""
    import ctypes as C
    lib = engine.lib
    S, ds = synth.dataset("stand", 1)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    dev = engine.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_init, d_act, d_tgt, d_gains, d_mask = t(init[:1]), t(act[:1]), t(tgt[:1]), t(gains[:1]), t(mask[:1])
    cost = torch.zeros(1, 3, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    p = lambda x: None if x is None else C.c_void_p(x.data_ptr())
    ids = (C.c_int * 1)(0)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def call(model=engine._handle, params=None, Cn=1, P=0, init_=d_init, S_=1, H_=1, motor=0, out=cost):
        return lib.spi_b200_eval_candidates(model, p(params), Cn, P, ids, p(init_), p(d_act), p(d_tgt), p(d_gains), p(d_mask),
                                            S_, H_, 4, motor, 0, C.c_float(1.0), p(out), None, p(status), stream)
    before = engine.launch_count()
    wide = torch.zeros(1, 17, device=dev)
    bad_id = (C.c_int * 1)(999)
    for kwargs, needle in [(dict(model=None), b"NULL"), (dict(Cn=0), b"positive"), (dict(S_=0), b"positive"),
                           (dict(H_=0), b"positive"), (dict(init_=None), b"NULL"), (dict(out=None), b"NULL"),
                           (dict(P=1), b"params is NULL"), (dict(motor=9), b"motor_model"),
                           (dict(params=wide, P=17), b"[0,16]")]:
        rc = call(**kwargs)
        assert rc < 0, kwargs
        assert needle in lib.spi_b200_last_error(), (kwargs, lib.spi_b200_last_error())
    rc = lib.spi_b200_eval_candidates(engine._handle, p(wide), 1, 1, bad_id, p(d_init), p(d_act), p(d_tgt), p(d_gains), p(d_mask),
                                      1, 1, 4, 0, 0, C.c_float(1.0), p(cost), None, p(status), stream)
    assert rc < 0 and b"param id" in lib.spi_b200_last_error()
    assert engine.launch_count() == before                      # nothing was launched by the rejected calls
    assert call() == 0                                          # C = S = H = 1, P = 0: the nominal model
    torch.cuda.synchronize()
    nominal = np.array([[engine.model.base.mass]], np.float32)   # the same model, stated as a 1-parameter candidate
    ref, _ = oracle_lib.eval_candidates(blob, nominal, [0], init[:1], act[:1, :1], tgt[:1], gains[:1], mask[:1],
                                        cost_denominator=1.0)
    np.testing.assert_allclose(cost.cpu().numpy(), ref, rtol=5e-5, atol=2e-6)
    assert int(status.item()) == 0


@pytest.mark.gpu
def test_asymmetric_calf_runs_on_the_fast_path(oracle_lib, nominal_model):
    """The harmonic calf inertia of the fast path (SPI_WS_CALF_HARMONIC) only relies on the joint-axis / joint-offset pattern
    of the chain, not on a symmetry of the calf: a calf with a lateral centre-of-mass offset and extra products of inertia is
    accepted by 'ws' and matches the oracle."""
    import copy
    from spi_active_b200.engine import RolloutEngine
    m = copy.deepcopy(nominal_model)
    for leg in range(4):
        calf = m.leg_bodies[3 * leg + 2]
        inertia = list(calf.inertia); inertia[3] += 2e-4; inertia[5] += 2e-4
        m.leg_bodies[3 * leg + 2] = type(calf)(calf.mass, [calf.com[0], calf.com[1] + 0.03, calf.com[2]], inertia)
    eng = RolloutEngine(m)
    eng.set_kernel("ws")
    S, ds, segs = _segs("sine", 5, eng.device)
    params = (np.linspace(0.7, 1.5, 5) * 6.921).astype(np.float32)[:, None]
    cost = eng.evaluate_candidates(torch.from_numpy(params), ["mass"], segs)
    torch.cuda.synchronize()
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    ref_cost, _ = oracle_lib.eval_candidates(eng.blob, params, [0], init, act, tgt, gains, mask, cost_denominator=denom)
    np.testing.assert_allclose(cost.cpu().numpy(), ref_cost, rtol=5e-5, atol=2e-6)
