"""CPU oracle and host logic against the golden vectors produced by the reference's own Python
(tests/golden/make_golden.py lists, per file, the reference functions that were executed).

Tolerances: windowing is exact; fp32 torch arithmetic of the reference vs the fp64 oracle agrees to fp32
rounding (rtol 2e-6 / atol 2e-5 on torques up to 35 N m); the end-to-end mass sweep agrees to 2e-6 absolute
(the reference path stores the state in fp32 tensors and sums the errors in fp32, the oracle runs in fp64).
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from spi_active_b200 import dataset as dsmod
from spi_active_b200 import go2_model as gm

GOLD = Path(__file__).resolve().parent / "golden"


def _gains(g, N):
    return np.tile(np.concatenate([g["kp"], g["kd"]])[None], (N, 1)).astype(np.float32)


@pytest.mark.parametrize("precision", [64, 32])
def test_torque_law_and_motor_models(oracle_lib, blob, precision):
    g = np.load(GOLD / "torques.npz")
    N = g["actions"].shape[0]
    gains = _gains(g, N)
    tol = dict(rtol=2e-6, atol=2e-5) if precision == 64 else dict(rtol=3e-6, atol=3e-5)
    # the clip of _pre_physics_step (legged_robot_base.py:186-187)
    np.testing.assert_array_equal(g["clipped_actions"], np.clip(g["actions"], -20.0, 20.0))
    cases = [
        ("tau_base", "none", 0, None), ("tau_omni", "none", gm.FLAG_HIP_HALF, None),
        ("tau_scalar_base", "act2tau_scalar", 0, np.repeat(g["scalar_gain"][:, None], 3, 1)),
        ("tau_scalar_omni", "act2tau_scalar", gm.FLAG_HIP_HALF, np.repeat(g["scalar_gain"][:, None], 3, 1)),
        ("tau_vec3_base", "act2tau_vec3", 0, g["vec3_gain"]), ("tau_vec3_omni", "act2tau_vec3", gm.FLAG_HIP_HALF, g["vec3_gain"]),
        ("tau_tanh_base", "act2tau_vec3_tanh", 0, g["tanh_a"]), ("tau_tanh_omni", "act2tau_vec3_tanh", gm.FLAG_HIP_HALF, g["tanh_a"]),
    ]
    for key, motor, flags, mp in cases:
        out = oracle_lib.compute_torques(blob, g["actions"], g["q"], g["qd"], gains, mp, gm.MOTOR_MODELS[motor], flags,
                                         precision=precision)
        np.testing.assert_allclose(out, g[key], err_msg=key, **tol)


@pytest.mark.parametrize("H", [3, 5])
def test_load_dataset_matches_reference(tmp_path, H, capsys):
    g = np.load(GOLD / "windowing.npz", allow_pickle=True)
    paths = []
    for i in range(2):
        rec = {k[len(f"in{i}_"):]: g[k] for k in g.files if k.startswith(f"in{i}_")}
        p = tmp_path / f"rec{i}.npz"
        np.savez(p, **rec)
        paths.append(p)
    total, ds = dsmod.load_dataset(paths, H)
    assert total == int(g[f"H{H}_total"])
    keys = [k[len(f"H{H}_"):] for k in g.files if k.startswith(f"H{H}_") and k != f"H{H}_total"]
    assert sorted(keys) == sorted(ds.keys())
    for k in keys:
        ref = g[f"H{H}_{k}"]
        assert ds[k].shape == ref.shape and ds[k].dtype == ref.dtype, k
        np.testing.assert_array_equal(ds[k], ref, err_msg=k)
    assert "Loaded 2 trajectory(s)" in capsys.readouterr().out   # same progress line as eval.py:169


def test_fim_reward_restatement(oracle_lib):
    g = np.load(GOLD / "fim.npz")
    rew, JJt = oracle_lib.fim_reward(g["root"], g["dof"], g["origins"], int(g["M"]), int(g["P"]), float(g["delta"]))
    np.testing.assert_allclose(rew, g["reward"], rtol=2e-5)
    np.testing.assert_allclose(np.trace(JJt, axis1=1, axis2=2), g["reward"][:: int(g["P"]) + 1], rtol=2e-5)


def test_weighted_cost_and_sweep_constants(oracle_lib):
    g = np.load(GOLD / "cost.npz")
    from spi_active_b200 import cem
    assert tuple(g["weights"]) == cem.COST_WEIGHTS
    np.testing.assert_allclose(oracle_lib.weighted_cost(g["costs"], cem.COST_WEIGHTS), g["totals"], rtol=1e-12)
    from spi_active_b200.landscape import MASS_SAMPLES, MASS_SCALE_MAX, MASS_SCALE_MIN, MAX_SAFE_ENV_BATCH
    assert (MASS_SCALE_MIN, MASS_SCALE_MAX) == tuple(g["mass_scale_range"])
    assert MASS_SAMPLES == int(g["mass_samples"]) and MAX_SAFE_ENV_BATCH == int(g["max_safe_env_batch"])


def test_quat_rotate_inverse(oracle_lib):
    g = np.load(GOLD / "quat.npz")
    np.testing.assert_allclose(oracle_lib.quat_rotate_inverse(g["q"], g["v"]), g["rotate_inverse"], atol=2e-6)


def _sweep_dataset(g, H, tmp_path):
    paths = []
    for i in range(3):
        rec = {k[len(f"rec{i}_"):]: g[k] for k in g.files if k.startswith(f"rec{i}_")}
        p = tmp_path / f"rec{i}.npz"
        np.savez(p, **rec)
        paths.append(p)
    return dsmod.load_dataset(paths, H)


@pytest.mark.parametrize("H,B", [(5, 16), (5, 32), (5, 4096), (3, 16), (3, 32), (3, 4096)])
def test_mass_sweep_matches_reference_pipeline(oracle_lib, blob, tmp_path, H, B):
    """The reference's mass_sweep -> apply_base_mass -> evaluate_batch -> LeggedRobotBase control path, executed
    for real on the oracle physics, against oracle.eval_candidates fed by OUR windowing/packing with
    strict_reference=True (quirks D2: chunk-dependent mask, D3: gains of row 0 of each chunk)."""
    g = np.load(GOLD / "mass_sweep.npz")
    total, ds = _sweep_dataset(g, H, tmp_path)
    batch = int(g[f"H{H}_B{B}_batch"])
    assert batch == min(B, total, 8192)
    segs = dsmod.pack_segments(dsmod.to_device(ds, "cpu"), env_batch=batch, strict_reference=True)
    params = (g["scales"] * float(g["ref_masses"][0])).astype(np.float32)[:, None]
    cost, status = oracle_lib.eval_candidates(
        blob, params, [gm.PARAM_IDS["mass"]], segs.seg_init.numpy(), segs.seg_actions.numpy(), segs.seg_target.numpy(),
        segs.seg_gains.numpy(), segs.seg_mask.numpy(), cost_denominator=segs.cost_denominator)
    assert status.sum() == 0
    np.testing.assert_allclose(cost, g[f"H{H}_B{B}_costs"], rtol=0, atol=2e-6)


def test_strict_mask_differs_from_default():
    """Sanity of the quirk itself: with B >= S only the first file contributes under the literal mask."""
    me = torch.zeros(30, dtype=torch.bool)
    me[9] = me[19] = me[29] = True
    strict = dsmod.reference_eval_mask(me, 4096, True)
    assert strict.sum().item() == 9 and strict[:9].all()
    chunked = dsmod.reference_eval_mask(me, 8, True)
    # chunks [0:8) [8:16) [16:24) [24:30): masked from the first boundary of each chunk on
    expect = torch.ones(30, dtype=torch.bool)
    expect[9:16] = False; expect[19:24] = False; expect[29:] = False
    assert torch.equal(chunked, expect)
    assert dsmod.reference_eval_mask(me, None, False).sum().item() == 27


def test_model_blob_from_the_reference_urdf(nominal_model):
    """The typed-in Go2 table (go2_nominal) against the blob built from the reference's own go2.urdf by model_from_urdf
    (tests/golden/urdf_blob.npz, make_golden.gen_urdf_blob): bit-identical.  With the reference checkout present (build
    container) the URDF is parsed again here."""
    g = np.load(GOLD / "urdf_blob.npz")
    blob = gm.build_model_blob(nominal_model)
    assert blob.shape == (gm.BLOB["SIZE"],) and blob.dtype == np.float32
    np.testing.assert_array_equal(blob, g["blob"])
    np.testing.assert_array_equal(nominal_model.body_masses_isaac_order(), g["body_masses"])
    assert abs(float(g["total_mass"]) - 15.019) < 1e-3
    urdf = Path("/root/reference/spigym/data/robots/go2/urdf/go2.urdf")
    if urdf.exists():
        np.testing.assert_array_equal(gm.build_model_blob(gm.model_from_urdf(urdf)), g["blob"])


def test_policy_from_reference_checkpoint():
    """PolicyMLP.from_checkpoint on a checkpoint written in the layout of PPO.save (agents/ppo/ppo.py:155-166) from the
    reference's real PPOActor; outputs = its act_inference (agents/modules/ppo_modules.py:73-75) to fp32 rounding."""
    from spi_active_b200.active import PolicyMLP
    g = np.load(GOLD / "ppo_actor.npz")
    pol = PolicyMLP.from_checkpoint(GOLD / "ppo_actor.pt", "cpu")
    assert [tuple(w.shape) for w in pol.weights] == [(32, 60), (16, 32), (8, 16), (12, 8)]
    assert [k for k in g["keys"]][0] == "std"                       # the action-noise parameter is skipped
    out = pol(torch.from_numpy(g["obs"]))
    np.testing.assert_allclose(out.numpy(), g["actions"], rtol=1e-5, atol=1e-6)
