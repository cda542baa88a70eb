"""GPU: the CUDA path (through the C-ABI) against the golden vectors produced by executing the reference's own
Python (tests/golden/make_golden.py).  fp32 kernel vs fp32 torch reference: rtol 3e-6 / atol 3e-5 on torques
(|tau| <= 35.55 N m); 2e-4 relative on the FIM reward (25 x P squared differences scaled by 1/delta^2);
2e-5 absolute on the end-to-end sweep costs (40 fp32 integration sub-steps)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from spi_active_b200 import dataset as dsmod
from spi_active_b200 import go2_model as gm
from spi_active_b200 import landscape

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def test_torque_kernel_matches_reference_torch(engine):
    g = np.load(GOLD / "torques.npz")
    N = g["actions"].shape[0]
    gains = np.tile(np.concatenate([g["kp"], g["kd"]])[None], (N, 1)).astype(np.float32)
    t = torch.from_numpy
    cases = [
        ("tau_base", "none", 0, None), ("tau_omni", "none", gm.FLAG_HIP_HALF, None),
        ("tau_scalar_base", "act2tau_scalar", 0, np.repeat(g["scalar_gain"][:, None], 3, 1)),
        ("tau_scalar_omni", "act2tau_scalar", gm.FLAG_HIP_HALF, np.repeat(g["scalar_gain"][:, None], 3, 1)),
        ("tau_vec3_base", "act2tau_vec3", 0, g["vec3_gain"]), ("tau_vec3_omni", "act2tau_vec3", gm.FLAG_HIP_HALF, g["vec3_gain"]),
        ("tau_tanh_base", "act2tau_vec3_tanh", 0, g["tanh_a"]), ("tau_tanh_omni", "act2tau_vec3_tanh", gm.FLAG_HIP_HALF, g["tanh_a"]),
    ]
    for key, motor, flags, mp in cases:
        out = engine.compute_torques(t(g["actions"]), t(g["q"]), t(g["qd"]), t(gains), None if mp is None else t(mp),
                                     motor, flags).cpu().numpy()
        np.testing.assert_allclose(out, g[key], rtol=3e-6, atol=3e-5, err_msg=key)


def test_fim_kernel_matches_reference_torch(engine, oracle_lib):
    g = np.load(GOLD / "fim.npz")
    M, P = int(g["M"]), int(g["P"])
    states = oracle_lib.fim_states(g["root"], g["dof"], g["origins"], M, P)
    JJt, trace = engine.fim_reward(torch.from_numpy(states), float(g["delta"]))
    np.testing.assert_allclose(trace.cpu().numpy(), g["reward"][:: P + 1], rtol=2e-4)
    _, ref_JJt = oracle_lib.fim_reward(g["root"], g["dof"], g["origins"], M, P, float(g["delta"]))
    np.testing.assert_allclose(JJt.cpu().numpy(), ref_JJt, rtol=2e-4, atol=1e-2)


def test_weighted_cost_kernel(engine):
    g = np.load(GOLD / "cost.npz")
    out = engine.weighted_cost(torch.from_numpy(g["costs"])).cpu().numpy()
    np.testing.assert_allclose(out, g["totals"], rtol=2e-6)


@pytest.mark.parametrize("H,B", [(5, 16), (5, 32), (5, 4096), (3, 16), (3, 4096)])
def test_mass_sweep_matches_reference_pipeline(engine, tmp_path, H, B):
    """landscape.mass_sweep on the GPU vs the reference's real mass_sweep/apply_base_mass/evaluate_batch/control
    code executed on the oracle physics (strict_reference packing reproduces quirks D2/D3)."""
    g = np.load(GOLD / "mass_sweep.npz")
    paths = []
    for i in range(3):
        rec = {k[len(f"rec{i}_"):]: g[k] for k in g.files if k.startswith(f"rec{i}_")}
        p = tmp_path / f"rec{i}.npz"
        np.savez(p, **rec)
        paths.append(p)
    total, ds = dsmod.load_dataset(paths, H)
    batch = min(B, total, landscape.MAX_SAFE_ENV_BATCH)
    segs = dsmod.pack_segments(dsmod.to_device(ds, engine.device), env_batch=batch, strict_reference=True)
    res = landscape.mass_sweep(engine, segs, g["ref_masses"], g["scales"])
    np.testing.assert_allclose(res, g[f"H{H}_B{B}_costs"], rtol=0, atol=2e-5)
    s_ref = landscape.summarize_landscape(g[f"H{H}_B{B}_costs"], g["scales"], 6.921, 15.019)
    s_gpu = landscape.summarize_landscape(res, g["scales"], 6.921, 15.019)
    assert s_ref.best_idx == s_gpu.best_idx


@pytest.mark.parametrize("B", [16, 4096])
def test_real_reference_env_golden(engine, tmp_path, nominal_model, B):
    """tests/golden/b1_env.npz: costs produced by the reference's REAL `LeggedRobotBase` env (built by its real
    `instantiate_env` on the B200Sim plugin, oracle backend) under its real `mass_sweep` / `apply_base_mass` /
    `evaluate_batch` (tests/test_b1_reference_env.py runs that live in the build container).  Here, on the GPU: (a) the fused
    operator `spi_b200_eval_candidates`, (b) the stepwise CUDA plugin path (B200Sim on `spi_b200_sim_step`, driven chunk by
    chunk like evaluate_batch does: row-0 gains, chunk mask) — both within 2e-5 of the golden."""
    from spi_active_b200.simulator import B200Sim
    from test_simulator_plugin import _config, _replay
    g = np.load(GOLD / "b1_env.npz")
    H = int(g["H"])
    paths = []
    for i in range(3):
        rec = {k[len(f"rec{i}_"):]: g[k] for k in g.files if k.startswith(f"rec{i}_")}
        p = tmp_path / f"rec{i}.npz"
        np.savez(p, **rec)
        paths.append(p)
    total, ds = dsmod.load_dataset(paths, H)
    batch = min(B, total, landscape.MAX_SAFE_ENV_BATCH)
    assert batch == int(g[f"B{B}_batch"])
    segs = dsmod.pack_segments(dsmod.to_device(ds, engine.device), env_batch=batch, strict_reference=True)
    res = landscape.mass_sweep(engine, segs, g["ref_masses"], g["scales"])
    np.testing.assert_allclose(res, g[f"B{B}_costs"], rtol=0, atol=2e-5)
    # (b) stepwise through the plugin on the CUDA backend, one candidate
    sim = B200Sim(config=_config(batch), device=str(engine.device))
    sim.set_headless(True); sim.setup(); sim.setup_terrain("plane"); sim.load_assets()
    sim.create_envs(batch, torch.zeros(batch, 3), torch.tensor([0, 0, 0.34, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], dtype=torch.float32))
    sim.get_dof_limits_properties(); sim.prepare_sim()
    mass = float(g["scales"][2] * g["ref_masses"][0])
    sums, me = np.zeros(3), ds["motion_ends"]
    for start in range(0, total, batch):
        end = min(start + batch, total)
        n = end - start
        chunk = {k: np.concatenate([v[start:end], np.repeat(v[start:start + 1], batch - n, axis=0)]) for k, v in ds.items()}
        per = _replay(sim, chunk, H, nominal_model, base_mass=mass)[:n]
        mask = ~(np.cumsum(me[start:end]) > 0)                                   # scripts/eval.py:279-280
        sums += (per * mask[:, None]).sum(axis=0)
    np.testing.assert_allclose(sums / float((~me).sum()), g[f"B{B}_costs"][2], rtol=0, atol=2e-5)
