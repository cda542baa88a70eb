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
