"""Boundary B1 for real: the reference's OWN env classes on top of the `simulator=b200` plugin.

`scripts/eval.py:313-321 instantiate_env` runs unmodified: load_env_config (hydra.compose stood in by tests/mini_hydra over
the reference's YAML groups + this repo's config/env/go2_test.yaml, config/obs/go2_test.yaml, config/simulator/b200.yaml),
`LeggedRobotBase(cfg.env.config, device)` -> `BaseTask.__init__` (get_class(simulator._target_), set_headless, setup,
setup_terrain, load_assets, create_envs, get_dof_limits_properties, find_rigid_body_indice, prepare_sim) -> `_init_buffers` ->
`set_is_evaluating` -> `reset_all` (random reset, set_*_state_tensor, one full `step` incl. `_post_physics_step`:
refresh_sim_tensors, termination, reward, observations, history).  Then the reference's real `mass_sweep` -> `apply_base_mass`
(through the gym-handle shim) -> `evaluate_batch` drive the plugin, and the three costs must equal the fused operator's
semantics (the oracle's `eval_candidates` on `pack_segments(strict_reference=True)`) to 2e-5.

Runs only where /root/reference exists (the build container); the GPU box checks the same numbers through
tests/golden/b1_env.npz (tests/test_gpu_golden.py).  The simulator class is the product's B200Sim with the CPU oracle behind
its backend interface (tests/b200_oracle_sim.py) — the CUDA backend needs a GPU.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="/root/reference is not present (GPU box)")

TARGET = "b200_oracle_sim.B200SimOracleBackend"
GOLD = Path(__file__).resolve().parent / "golden"
SCALES = (0.6, 1.0, 1.7)


@pytest.fixture(scope="module")
def env16():
    return rh.instantiate_env(16, TARGET)


def test_real_env_constructs_and_steps_on_the_plugin(env16):
    R = rh.import_reference()
    env = env16
    from spi_active_b200.simulator import B200Sim
    assert type(env) is R.LeggedRobotBase and isinstance(env.simulator, B200Sim)
    # every public method of the reference's BaseSimulator (base_simulator.py:6-171) is provided by the plugin (it subclasses
    # the reference class itself when spigym is importable at its import time, a local mirror otherwise)
    for name, member in vars(R.bt.BaseSimulator).items():
        if callable(member) and not name.startswith("__"):
            assert callable(getattr(env.simulator, name, None)), name
    assert env.num_envs == 16 and env.dim_actions == 12 and abs(env.dt - 0.02) < 1e-12
    assert env.obs_buf_dict["actor_obs"].shape == (16, 42) and env.obs_buf_dict["critic_obs"].shape == (16, 45)
    np.testing.assert_allclose(env.p_gains.numpy(), 25.0)
    np.testing.assert_allclose(env.d_gains.numpy(), 0.6)
    assert env.feet_indices.tolist() == [4, 8, 14, 18]
    # three more full steps (incl. _post_physics_step) with random actions: finite, robots stay near their grid origins
    g = torch.Generator().manual_seed(0)
    for _ in range(3):
        obs, rew, reset, extras = env.step({"actions": torch.randn(16, 12, generator=g)})
    assert torch.isfinite(obs["actor_obs"]).all() and int(reset.sum()) == 0
    assert (env.simulator.robot_root_states[:, :3] - env.env_origins).abs().max() < 0.5
    assert int(env.episode_length_buf[0]) == 4
    # the tensors are views: an in-place write to dof_pos is visible in dof_state (scripts/eval.py:261-266 relies on it)
    env.simulator.dof_pos[3, 5] = 0.123
    assert float(env.simulator.dof_state.view(16, 12, 2)[3, 5, 0]) == pytest.approx(0.123)


def _recordings():
    import synth
    recs = []
    for name, steps, kp, kd in (("jump", 40, 25.0, 0.6), ("sine", 33, 22.0, 0.5), ("stand", 21, 28.0, 0.7)):
        r = dict(synth.recording(name, steps))
        r["pd_gain_kp"] = np.full(12, kp, np.float32)
        r["pd_gain_kd"] = np.full(12, kd, np.float32)
        r.update(timestamps=np.arange(steps) * 0.02, sim_duration=steps * 0.02, data_frequency=50, robot_type="go2")
        recs.append(r)
    return recs


def run_reference_sweep(tmp_path, batch_sizes=(16, 4096), H=5, scales=SCALES):
    """-> {B: (batch, costs[len(scales), 3])} from the reference's real mass_sweep through the real env on the plugin."""
    R = rh.import_reference()
    paths = []
    for i, r in enumerate(_recordings()):
        p = Path(tmp_path) / f"rec{i}.npz"
        np.savez(p, **r)
        paths.append(p)
    total, ds_np = R.ev.load_dataset(paths, H)
    out = {}
    for B in batch_sizes:
        batch = min(B, total, R.ml.MAX_SAFE_ENV_BATCH)                  # mass_landscape.py:146
        env = rh.instantiate_env(batch, TARGET)
        dataset = R.ev.to_device(ds_np, env.device)
        ref_masses = R.ev.capture_reference_masses(env.simulator)
        out[B] = (batch, R.ml.mass_sweep(env, dataset, H, ref_masses, np.asarray(scales)))
    return total, ds_np, ref_masses, out


def test_reference_mass_sweep_through_the_plugin_equals_the_fused_operator(tmp_path, oracle_lib, blob):
    from spi_active_b200 import go2_model as gm
    from spi_active_b200.dataset import pack_segments, to_device
    total, ds_np, ref_masses, out = run_reference_sweep(tmp_path)
    np.testing.assert_allclose(ref_masses, gm.go2_nominal().body_masses_isaac_order())
    gold = np.load(GOLD / "b1_env.npz")
    for B, (batch, costs) in out.items():
        segs = pack_segments(to_device(ds_np, "cpu"), env_batch=batch, strict_reference=True)
        ref, st = oracle_lib.eval_candidates(blob, (np.asarray(SCALES) * float(ref_masses[0])).astype(np.float32)[:, None],
                                             [gm.PARAM_IDS["mass"]], segs.seg_init.numpy(), segs.seg_actions.numpy(),
                                             segs.seg_target.numpy(), segs.seg_gains.numpy(), segs.seg_mask.numpy(),
                                             cost_denominator=segs.cost_denominator)
        assert st.sum() == 0
        np.testing.assert_allclose(costs, ref, atol=2e-5, rtol=0, err_msg=f"B={B}")
        np.testing.assert_allclose(costs, gold[f"B{B}_costs"], atol=1e-6, rtol=0)      # the committed golden reproduces
