"""Host-side logic that needs no GPU: candidate sharding + cost all-gather over gloo (world_size 2), the CEM
refit rule, the built-in TPE used when optuna is absent, result-file schemas, recorders."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from spi_active_b200 import cem, landscape, recorders
from spi_active_b200 import go2_model as gm

ROOT = Path(__file__).resolve().parent.parent


def test_shard_range_partitions_exactly():
    for total, world in ((16384, 8), (4096, 1), (12, 4)):
        spans = [cem.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert len({b - a for a, b in spans}) == 1
    with pytest.raises(ValueError):
        cem.shard_range(10, 0, 4)


def _fake_cost(params):
    """A deterministic stand-in for the rollout cost: distance to a known optimum."""
    target = np.array([7.0, 0.02, 0.0, -0.005], dtype=np.float64)
    return ((params.astype(np.float64) - target) ** 2 * np.array([1.0, 50.0, 50.0, 50.0])).sum(axis=1).astype(np.float32)


def _gloo_worker(rank, world, port, C, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)          # same population on every rank (the device RNG is keyed by global index)
    mean, std = np.array([9.0, 0.0, 0.0, 0.0]), np.array([2.0, 0.05, 0.05, 0.05])
    for it in range(6):
        params = (mean + rng.standard_normal((C, 4)) * std).astype(np.float32)
        c0, c1 = cem.shard_range(C, rank, world)
        local = torch.from_numpy(_fake_cost(params[c0:c1]))
        total = cem.gather_costs(local, world).numpy()
        mean, std, best, best_cost = cem.cem_refit_numpy(params, total, max(2, C // 10), 0.7, mean, std)
    np.save(Path(out_dir) / f"rank{rank}.npy", np.concatenate([mean, std, best, [best_cost]]))
    dist.destroy_process_group()


def test_sharded_cem_over_gloo_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    C, world = 256, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_gloo_worker, args=(world, port, C, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    np.testing.assert_array_equal(r0, r1)            # every rank ends with the same distribution
    # single-process run of the same loop
    rng = np.random.default_rng(0)
    mean, std = np.array([9.0, 0.0, 0.0, 0.0]), np.array([2.0, 0.05, 0.05, 0.05])
    for it in range(6):
        params = (mean + rng.standard_normal((C, 4)) * std).astype(np.float32)
        mean, std, best, best_cost = cem.cem_refit_numpy(params, _fake_cost(params), max(2, C // 10), 0.7, mean, std)
    np.testing.assert_allclose(r0, np.concatenate([mean, std, best, [best_cost]]), rtol=1e-12)
    assert abs(mean[0] - 7.0) < 0.3                 # and it converges towards the optimum


def test_cem_refit_rule_handles_nonfinite_and_ties():
    params = np.arange(12, dtype=np.float32).reshape(6, 2)
    cost = np.array([3.0, np.nan, 1.0, 1.0, np.inf, 2.0], dtype=np.float32)
    mean, std, best, bc = cem.cem_refit_numpy(params, cost, 3, 1.0, np.zeros(2), np.ones(2))
    np.testing.assert_allclose(mean, params[[2, 3, 5]].mean(0))     # stable order: index 2 before 3; nan/inf last
    np.testing.assert_array_equal(best, params[2]); assert bc == 1.0
    assert (std > 0).all()


def test_default_cem_config_is_physically_valid(nominal_model):
    cfg = cem.default_full_config(nominal_model)
    assert cfg.names == cem.FULL_PARAM_NAMES and len(cfg.mean) == 10
    assert (np.asarray(cfg.lo) < np.asarray(cfg.mean)).all() and (np.asarray(cfg.mean) < np.asarray(cfg.hi)).all()
    assert np.asarray(cfg.lo)[[0, 4, 5, 6, 7, 8, 9]].min() > 0     # mass, inertia, motor a stay positive


def test_tpe_finds_minimum_of_a_bowl_and_replays_enqueued_trial():
    f = lambda x: (x - 1.013) ** 2 + 0.028
    best_x, best_v, trials = landscape.optimize_mass(f, n_trials=50, seed=0)
    assert trials[0][0] == landscape.INITIAL_MASS_SCALE == 3.0            # quirk D7: evaluated although out of range
    assert all(landscape.MASS_SCALE_MIN <= x <= landscape.MASS_SCALE_MAX for x, _ in trials[1:])
    assert len(trials) == 50 and abs(best_x - 1.013) < 0.03
    # determinism under the seed
    assert landscape.optimize_mass(f, n_trials=50, seed=0)[0] == best_x


def test_result_file_schemas(tmp_path):
    scales = np.linspace(landscape.MASS_SCALE_MIN, landscape.MASS_SCALE_MAX, landscape.MASS_SAMPLES)
    costs = np.stack([(scales - 1.0) ** 2 * 0.01 + 0.001, (scales - 1.0) ** 2 * 0.005 + 0.002,
                      (scales - 1.0) ** 2 * 0.05 + 0.04], axis=1).astype(np.float32)
    s = landscape.write_landscape_results(tmp_path / "landscape_results.txt", "all", 5, costs, scales, 6.921, 15.019)
    lines = (tmp_path / "landscape_results.txt").read_text().splitlines()
    assert lines[0] == "Mass Landscape Results" and lines[1] == "=" * 60
    assert lines[11] == "scale,base_mass_kg,total_mass_kg,base_pos,base_quat,joint_pos,total_cost"
    assert len(lines) == 12 + landscape.MASS_SAMPLES
    assert abs(s.best_scale - scales[np.argmin(np.abs(scales - 1.0))]) < 1e-9
    assert abs(s.cost_percentages.sum() - 100) < 1e-9
    landscape.write_optimization_results(tmp_path / "optimization_results.txt", "all", 5, 50, 6.921, 7.006, 0.028378)
    txt = (tmp_path / "optimization_results.txt").read_text()
    assert "Mass Optimization Results (Optuna)" in txt and "Optimal base mass: 7.006 kg" in txt and "Best cost: 0.028378" in txt


def test_load_config_paths():
    p = landscape.load_config("all", Path("/data"))
    assert [x.name for x in p] == ["go2_jump_data.npz", "go2_sine_data.npz", "go2_stand_data.npz", "go2_walk_data.npz"]
    assert p[0].parent == Path("/data/spigym/data/sysid_bag")


def test_action_laws_match_reference_recorders():
    """scripts/data/sine.py:41-45, jump.py:17-18,41-45: thigh = +A sin(2 pi f t), calf = -same, dt 0.02."""
    for name, amp, f in (("sine", 0.8, 1.0), ("jump", 1.2, 1.5)):
        a = recorders.action_law(name)
        t = np.arange(250) * 0.02
        ph = amp * np.sin(2 * np.pi * f * t)
        assert a.shape == (250, 12)
        np.testing.assert_allclose(a[:, [1, 4, 7, 10]], np.repeat(ph[:, None], 4, 1), atol=1e-6)
        np.testing.assert_allclose(a[:, [2, 5, 8, 11]], -np.repeat(ph[:, None], 4, 1), atol=1e-6)
        assert np.abs(a[:, [0, 3, 6, 9]]).max() == 0
    assert np.abs(recorders.action_law("stand")).max() == 0 and recorders.action_law("walk").shape == (1000, 12)


def test_recording_schema_roundtrip(tmp_path, oracle_lib, blob, nominal_model):
    import synth
    from spi_active_b200 import dataset as dsmod
    rec = synth.recording("stand", 20)
    dsmod.save_recording(tmp_path / "go2_stand_data.npz", {k: v for k, v in rec.items() if not k.startswith("pd_")},
                         recorders.CONTROL_DT, rec["pd_gain_kp"], rec["pd_gain_kd"])
    with np.load(tmp_path / "go2_stand_data.npz") as z:
        for k in ("joint_positions", "joint_velocities", "joint_torques", "actions", "base_positions", "base_orientations",
                  "base_linear_velocities", "base_angular_velocities", "timestamps", "sim_duration", "data_frequency",
                  "robot_type", "pd_gain_kp", "pd_gain_kd"):      # scripts/data/common.py:83-100
            assert k in z.files, k
        assert int(z["data_frequency"]) == 50 and str(z["robot_type"]) == "go2"
    total, ds = dsmod.load_dataset(tmp_path / "go2_stand_data.npz", 5)
    assert total == 15 and ds["motion_ends"].sum() == 1 and ds["motion_ends"][-1]


def test_record_data_cli_writes_files_the_loader_reads(tmp_path, oracle_lib, blob, nominal_model, monkeypatch):
    """scripts/record_data.py (counterpart of the reference's scripts/data/*.py): file names and directory layout are what
    scripts/config/*.yaml / landscape.load_config resolve, the schema is what load_dataset reads.  The oracle stands in for
    the engine here (no GPU)."""
    import importlib.util
    from oracle import oracle as orc
    from spi_active_b200 import dataset as dsmod, go2_model as gm, landscape
    spec = importlib.util.spec_from_file_location("record_data", ROOT / "scripts" / "record_data.py")
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    monkeypatch.setitem(recorders.DURATION_STEPS, "stand", 30)
    monkeypatch.setitem(recorders.DURATION_STEPS, "sine", 30)

    def rollout_fn(init, actions):
        st = orc.rollout_states(blob, np.array([[nominal_model.base.mass]], np.float32), [gm.PARAM_IDS["mass"]], init[None],
                                actions[None])
        return st[0, 0]
    out_dir = tmp_path / landscape.DATA_SUBDIR
    out_dir.mkdir(parents=True)
    for n in ("stand", "sine"):
        p = mod.record_to_file(n, rollout_fn, nominal_model, out_dir)
        assert p.name == f"go2_{n}_data.npz"
    assert landscape.load_config("stand", tmp_path) == [out_dir / "go2_stand_data.npz"]
    total, ds = dsmod.load_dataset([out_dir / "go2_stand_data.npz", out_dir / "go2_sine_data.npz"], 5)
    assert total == 50 and ds["motion_ends"].sum() == 2
    np.testing.assert_allclose(ds["pd_gain_kp"][0], nominal_model.kp[0]); np.testing.assert_allclose(ds["pd_gain_kd"][0], nominal_model.kd[0])


def test_bench_reference_arm_runs_without_gpu():
    """`bench.py --impl reference` is CPU-only and prints the contract's JSON line."""
    import json
    import subprocess
    env = dict(os.environ, SPI_BENCH_REFERENCE_BUDGET_S="3")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "candidate-env steps/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
