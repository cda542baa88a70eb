"""`simulator=b200` plugin (boundary B1): host logic on CPU with an oracle-backed stand-in for the CUDA engine, and
on the GPU the stepwise path (set state -> H x decimation x (torque, simulate) -> errors, i.e. what the reference's
evaluate_batch does through BaseSimulator) against the fused operator on the same inputs."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from spi_active_b200 import dataset as dsmod
from spi_active_b200 import go2_model as gm
from spi_active_b200.simulator import B200Sim

import synth


class OracleBackend:
    """Same two-method shape as RolloutEngine (sim_step, close), physics by the CPU oracle — test infrastructure."""

    def __init__(self, blob):
        self.blob = blob

    def sim_step(self, state, torques, n_steps=1, params=None, param_names=(), flags=0, foot_force=None, ext_wrench=None):
        from oracle import oracle as orc
        ids = [gm.PARAM_IDS[n] for n in param_names]
        new, ff = orc.sim_step(self.blob, state.numpy().astype(np.float64), torques.numpy().astype(np.float64), n_steps,
                               params=None if params is None else params.numpy(), param_ids=ids, flags=flags,
                               return_foot_force=True, ext_wrench=None if ext_wrench is None else ext_wrench.numpy())
        state.copy_(torch.from_numpy(new.astype(np.float32)))
        if foot_force is not None:
            foot_force.copy_(torch.from_numpy(ff.astype(np.float32)))
        return state

    def body_states(self, state, out=None):
        from oracle import oracle as orc
        bs = torch.from_numpy(orc.body_states(self.blob, state.numpy()).astype(np.float32))
        if out is None:
            return bs
        out.copy_(bs)
        return out

    def close(self):
        pass


def _config(num_envs, params_dict=None, body_spheres=False):
    m = gm.go2_nominal()
    b200 = SimpleNamespace(contact=SimpleNamespace(body_spheres=True)) if body_spheres else None
    return SimpleNamespace(
        num_envs=num_envs, headless=True, params_dict=params_dict,
        simulator=SimpleNamespace(config=SimpleNamespace(sim=SimpleNamespace(fps=200, control_decimation=4, substeps=1),
                                                         b200=b200)),
        robot=SimpleNamespace(dof_names=list(gm.DOF_NAMES), body_names=list(gm.BODY_NAMES),
                              dof_vel_limit_list=list(m.qd_limit), dof_effort_limit_list=list(m.torque_limit)),
        rewards=SimpleNamespace(reward_limit=SimpleNamespace(soft_dof_pos_limit=0.9)),
        termination_scales=SimpleNamespace(termination_close_to_dof_pos_limit=0.98),
        terrain=SimpleNamespace(mesh_type="plane"))


def _make_sim(num_envs, device, backend=None, params_dict=None, body_spheres=False):
    sim = B200Sim(config=_config(num_envs, params_dict, body_spheres), device=device, backend=backend)
    sim.set_headless(True)
    sim.setup()
    sim.setup_terrain("plane")
    sim.load_assets()
    init = torch.tensor([0, 0, 0.34, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], dtype=torch.float32)
    sim.create_envs(num_envs, torch.zeros(num_envs, 3), init)
    limits = sim.get_dof_limits_properties()
    sim.prepare_sim()
    return sim, limits


def _replay(sim, ds, H, model, base_mass=None):
    """evaluate_batch's body for one chunk (scripts/eval.py:243-296) + LeggedRobotBase.step's control path
    (legged_robot_base.py:185-209, 545, 557), written against the BaseSimulator surface only."""
    N, dev = sim.num_envs, sim.all_root_states.device
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev)
    if base_mass is not None:   # apply_base_mass (scripts/eval.py:204-214) through the gym shim
        ref = model.body_masses_isaac_order()
        for env_ptr, actor in zip(sim.envs, sim.robot_handles):
            props = sim.gym.get_actor_rigid_body_properties(env_ptr, actor)
            for idx, prop in enumerate(props):
                prop.mass = float(ref[idx]) if idx else float(base_mass)
            sim.gym.set_actor_rigid_body_properties(env_ptr, actor, props, recomputeInertia=True)
        sim.gym.refresh_mass_matrix_tensors(sim.sim)
    env_ids = torch.arange(N, device=dev)
    root = sim.robot_root_states.clone()
    root[:, 0:3], root[:, 3:7] = t(ds["init_base_pos"]), t(ds["init_base_ori"])
    root[:, 7:10], root[:, 10:13] = t(ds["init_base_lin_vel"]), t(ds["init_base_ang_vel"])
    sim.set_actor_root_state_tensor(env_ids, root)
    dof = sim.dof_state.view(N, 12, 2)
    dof[:, :, 0], dof[:, :, 1] = t(ds["init_joint_pos"]), t(ds["init_joint_vel"])     # in-place view write (eval.py:261-266)
    sim.set_dof_state_tensor(env_ids, sim.dof_state)
    sim.refresh_sim_tensors()
    kp, kd = t(ds["pd_gain_kp"][0]), t(ds["pd_gain_kd"][0])
    qdef, tlim = t(model.q_default), t(model.torque_limit)
    acts = t(ds["action_sequences"])
    for k in range(H):
        a = torch.clip(acts[:, k], -model.action_clip, model.action_clip)
        for _ in range(model.control_decimation):
            tau = kp * (a * model.action_scale + qdef - sim.dof_pos) - kd * sim.dof_vel
            sim.apply_torques_at_dof(torch.clip(tau, -tlim, tlim))
            sim.simulate_at_each_physics_step()
        sim.refresh_sim_tensors()
    e_pos = torch.norm(sim.robot_root_states[:, 0:3] - t(ds["target_base_pos"]), dim=1)
    e_quat = torch.norm(sim.robot_root_states[:, 3:7] - t(ds["target_base_ori"]), dim=1)
    e_joint = torch.norm(sim.dof_pos - t(ds["target_joint_pos"]), dim=1)
    return torch.stack([e_pos, e_quat, e_joint], dim=1).cpu().numpy()


def test_plugin_surface_and_limits(blob, nominal_model):
    sim, (pos_lim, vel_lim, tau_lim) = _make_sim(5, "cpu", OracleBackend(blob))
    assert (sim.num_dof, sim.num_bodies) == (12, 19) and sim.sim_dt == 0.005 and sim.device == "cpu"
    assert sim.find_rigid_body_indice("FR_foot") == 8 and sim.find_rigid_body_indice("RR_foot") == 18
    np.testing.assert_allclose(tau_lim.numpy(), nominal_model.torque_limit)
    mid = (np.asarray(nominal_model.q_lower) + np.asarray(nominal_model.q_upper)) / 2
    rng = np.asarray(nominal_model.q_upper) - np.asarray(nominal_model.q_lower)
    np.testing.assert_allclose(pos_lim[:, 0].numpy(), mid - 0.45 * rng, rtol=1e-6)        # soft limits (isaacgym.py:343-346)
    np.testing.assert_allclose(sim.dof_pos_limits_termination[:, 1].numpy(), mid + 0.49 * rng, rtol=1e-6)
    # tensor attributes share storage the way the env layer relies on (eval.py:261-266)
    sim.dof_state.view(5, 12, 2)[:, :, 0] = 0.25
    assert float(sim.dof_pos[3, 7]) == 0.25 and sim.base_quat.data_ptr() == sim.robot_root_states[:, 3:7].data_ptr()
    assert sim.contact_forces.shape == (5, 19, 3) and sim._rigid_body_pos.shape == (5, 19, 3)
    props = sim.gym.get_actor_rigid_body_properties(sim.envs[0], sim.robot_handles[0])
    assert len(props) == 19 and abs(sum(p.mass for p in props) - 15.019) < 1e-3 and abs(props[0].com.x - 0.021112) < 1e-7
    with pytest.raises(NotImplementedError):
        sim.setup_viewer()
    with pytest.raises(NotImplementedError):
        sim.setup_terrain("trimesh")
    with pytest.raises(AssertionError):
        bad = B200Sim(config=_config(1), device="cpu", backend=OracleBackend(blob)); bad.robot_config.dof_names[0] = "x"; bad.load_assets()


def test_contact_block_of_the_yaml_reaches_the_model(blob):
    """config/simulator/b200.yaml `b200.contact.*` -> ContactParams of the model the engine is built from."""
    cfg = _config(2)
    cfg.simulator.config.b200 = SimpleNamespace(inertia_keep=True, strict_inertiay=False,
                                                contact=SimpleNamespace(kn=12345.0, cn=0.25, mu=0.8, dt=20.0, nsub=4))
    sim = B200Sim(config=cfg, device="cpu", backend=OracleBackend(blob))
    c = sim.model.contact
    assert (c.kn, c.cn, c.mu, c.dt, c.nsub) == (12345.0, 0.25, 0.8, 20.0, 4) and sim.inertia_keep
    b = gm.build_model_blob(sim.model)
    assert b[gm.BLOB["CONTACT_KN"]] == np.float32(12345.0) and b[gm.BLOB["NSUB"]] == 4.0
    import yaml
    from pathlib import Path
    y = yaml.safe_load((Path(__file__).resolve().parent.parent / "config" / "simulator" / "b200.yaml").read_text())
    d = gm.ContactParams()                      # the shipped yaml states the defaults
    assert y["simulator"]["config"]["b200"]["contact"] == dict(kn=d.kn, cn=d.cn, mu=d.mu, dt=d.dt, nsub=d.nsub,
                                                               body_spheres=False)


def test_params_dict_overrides(blob):
    """isaacgym_active_sysid.py:39-94: per-env values; `inertiaiy` (sic) accepted; strict flag drops inertiay."""
    pd = {"mass": {"body_name": "base", "value": [9.39, 9.49]}, "comx": {"body_name": "base", "value": [0.0, 0.1]},
          "inertiaiy": {"body_name": "base", "value": [0.005, 0.105]}, "motor_model_hip_a": {"value": [20.0, 20.1]}}
    sim, _ = _make_sim(2, "cpu", OracleBackend(blob), params_dict=pd)
    np.testing.assert_allclose(sim._params_host[:, 0], [9.39, 9.49])
    np.testing.assert_allclose(sim._params_host[:, 1], [0.0, 0.1])
    np.testing.assert_allclose(sim._params_host[:, 5], [0.005, 0.105])


def test_stepwise_replay_on_oracle_backend_matches_fused_oracle(oracle_lib, blob, nominal_model):
    S, ds = synth.dataset("sine", 5, steps=30)
    sim, _ = _make_sim(S, "cpu", OracleBackend(blob))
    per = _replay(sim, ds, 5, nominal_model, base_mass=8.5)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    _, _, ref = oracle_lib.eval_candidates(blob, np.array([[8.5]], np.float32), [0], init, act, tgt, gains, mask,
                                           return_per_seg=True)
    np.testing.assert_allclose(per, ref[0], atol=2e-6)     # fp32 tensors between steps vs fp64 throughout
    assert float(sim.contact_forces[:, [4, 8, 14, 18], 2].sum()) > 0


@pytest.mark.gpu
def test_stepwise_replay_on_gpu_matches_fused_operator(engine, nominal_model):
    S, ds = synth.dataset("jump", 5, steps=60)
    sim, _ = _make_sim(S, str(engine.device))
    per = _replay(sim, ds, 5, nominal_model, base_mass=8.5)
    segs = dsmod.pack_segments(dsmod.to_device(ds, engine.device))
    _, fused = engine.evaluate_candidates(torch.tensor([[8.5]]), ["mass"], segs, return_per_seg=True)
    np.testing.assert_allclose(per, fused[0].cpu().numpy(), atol=2e-5)
    assert float(sim.contact_forces[:, [4, 8, 14, 18], 2].abs().sum()) > 0


# ---- rigid-body state tensor (forward kinematics) ------------------------------------------------
def _random_states(n, seed, model):
    rng = np.random.default_rng(seed)
    s = np.zeros((n, 37))
    s[:, 0:3] = rng.uniform(-1, 1, (n, 3)); s[:, 2] = rng.uniform(0.2, 0.5, n)
    q = rng.standard_normal((n, 4)); s[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    s[:, 7:13] = rng.uniform(-1, 1, (n, 6))
    s[:, 13:25] = np.asarray(model.q_default) + rng.uniform(-0.4, 0.4, (n, 12))
    s[:, 25:37] = rng.uniform(-3, 3, (n, 12))
    return s


def test_body_states_oracle_kinematics(oracle_lib, blob, nominal_model):
    """The FK restatement against independent facts: URDF geometry at the zero pose, the foot height the contact model
    uses, and velocities as the time derivative of positions / orientations along a free-flight trajectory."""
    from scipy.spatial.transform import Rotation
    m = nominal_model
    z = np.zeros((1, 37)); z[0, 2] = 0.4; z[0, 6] = 1.0
    bs = oracle_lib.body_states(blob, z)[0]
    names = gm.BODY_NAMES
    np.testing.assert_allclose(bs[names.index("FL_hip"), :3], [0.1934, 0.0465, 0.4], atol=1e-7)
    np.testing.assert_allclose(bs[names.index("RR_thigh"), :3], [-0.1934, -0.0465 - 0.0955, 0.4], atol=1e-7)
    np.testing.assert_allclose(bs[names.index("FR_foot"), :3], [0.1934, -0.0465 - 0.0955, 0.4 - 0.426], atol=1e-7)
    np.testing.assert_allclose(bs[names.index("Head_lower"), :3], [0.293, 0.0, 0.34], atol=1e-7)
    assert np.allclose(bs[:, 3:7], [0, 0, 0, 1]) and np.allclose(bs[:, 7:13], 0)
    # velocities = d/dt of the kinematics: integrate the state itself (constant twist / joint rates) over a small dt
    s0 = _random_states(6, 1, m)
    dt = 1e-6
    s1 = s0.copy()
    s1[:, 0:3] += dt * s0[:, 7:10]
    rot = Rotation.from_rotvec(dt * s0[:, 10:13]) * Rotation.from_quat(s0[:, 3:7])     # world-frame angular velocity
    s1[:, 3:7] = rot.as_quat()
    s1[:, 13:25] += dt * s0[:, 25:37]
    b0, b1 = oracle_lib.body_states(blob, s0), oracle_lib.body_states(blob, s1)
    np.testing.assert_allclose((b1[..., 0:3] - b0[..., 0:3]) / dt, b0[..., 7:10], atol=2e-5)
    dR = Rotation.from_quat(b1[..., 3:7].reshape(-1, 4)) * Rotation.from_quat(b0[..., 3:7].reshape(-1, 4)).inv()
    np.testing.assert_allclose(dR.as_rotvec().reshape(6, 19, 3) / dt, b0[..., 10:13], atol=2e-5)
    # the foot sphere the contact model uses sits at foot origin + R (-0.002, 0, 0)
    f = names.index("RL_foot")
    c = b0[:, f, :3] + Rotation.from_quat(b0[:, f, 3:7]).apply(np.asarray(m.foot_sphere_offset))
    calf = names.index("RL_calf")
    c2 = b0[:, calf, :3] + Rotation.from_quat(b0[:, calf, 3:7]).apply(blob[214 + 6:214 + 9].astype(np.float64))
    np.testing.assert_allclose(c, c2, atol=1e-7)


def test_plugin_exposes_all_rigid_bodies(blob, nominal_model):
    sim, _ = _make_sim(3, "cpu", backend=OracleBackend(blob))
    sim.refresh_sim_tensors()
    assert sim._rigid_body_pos.shape == (3, 19, 3) and sim._rigid_body_rot.shape == (3, 19, 4)
    feet = [sim.find_rigid_body_indice(f"{leg}_foot") for leg in gm.LEGS]
    assert feet == [4, 8, 14, 18]
    z = sim._rigid_body_pos[:, feet, 2]
    assert float(z.min()) > 0.0 and float(z.max()) < 0.34          # feet hang below the base at the default pose
    assert torch.allclose(sim._rigid_body_pos[:, 0], sim.robot_root_states[:, 0:3])


@pytest.mark.gpu
def test_body_states_kernel_matches_oracle(engine, oracle_lib, blob, nominal_model):
    s = _random_states(257, 2, nominal_model)
    out = engine.body_states(torch.from_numpy(s.astype(np.float32)).to(engine.device)).cpu().numpy()
    ref = oracle_lib.body_states(blob, s.astype(np.float32))
    np.testing.assert_allclose(out[..., 0:3], ref[..., 0:3], atol=2e-6)
    sign = np.sign((out[..., 3:7] * ref[..., 3:7]).sum(-1, keepdims=True))
    assert (sign > 0).all()                                         # same hemisphere: composed, not re-extracted
    np.testing.assert_allclose(out[..., 3:7], ref[..., 3:7], atol=2e-6)
    np.testing.assert_allclose(out[..., 7:13], ref[..., 7:13], atol=2e-5)


# ---- external forces: IsaacGym.apply_rigid_body_force_at_pos_tensor (isaacgym.py:609-613) ----------------------------------
def _free_flight_sim(num_envs, device, backend, seed=3, vel_scale=1.0):
    from test_oracle_physics import _rand_state
    sim, _ = _make_sim(num_envs, device, backend)
    rng = np.random.default_rng(seed)
    m = gm.go2_nominal()
    st = np.stack([_rand_state(rng, m, height=5.0) for _ in range(num_envs)]).astype(np.float32)
    st[:, 7:13] *= vel_scale; st[:, 25:37] *= vel_scale
    sim._state.copy_(torch.from_numpy(st).to(sim._state.device))
    sim.refresh_sim_tensors()
    return sim, st


def test_external_force_changes_momentum_by_the_impulse(blob, nominal_model):
    """A world-frame force F at a world point p on any of the 19 Isaac bodies, for one physics step: total linear momentum
    changes by F dt and angular momentum about the world origin by (p x F) dt relative to the unforced step (free flight, zero
    joint torques) — independent of which body carries it, which the [torque; force]-per-moving-body conversion must preserve."""
    from test_oracle_physics import _momentum
    dt = nominal_model.dt
    cases = [("base", [30.0, -20.0, 50.0]), ("RL_thigh", [0.0, 40.0, 10.0]), ("FR_foot", [15.0, 5.0, 60.0]),
             ("Head_lower", [-25.0, 0.0, 5.0]), ("FL_calf", [10.0, -30.0, 0.0])]
    N = len(cases)
    # slow states: the wrench is held in the BODY frame over the step's two sub-steps, so a body turning at w rad/s sees the
    # force direction drift by w * dt / 2 (1 - 2 % for a swinging calf); at 0.1x the rates the mapping itself is checked tightly
    ref, st0 = _free_flight_sim(N, "cpu", OracleBackend(blob), vel_scale=0.1)
    ref.apply_torques_at_dof(torch.zeros(N, 12)); ref.simulate_at_each_physics_step()
    sim, _ = _free_flight_sim(N, "cpu", OracleBackend(blob), vel_scale=0.1)
    force, pos = torch.zeros(N, 19, 3), torch.zeros(N, 19, 3)
    for e, (name, F) in enumerate(cases):
        b = sim.find_rigid_body_indice(name)
        force[e, b] = torch.tensor(F)
        pos[e, b] = sim._rigid_body_pos[e, b] + torch.tensor([0.03, -0.02, 0.05])      # off the link origin: a torque arm
    sim.apply_torques_at_dof(torch.zeros(N, 12))
    sim.apply_rigid_body_force_at_pos_tensor(force, pos)
    sim.simulate_at_each_physics_step()
    for e, (name, F) in enumerate(cases):
        b = sim.find_rigid_body_indice(name)
        P1, L1, _ = _momentum(nominal_model, sim._state[e].numpy().astype(np.float64))
        P0, L0, _ = _momentum(nominal_model, ref._state[e].numpy().astype(np.float64))
        F = np.asarray(F, np.float64)
        # (a 60 N push on a free 0.2 kg calf spins it up to ~8 rad/s within the first sub-step, so even from rest the held
        # body-frame direction drifts by ~1 % in the second one: tolerance 2 % of the impulse)
        np.testing.assert_allclose(P1 - P0, F * dt, rtol=2e-2, atol=5e-3, err_msg=name)
        np.testing.assert_allclose(L1 - L0, np.cross(pos[e, b].numpy().astype(np.float64), F) * dt, rtol=2e-2, atol=2e-2,
                                   err_msg=name)
    # the force is consumed by that step: the next one is unforced again
    a = sim._state.clone(); sim.simulate_at_each_physics_step()
    r = ref._state.clone(); ref._state.copy_(a); ref.simulate_at_each_physics_step()
    np.testing.assert_allclose(sim._state.numpy(), ref._state.numpy(), atol=1e-6)
    assert not np.allclose(a.numpy(), r.numpy(), atol=1e-4)


@pytest.mark.gpu
def test_external_force_kernel_matches_oracle(engine, oracle_lib, blob, nominal_model):
    """spi_b200_sim_step_ext against the oracle for random wrenches on all 13 moving bodies, and the plugin's conversion of
    world (force, position) pairs on the GPU against the same plugin on the oracle backend."""
    N = 33
    rng = np.random.default_rng(8)
    cpu, st0 = _free_flight_sim(N, "cpu", OracleBackend(blob), seed=5)
    gpu, _ = _free_flight_sim(N, str(engine.device), None, seed=5)
    ext = (rng.standard_normal((N, 13, 6)) * np.array([2, 2, 2, 30, 30, 30])).astype(np.float32)
    tau = rng.uniform(-5, 5, (N, 12)).astype(np.float32)
    s = torch.from_numpy(st0.copy()).to(engine.device)
    engine.sim_step(s, torch.from_numpy(tau), 3, ext_wrench=torch.from_numpy(ext))
    ref = oracle_lib.sim_step(blob, st0.astype(np.float64), tau.astype(np.float64), 3, ext_wrench=ext)
    np.testing.assert_allclose(s.cpu().numpy()[:, :7], ref[:, :7], atol=2e-5)
    np.testing.assert_allclose(s.cpu().numpy()[:, 13:25], ref[:, 13:25], atol=2e-5)
    np.testing.assert_allclose(s.cpu().numpy()[:, 7:13], ref[:, 7:13], atol=1e-3)
    np.testing.assert_allclose(s.cpu().numpy()[:, 25:37], ref[:, 25:37], atol=5e-3)
    force = torch.from_numpy(rng.standard_normal((N, 19, 3)).astype(np.float32) * 20)
    pos = cpu._rigid_body_pos.clone() + 0.05 * torch.from_numpy(rng.standard_normal((N, 19, 3)).astype(np.float32))
    for sim in (cpu, gpu):
        dev = sim._state.device
        sim.apply_torques_at_dof(torch.from_numpy(tau).to(dev))
        sim.apply_rigid_body_force_at_pos_tensor(force.to(dev), pos.to(dev))
        sim.simulate_at_each_physics_step()
    np.testing.assert_allclose(gpu._state.cpu().numpy()[:, :7], cpu._state.numpy()[:, :7], atol=2e-5)
    np.testing.assert_allclose(gpu._state.cpu().numpy()[:, 13:25], cpu._state.numpy()[:, 13:25], atol=2e-5)
    np.testing.assert_allclose(gpu._state.cpu().numpy()[:, 7:13], cpu._state.numpy()[:, 7:13], atol=1e-3)


# ---- non-foot contacts (b200.contact.body_spheres; isaacgym.py:577 contact_forces[N, 19, 3]) ----------------------------------
def _belly_scenario(sim, steps):
    """Legs folded over the back (feet above the trunk), trunk dropped from 10 cm: the robot can only land on its trunk
    corners, head and hips.  PD torques hold the folded pose."""
    N, dev = sim.num_envs, sim.all_root_states.device
    env_ids = torch.arange(N, device=dev)
    root = sim.robot_root_states.clone()
    root[:, 0:3] = torch.tensor([0.0, 0.0, 0.16], device=dev)
    root[:, 3:7] = torch.tensor([0.0, 0.0, 0.0, 1.0], device=dev)
    root[:, 7:13] = 0.0
    sim.set_actor_root_state_tensor(env_ids, root)
    q_fold = torch.tensor([0.0, 2.6, -0.9] * 4, dtype=torch.float32, device=dev)
    dof = sim.dof_state.view(N, 12, 2)
    dof[:, :, 0], dof[:, :, 1] = q_fold, 0.0
    sim.set_dof_state_tensor(env_ids, sim.dof_state)
    sim.refresh_sim_tensors()
    for _ in range(steps):
        tau = 30.0 * (q_fold - sim.dof_pos) - 0.8 * sim.dof_vel
        sim.apply_torques_at_dof(torch.clip(tau, -23.7, 23.7))
        sim.simulate_at_each_physics_step()
    sim.refresh_sim_tensors()


def test_body_spheres_carry_the_robot_on_its_belly(blob, nominal_model):
    """With b200.contact.body_spheres the trunk / head / hip shapes of the URDF collide with the plane: a robot dropped with
    its legs folded away comes to rest on them, they carry its weight, and contact_forces reports them on the right bodies;
    without the option (default, = the fused operator's model) the same robot falls through the plane."""
    sim, _ = _make_sim(2, "cpu", OracleBackend(blob), body_spheres=True)
    _belly_scenario(sim, 400)                                   # 2 s
    weight = float(np.sum(nominal_model.body_masses_isaac_order())) * 9.81
    cf = sim.contact_forces[0]
    feet = [sim.find_rigid_body_indice(f"{leg}_foot") for leg in gm.LEGS]
    assert torch.all(cf[feet] == 0)                             # the feet are in the air
    assert abs(float(cf[:, 2].sum()) - weight) < 0.03 * weight  # at rest: the body contacts carry the weight
    assert float(cf[sim.find_rigid_body_indice("base"), 2]) > 0 or float(cf[sim.find_rigid_body_indice("Head_lower"), 2]) > 0
    z = float(sim.robot_root_states[0, 2])
    assert 0.04 < z < 0.12, z                                   # lying on the trunk, not fallen through
    assert float(sim.robot_root_states[0, 7:13].abs().max()) < 0.05
    # default: feet only -> nothing stops the trunk
    sim0, _ = _make_sim(1, "cpu", OracleBackend(blob))
    _belly_scenario(sim0, 100)
    assert float(sim0.robot_root_states[0, 2]) < 0.0
    non_feet = [i for i in range(19) if i not in feet]
    assert torch.all(sim0.contact_forces[:, non_feet] == 0)


@pytest.mark.gpu
def test_body_spheres_on_gpu_match_the_oracle_backend(engine, blob):
    """The same drop through the CUDA engine (spi_b200_body_states + spi_b200_sim_step_ext) and through the oracle backend."""
    dev = str(engine.device)
    sim_g, _ = _make_sim(3, dev, None, body_spheres=True)
    sim_c, _ = _make_sim(3, "cpu", OracleBackend(blob), body_spheres=True)
    _belly_scenario(sim_g, 60)
    _belly_scenario(sim_c, 60)
    np.testing.assert_allclose(sim_g.robot_root_states.cpu().numpy(), sim_c.robot_root_states.numpy(), atol=2e-3)
    np.testing.assert_allclose(sim_g.dof_pos.cpu().numpy(), sim_c.dof_pos.numpy(), atol=2e-3)
    scale = float(sim_c.contact_forces.abs().max())
    np.testing.assert_allclose(sim_g.contact_forces.cpu().numpy(), sim_c.contact_forces.numpy(), atol=0.02 * scale)
