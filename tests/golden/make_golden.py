"""Generate the committed golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN PYTHON.

Run in the build container (it needs /root/reference; the GPU box and the test-suite never do):

    python tests/golden/make_golden.py

The reference (LeCAR-Lab/SPI-Active) is pure Python around the closed-source Isaac Gym binary, and its
hot-path modules import hydra / omegaconf / optuna / matplotlib / isaacgym, none of which exist here.
Those imports are replaced by inert stubs (MagicMock modules); every function exercised below is then the
reference's real, unmodified code object, called with plain torch tensors:

  torques.npz        LeggedRobotBase._pre_physics_step + _compute_torques
                         (spigym/envs/legged_base_task/legged_robot_base.py:185-198, 529-560)
                     go2_omni_interface._compute_torques, hip x0.5
                         (spigym/envs/locomotion/go2_omni.py:423-465)
                     ActiveSysId_OpenLoop.act2tau_scalar / act2tau_vec3 / act2tau_vec3_tanh
                         (spigym/envs/sysid/active_sysid_openloop.py:356-400)
  windowing.npz      scripts/eval.py:101-171 load_dataset on two small synthetic recordings
  fim.npz            ActiveSysId_OpenLoop._reward_fisher_information_matrix (:402-426)
  cost.npz           scripts/mass_opt.py:62-76 compute_cost; mass_landscape.py COST_COEFF / argmin (:162-171)
  quat.npz           spigym/utils/torch_utils.py quat_rotate_inverse / quat_rotate (:49-92)
  active_obs.npz     the 900-dim actor observation of the active-exploration env over 18 steps: real
                     go2_omni._step_contact_targets (:348-377), every _get_obs_* getter (go2_omni.py:627-657,
                     legged_robot_base.py:801-844), _pre_compute_observations_callback (:261-269),
                     _compute_observations / _post_config_observation_callback (:511-527), helpers.parse_observation,
                     env_utils/history_handler.HistoryHandler
  mass_sweep.npz     scripts/mass_landscape.py:111-128 mass_sweep -> scripts/eval.py:204-214 apply_base_mass
                     -> :217-310 evaluate_batch -> LeggedRobotBase._pre_physics_step/_physics_step/
                     _apply_force_in_physics_step/_compute_torques, all real, driving tests/oracle_sim.OracleSim
                     (our CPU oracle as the physics engine in place of PhysX).  Pins the replay / masking /
                     averaging / PD-gain / control-decimation semantics end to end, including quirks D2
                     (chunk-dependent eval_mask) and D3 (gains of row 0 of each chunk).

  urdf_blob.npz      the 312-float model blob built from the reference's own spigym/data/robots/go2/urdf/go2.urdf by
                     spi_active_b200.go2_model.model_from_urdf (pins the typed-in table of go2_nominal())
  ppo_actor.pt/.npz  a checkpoint in the layout of PPO.save (spigym/agents/ppo/ppo.py:155-166) holding the state dict of the
                     reference's real PPOActor (agents/modules/ppo_modules.py, modules.py:47-63; a small 60-32-16-8-12 ELU
                     instance of config/algo/ppo.yaml:31-39) + observations and its act_inference outputs

  b1_env.npz         the reference's real mass_sweep / apply_base_mass / evaluate_batch driving its real
                     LeggedRobotBase(BaseTask) env built by its real instantiate_env on the product's B200Sim plugin (oracle
                     backend; tests/ref_harness.py, tests/test_b1_reference_env.py): recordings + costs for two chunkings

  commands.npz       ActiveSysId.sample_commands / _sample_constant / _sample_polynomial / _sample_bezier
                     (spigym/agents/sysid/active_sysid.py:259-410) with a stub trial.suggest_float replaying seeded values

What this cannot pin is the rigid-body physics itself: PhysX is not available.
"""
from __future__ import annotations

import importlib
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace
from unittest import mock

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

STUBBED = ["hydra", "hydra.core", "hydra.core.config_store", "hydra.utils", "hydra.core.hydra_config", "omegaconf",
           "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "mpl_toolkits",
           "mpl_toolkits.mplot3d", "termcolor", "easydict", "isaacgym", "optuna", "optuna.samplers", "loguru", "rich",
           "rich.progress", "ipdb", "pynput"]


def import_reference():
    if not REF.exists():
        raise SystemExit("/root/reference is not present: golden vectors can only be regenerated in the build container")
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REF / "isaac_utils"))
    for name in STUBBED:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = mock.MagicMock(name=name)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import scripts.eval as ev
    import scripts.mass_landscape as ml
    import scripts.mass_opt as mo
    from spigym.envs.legged_base_task.legged_robot_base import LeggedRobotBase
    from spigym.envs.locomotion.go2_omni import go2_omni_interface
    from spigym.envs.sysid.active_sysid_openloop import ActiveSysId_OpenLoop
    import spigym.utils.torch_utils as tu
    return SimpleNamespace(ev=ev, ml=ml, mo=mo, LeggedRobotBase=LeggedRobotBase, go2_omni=go2_omni_interface,
                           ActiveSysId=ActiveSysId_OpenLoop, tu=tu)


def control_config(decimation=4):
    """The values of spigym/config/robot/go2/go2.yaml:107-109 and config/simulator/isaacgym.yaml:16 as the
    attribute tree the reference methods read."""
    return SimpleNamespace(
        robot=SimpleNamespace(control=SimpleNamespace(action_scale=0.25, action_clip_value=20.0, control_type="P",
                                                      clip_torques=True)),
        domain_rand=SimpleNamespace(randomize_ctrl_delay=False, randomize_torque_rfi=False),
        simulator=SimpleNamespace(config=SimpleNamespace(sim=SimpleNamespace(control_decimation=decimation))),
    )


# ------------------------------------------------------------------------------------------------
def gen_torques(R, rng):
    from spi_active_b200 import go2_model as gm
    m = gm.go2_nominal()
    N = 64
    actions = rng.uniform(-30, 30, (N, 12)).astype(np.float32)
    q = rng.uniform(-2.5, 2.5, (N, 12)).astype(np.float32)
    qd = rng.uniform(-25, 25, (N, 12)).astype(np.float32)
    kp = rng.uniform(15, 40, 12).astype(np.float32)
    kd = rng.uniform(0.3, 1.2, 12).astype(np.float32)
    self = SimpleNamespace(
        config=control_config(), device="cpu", num_envs=N, log_dict={},
        p_gains=torch.from_numpy(kp), d_gains=torch.from_numpy(kd),
        _kp_scale=torch.ones(N, 12), _kd_scale=torch.ones(N, 12),
        default_dof_pos=torch.tensor(m.q_default, dtype=torch.float32).unsqueeze(0),
        torque_limits=torch.tensor(m.torque_limit, dtype=torch.float32),
        simulator=SimpleNamespace(dof_pos=torch.from_numpy(q), dof_vel=torch.from_numpy(qd)),
    )
    R.LeggedRobotBase._pre_physics_step(self, torch.from_numpy(actions.copy()))
    clipped = self.actions_after_delay.clone()
    tau_base = R.LeggedRobotBase._compute_torques(self, clipped.clone())
    tau_omni = R.go2_omni._compute_torques(self, clipped.clone())
    out = dict(actions=actions, q=q, qd=qd, kp=kp, kd=kd, clipped_actions=clipped.numpy(),
               tau_base=tau_base.numpy(), tau_omni=tau_omni.numpy())
    # motor models on top of the clipped nominal torques (active_sysid_openloop.py:184-186)
    g_scalar = rng.uniform(0.6, 1.3, N).astype(np.float32)
    g3 = rng.uniform(0.6, 1.3, (N, 3)).astype(np.float32)
    a3 = rng.uniform(8.0, 40.0, (N, 3)).astype(np.float32)
    out["scalar_gain"] = g_scalar
    out["vec3_gain"] = g3
    out["tanh_a"] = a3
    for label, nominal in (("base", tau_base), ("omni", tau_omni)):
        out[f"tau_scalar_{label}"] = R.ActiveSysId.act2tau_scalar(self, nominal.clone(), scalar_gain=g_scalar.tolist()).numpy()
        out[f"tau_vec3_{label}"] = R.ActiveSysId.act2tau_vec3(self, nominal.clone(), hip_gain=g3[:, 0].tolist(),
                                                              thigh_gain=g3[:, 1].tolist(), calf_gain=g3[:, 2].tolist()).numpy()
        self.params_dict = {"motor_model_hip_a": {"value": a3[:, 0].tolist()},
                            "motor_model_thigh_a": {"value": a3[:, 1].tolist()},
                            "motor_model_calf_a": {"value": a3[:, 2].tolist()}}
        out[f"tau_tanh_{label}"] = R.ActiveSysId.act2tau_vec3_tanh(self, nominal.clone()).numpy()
    np.savez(HERE / "torques.npz", **out)
    print("torques.npz", {k: v.shape for k, v in out.items() if k.startswith("tau")})


def synthetic_recording(rng, T):
    quat = rng.standard_normal((T, 4)); quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    return dict(joint_positions=rng.standard_normal((T, 12)).astype(np.float32),
                joint_velocities=rng.standard_normal((T, 12)).astype(np.float32),
                joint_torques=rng.standard_normal((T, 12)).astype(np.float32),
                actions=rng.standard_normal((T, 12)),  # float64 on purpose: load_dataset casts (eval.py:121)
                base_positions=rng.standard_normal((T, 3)).astype(np.float32),
                base_orientations=quat.astype(np.float32),
                base_linear_velocities=rng.standard_normal((T, 3)).astype(np.float32),
                base_angular_velocities=rng.standard_normal((T, 3)).astype(np.float32),
                timestamps=np.arange(T) * 0.02, sim_duration=T * 0.02, data_frequency=50, robot_type="go2",
                pd_gain_kp=np.full(12, 20.0 + T, np.float32), pd_gain_kd=np.full(12, 0.5 + 0.01 * T, np.float32))


def gen_windowing(R, rng):
    recs = [synthetic_recording(rng, 12), synthetic_recording(rng, 9)]
    out = {}
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i, r in enumerate(recs):
            p = Path(d) / f"rec{i}.npz"
            np.savez(p, **r)
            paths.append(p)
            for k, v in r.items():
                out[f"in{i}_{k}"] = np.asarray(v)
        for H in (3, 5):
            total, ds = R.ev.load_dataset(paths, H)
            out[f"H{H}_total"] = np.int64(total)
            for k, v in ds.items():
                out[f"H{H}_{k}"] = np.asarray(v)
    np.savez(HERE / "windowing.npz", **out)
    print("windowing.npz", int(out["H3_total"]), int(out["H5_total"]))


def gen_fim(R, rng):
    M, P = 6, 4
    N = M * (P + 1)
    root = rng.standard_normal((N, 13)).astype(np.float32)
    dof = rng.standard_normal((N, 12)).astype(np.float32)
    origins = rng.standard_normal((N, 3)).astype(np.float32)
    idx = torch.arange(0, N).view(M, P + 1)
    self = SimpleNamespace(simulator=SimpleNamespace(robot_root_states=torch.from_numpy(root), dof_pos=torch.from_numpy(dof)),
                           env_origins=torch.from_numpy(origins), main_idx=idx[:, 0:1], aux_idx=idx[:, 1:],
                           delta_param=0.1, param_dim=P)
    rew = R.ActiveSysId._reward_fisher_information_matrix(self).numpy()
    np.savez(HERE / "fim.npz", root=root, dof=dof, origins=origins, M=M, P=P, delta=np.float32(0.1), reward=rew)
    print("fim.npz", rew.shape)


def gen_cost(R, rng):
    costs = rng.uniform(0, 0.1, (20, 3)).astype(np.float32)
    totals = np.array([R.mo.compute_cost(row.astype(np.float64)) for row in costs], dtype=np.float64)
    w = R.ml.COST_COEFF
    np.savez(HERE / "cost.npz", costs=costs, totals=totals,
             weights=np.array([w["base_pos"], w["base_quat"], w["joint_pos"]]),
             mass_scale_range=np.array([R.ml.MASS_SCALE_MIN, R.ml.MASS_SCALE_MAX]), mass_samples=R.ml.MASS_SAMPLES,
             max_safe_env_batch=R.ml.MAX_SAFE_ENV_BATCH)
    print("cost.npz", totals[:3])


def gen_quat(R, rng):
    N = 32
    q = rng.standard_normal((N, 4)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
    v = rng.standard_normal((N, 3)).astype(np.float32)
    out = dict(q=q, v=v, rotate_inverse=R.tu.quat_rotate_inverse(torch.from_numpy(q), torch.from_numpy(v)).numpy())
    if hasattr(R.tu, "quat_rotate"):
        out["rotate"] = R.tu.quat_rotate(torch.from_numpy(q), torch.from_numpy(v)).numpy()
    np.savez(HERE / "quat.npz", **out)
    print("quat.npz")


# ------------------------------------------------------------------------------------------------
def make_harness_env(R, num_envs):
    """An env object whose control path is the reference's real code and whose simulator is OracleSim."""
    from oracle_sim import OracleSim
    from spi_active_b200 import go2_model as gm
    m = gm.go2_nominal()

    class HarnessEnv:
        _pre_physics_step = R.LeggedRobotBase._pre_physics_step
        _physics_step = R.LeggedRobotBase._physics_step
        _apply_force_in_physics_step = R.LeggedRobotBase._apply_force_in_physics_step
        _compute_torques = R.LeggedRobotBase._compute_torques

        def render(self):
            pass

        def step(self, actor_state):
            # LeggedRobotBase.step (:169-183) without _post_physics_step: that only refreshes the tensors
            # and computes rewards / observations / resets that evaluate_batch discards (terminations are
            # off in the replay config: SURVEY.md §8a row 7)
            self._pre_physics_step(actor_state["actions"])
            self._physics_step()
            self.simulator.refresh_sim_tensors()

        def reset_all(self):
            # base_task.py:90-100 randomises the state and burns one env step; evaluate_batch overwrites
            # the state right after, so it has no observable effect (quirk D5)
            pass

    env = HarnessEnv()
    N = num_envs
    env.config = control_config()
    env.device, env.num_envs, env.dim_actions, env.log_dict = "cpu", N, 12, {}
    env.simulator = OracleSim(N, m)
    env.p_gains = torch.tensor(m.kp, dtype=torch.float32)
    env.d_gains = torch.tensor(m.kd, dtype=torch.float32)
    env._kp_scale, env._kd_scale = torch.ones(N, 12), torch.ones(N, 12)
    env.default_dof_pos = torch.tensor(m.q_default, dtype=torch.float32).unsqueeze(0)
    env.torque_limits = torch.tensor(m.torque_limit, dtype=torch.float32)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt)
    env.torques, env.actions, env.last_actions, env.actions_after_delay = z(N, 12), z(N, 12), z(N, 12), z(N, 12)
    env.last_dof_vel, env.feet_air_time = z(N, 12), z(N, 4)
    env.episode_length_buf, env.reset_buf = z(N, dt=torch.long), z(N, dt=torch.long)
    return env


def gen_mass_sweep(R, rng):
    import synth
    from spi_active_b200 import recorders
    out = {}
    # three short recordings with different PD gains per file (exercises D3), recorded with the oracle
    recs = []
    for name, steps, kp, kd in (("jump", 40, 25.0, 0.6), ("sine", 33, 22.0, 0.5), ("stand", 21, 28.0, 0.7)):
        r = dict(synth.recording(name, steps))
        r["pd_gain_kp"] = np.full(12, kp, np.float32)
        r["pd_gain_kd"] = np.full(12, kd, np.float32)
        r.update(timestamps=np.arange(steps) * 0.02, sim_duration=steps * 0.02, data_frequency=50, robot_type="go2")
        recs.append(r)
    scales = np.array([0.6, 1.0, 1.7])
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i, r in enumerate(recs):
            p = Path(d) / f"rec{i}.npz"
            np.savez(p, **r)
            paths.append(p)
            for k in ("joint_positions", "joint_velocities", "actions", "base_positions", "base_orientations",
                      "base_linear_velocities", "base_angular_velocities", "pd_gain_kp", "pd_gain_kd"):
                out[f"rec{i}_{k}"] = np.asarray(r[k])
        for H in (5, 3):
            total, ds_np = R.ev.load_dataset(paths, H)
            for B in (16, 32, 4096):
                batch = min(B, total, R.ml.MAX_SAFE_ENV_BATCH)   # mass_landscape.py:146
                env = make_harness_env(R, batch)
                dataset = R.ev.to_device(ds_np, env.device)
                ref_masses = R.ev.capture_reference_masses(env.simulator)
                res = R.ml.mass_sweep(env, dataset, H, ref_masses, scales)
                out[f"H{H}_B{B}_costs"] = res
                out[f"H{H}_B{B}_batch"] = np.int64(batch)
                print(f"mass_sweep H={H} B={B} (batch {batch}) S={total}:", res.tolist())
        out["ref_masses"] = ref_masses
    out["scales"] = scales
    np.savez(HERE / "mass_sweep.npz", **out)


def gen_active_obs(R, rng):
    """The 900-dim actor observation of the active-exploration env over a few steps, produced by the reference's real
    go2_omni._step_contact_targets, LeggedRobotBase._pre_compute_observations_callback / _compute_observations /
    _post_config_observation_callback, every _get_obs_* getter, helpers.parse_observation and HistoryHandler."""
    from spigym.envs.env_utils.history_handler import HistoryHandler
    from spi_active_b200 import active as act, go2_model as gm
    N, T = 7, 18
    keys = list(act.OBS_DIMS)
    obs_cfg = SimpleNamespace(
        obs_dict={"actor_obs": ["base_ang_vel", "projected_gravity", "command_lin_vel", "command_ang_vel",
                                "command_body_height", "command_gait_freq", "command_gait_phase",
                                "command_footswing_height", "command_body_attitude", "command_stance", "clock_inputs",
                                "dof_pos", "dof_vel", "actions", "short_history"]},
        obs_auxiliary={"short_history": {k: 14 for k in ["base_ang_vel", "projected_gravity", "dof_pos", "dof_vel", "actions",
                                                         "command_lin_vel", "command_ang_vel", "command_body_height",
                                                         "command_gait_freq", "command_gait_phase", "command_footswing_height",
                                                         "command_body_attitude", "command_stance", "clock_inputs"]}},
        obs_scales=dict(act.OBS_SCALES, short_history=1.0), noise_scales={k: 0.0 for k in keys + ["short_history"]},
        commands=SimpleNamespace(limit_body_height=[-0.25, 0.15], num_commands=14))

    class Stub:
        pass
    for name in dir(R.go2_omni):
        if name.startswith("_get_obs_"):
            setattr(Stub, name, getattr(R.go2_omni, name))
    Stub._step_contact_targets = R.go2_omni._step_contact_targets
    Stub._pre_compute_observations_callback = R.LeggedRobotBase._pre_compute_observations_callback
    Stub._compute_observations = R.LeggedRobotBase._compute_observations
    Stub._post_config_observation_callback = R.LeggedRobotBase._post_config_observation_callback
    self = Stub()
    self.config = SimpleNamespace(obs=obs_cfg)
    self.is_evaluating, self.dt, self.device, self.num_envs = True, 0.02, "cpu", N
    z = lambda *s: torch.zeros(*s, dtype=torch.float32)
    self.gait_indices, self.clock_inputs = z(N), z(N, 4)
    self.doubletime_clock_inputs, self.halftime_clock_inputs, self.desired_contact_states = z(N, 4), z(N, 4), z(N, 4)
    self.simulator = SimpleNamespace(robot_root_states=z(N, 13), dof_pos=z(N, 12), dof_vel=z(N, 12))
    self.simulator.base_quat = self.simulator.robot_root_states[:, 3:7]
    self.base_quat, self.rpy = z(N, 4), z(N, 3)
    self.base_lin_vel, self.base_ang_vel, self.projected_gravity = z(N, 3), z(N, 3), z(N, 3)
    self.gravity_vec = torch.tensor([0.0, 0.0, -1.0]).repeat(N, 1)
    self.default_dof_pos = torch.tensor(gm.go2_nominal().q_default, dtype=torch.float32).unsqueeze(0)
    self.history_handler = HistoryHandler(N, obs_cfg.obs_auxiliary, dict(act.OBS_DIMS), "cpu")
    ranges = np.asarray(act.COMMAND_RANGES)
    out = dict(states=[], actions=[], commands_clock=[], commands_obs=[], actor_obs=[], clock=[], gait=[])
    for t in range(T):
        cmd_clock = rng.uniform(ranges[:, 0], ranges[:, 1], (N, 14)).astype(np.float32)
        cmd_clock[:, 8] = rng.uniform(0.3, 0.7, N)            # exercise the stance / swing warp with durations != 0.5
        cmd_obs = rng.uniform(ranges[:, 0], ranges[:, 1], (N, 14)).astype(np.float32)
        state = rng.standard_normal((N, 37)).astype(np.float32)
        state[:, 3:7] /= np.linalg.norm(state[:, 3:7], axis=1, keepdims=True)
        if t == 5:
            state[:, 25:37] *= 3000.0                          # exercise the +-100 observation clip
        actions = rng.uniform(-3, 3, (N, 12)).astype(np.float32)
        self.commands = torch.from_numpy(cmd_clock.copy())
        self._step_contact_targets()
        self.simulator.robot_root_states[:] = torch.from_numpy(state[:, :13])
        self.simulator.dof_pos[:] = torch.from_numpy(state[:, 13:25])
        self.simulator.dof_vel[:] = torch.from_numpy(state[:, 25:37])
        self.actions = torch.from_numpy(actions.copy())
        self.commands = torch.from_numpy(cmd_obs.copy())       # _update_tasks_callback runs before the observations
        self._pre_compute_observations_callback()
        self._compute_observations()
        obs = torch.clip(self.obs_buf_dict["actor_obs"], -100.0, 100.0)
        for key in self.history_handler.history.keys():       # legged_robot_base.py:249-250
            self.history_handler.add(key, self.hist_obs_dict[key])
        for k, v in (("states", state), ("actions", actions), ("commands_clock", cmd_clock), ("commands_obs", cmd_obs),
                     ("actor_obs", obs.numpy().copy()), ("clock", self.clock_inputs.numpy().copy()),
                     ("gait", self.gait_indices.numpy().copy())):
            out[k].append(v)
    np.savez(HERE / "active_obs.npz", **{k: np.stack(v) for k, v in out.items()})
    print("active_obs.npz", np.stack(out["actor_obs"]).shape)


def gen_urdf_blob():
    from spi_active_b200 import go2_model as gm
    urdf = REF / "spigym" / "data" / "robots" / "go2" / "urdf" / "go2.urdf"
    m = gm.model_from_urdf(urdf)
    blob = gm.build_model_blob(m)
    np.savez(HERE / "urdf_blob.npz", blob=blob, body_masses=m.body_masses_isaac_order(), total_mass=np.float64(m.total_mass()))
    print("urdf_blob.npz", blob.shape, "== go2_nominal():", bool(np.array_equal(blob, gm.build_model_blob(gm.go2_nominal()))))


class AttrDict(dict):
    """omegaconf / easydict stand-in: the reference indexes its module config both ways (modules.py:13, 19)."""
    __getattr__ = dict.__getitem__


def gen_ppo_actor():
    from spigym.agents.modules.ppo_modules import PPOActor, PPOCritic
    torch.manual_seed(5)
    dims = dict(actor_obs=60, critic_obs=70)
    layer = lambda: AttrDict(type="MLP", hidden_dims=[32, 16, 8], activation="ELU")      # ppo.yaml:31-39, scaled down
    actor = PPOActor(obs_dim_dict=dims, module_config_dict=AttrDict(input_dim=["actor_obs"], output_dim=["robot_action_dim"],
                                                                   layer_config=layer()), num_actions=12, init_noise_std=0.8)
    critic = PPOCritic(dims, AttrDict(type="MLP", input_dim=["critic_obs"], output_dim=[1], layer_config=layer()))
    with torch.no_grad():
        for prm in actor.parameters():
            prm.add_(0.3 * torch.randn_like(prm))          # non-trivial biases too
    torch.save({"actor_model_state_dict": actor.state_dict(), "critic_model_state_dict": critic.state_dict(),
                "actor_optimizer_state_dict": {}, "critic_optimizer_state_dict": {}, "iter": 123, "infos": None},
               HERE / "ppo_actor.pt")
    obs = torch.randn(9, 60) * 2.0
    with torch.no_grad():
        out = actor.act_inference(obs)
    np.savez(HERE / "ppo_actor.npz", obs=obs.numpy(), actions=out.numpy(), keys=np.array(list(actor.state_dict().keys())))
    print("ppo_actor.pt", list(actor.state_dict().keys()))


def gen_b1_env():
    import test_b1_reference_env as tb
    out = {}
    with tempfile.TemporaryDirectory() as d:
        total, ds_np, ref_masses, res = tb.run_reference_sweep(d)
    for i, r in enumerate(tb._recordings()):
        for k in ("joint_positions", "joint_velocities", "actions", "base_positions", "base_orientations",
                  "base_linear_velocities", "base_angular_velocities", "pd_gain_kp", "pd_gain_kd"):
            out[f"rec{i}_{k}"] = np.asarray(r[k])
    for B, (batch, costs) in res.items():
        out[f"B{B}_costs"], out[f"B{B}_batch"] = costs, np.int64(batch)
    out["scales"], out["ref_masses"], out["H"] = np.array([0.6, 1.0, 1.7]), ref_masses, np.int64(5)
    np.savez(HERE / "b1_env.npz", **out)
    print("b1_env.npz", {k: v.tolist() for k, v in out.items() if k.endswith("_costs")})


def gen_commands(rng):
    """The reference's real command samplers; `self` carries what _init_config (active_sysid.py:101-121) would load from
    config/algo/active_sysid.yaml:23-50."""
    import yaml
    from spigym.agents.sysid.active_sysid import ActiveSysId
    cc = yaml.safe_load((REF / "spigym" / "config" / "algo" / "active_sysid.yaml").read_text())["algo"]["config"]["command"]
    dt = 0.02
    out = dict(num_updates=np.int64(cc["rollout_length"] // cc["horizon_length"]),
               steps_per_update=np.int64(int(cc["horizon_length"] / dt)))
    for mode in ("constant", "polynomial", "bezier"):
        self = SimpleNamespace(default_command=np.array(cc["default_command"], dtype=np.float32),
                               command_sampling_idxs=cc["command_sampling_idxs"], command_sampling_mode=mode,
                               num_steps_per_update=int(out["steps_per_update"]), poly_degree=cc["poly_degree"],
                               num_bezier_points=cc["num_bezier_points"])
        for name in ("_sample_constant", "_sample_polynomial", "_sample_bezier", "_binomial_coefficient"):
            setattr(self, name, getattr(ActiveSysId, name).__get__(self))
        names, values, lows, highs = [], [], [], []

        class Trial:
            def suggest_float(_, name, low, high):
                v = float(rng.uniform(low, high)) if mode != "bezier" else float(rng.uniform(low - 0.3 * (high - low), high + 0.3 * (high - low)))
                names.append(name); values.append(v); lows.append(low); highs.append(high)
                return v
        ranges = np.array(cc["command_ranges"], dtype=np.float32)
        cmds = ActiveSysId.sample_commands(self, Trial(), ranges, int(out["num_updates"]))
        out[f"{mode}_names"], out[f"{mode}_values"] = np.array(names), np.array(values)
        out[f"{mode}_low"], out[f"{mode}_high"] = np.array(lows), np.array(highs)
        out[f"{mode}_commands"] = cmds
        print("commands.npz", mode, cmds.shape, len(names))
    np.savez(HERE / "commands.npz", **out)


def main():
    R = import_reference()
    gen_commands(np.random.default_rng(11))
    gen_urdf_blob()
    gen_b1_env()
    gen_ppo_actor()
    rng = np.random.default_rng(20251017)
    gen_torques(R, rng)
    gen_windowing(R, rng)
    gen_fim(R, rng)
    gen_cost(R, rng)
    gen_quat(R, rng)
    gen_mass_sweep(R, rng)
    gen_active_obs(R, np.random.default_rng(7))


if __name__ == "__main__":
    main()
