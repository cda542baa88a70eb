"""`simulator=b200` — the stepwise simulator plugin (boundary B1 of SURVEY.md §8b).

Mirrors the reference's plugin interface method for method:
  * spigym/simulator/base_simulator/base_simulator.py:6-171   BaseSimulator (17 methods)
  * spigym/simulator/isaacgym/isaacgym.py:31-62      setup            (sim_dt, device)
  * :170-180, 219-224                                load_assets      (num_dof / num_bodies / names, asserted vs config)
  * :226-272                                         create_envs
  * :314-349, 519-531                                get_dof_limits_properties (hard / soft / termination limits)
  * :533-534                                         find_rigid_body_indice
  * :536-577                                         prepare_sim      (the tensors the env layer reads directly)
  * :579-589                                         refresh_sim_tensors
  * :598-599, 601-620, 622-626                       apply_torques_at_dof / set_*_state_tensor / simulate_at_each_physics_step
  * spigym/simulator/isaacgym/isaacgym_active_sysid.py:34-94   per-env `params_dict` overrides (mass, com*, inertia*)
and the raw Isaac Gym handle surface scripts/eval.py:185-214 reaches through (`gym`, `sim`, `envs`, `robot_handles`).

The reference instantiates `get_class(config.simulator._target_)(config=<env config>, device=<str>)`
(spigym/envs/base_task/base_task.py:27-28); `config/simulator/b200.yaml` in this repo points `_target_` here.

Physics = `spi_b200_sim_step` (CUDA).  There is no CPU fallback: the default backend is the CUDA RolloutEngine and
raises without a GPU.  (`backend=` exists so that tests can exercise the host logic of this class on a CPU box with an
object of the same two-method shape; the product never passes it.)
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional, Sequence

import numpy as np
import torch

from . import go2_model as gm

try:  # subclass the reference's base class when the reference is importable, else an identical local mirror
    from spigym.simulator.base_simulator.base_simulator import BaseSimulator  # type: ignore
except Exception:  # pragma: no cover - the reference is not a dependency
    class BaseSimulator:  # noqa: D401 - same surface as base_simulator.py:6-171
        def __init__(self, config, device):
            self.config = config
            self.sim_device = device
            self.headless = False

        def set_headless(self, headless):
            self.headless = headless


def _get(cfg, path: str, default=None):
    """cfg.a.b.c for OmegaConf / SimpleNamespace / dict trees; `default` when any level is missing."""
    cur = cfg
    for key in path.split("."):
        if cur is None:
            return default
        if isinstance(cur, dict):
            cur = cur.get(key, None)
        else:
            cur = getattr(cur, key, None)
    return default if cur is None else cur


# isaacgym_active_sysid.py:61-94 setter names -> engine parameter names (the `inertiaiy` typo, quirk D8, is
# accepted as an alias; whether `inertiay` takes effect is the engine flag SPI_FLAG_STRICT_INERTIAY)
PARAMS_DICT_KEYS = {"mass": "mass", "comx": "comx", "comy": "comy", "comz": "comz", "inertiax": "inertiax",
                    "inertiay": "inertiay", "inertiaiy": "inertiay", "inertiaz": "inertiaz"}


# the 19 Isaac Gym bodies (gm.BODY_NAMES order) -> the 13 moving bodies of the engine (base, then hip / thigh / calf of FL,
# FR, RL, RR): feet are rigidly attached to their calf, Head_upper / Head_lower to the base
MOVING_BODY_OF_ISAAC_BODY = [0, 1, 2, 3, 3, 4, 5, 6, 6, 0, 0, 7, 8, 9, 9, 10, 11, 12, 12]
ISAAC_BODY_OF_MOVING_BODY = [0, 1, 2, 3, 5, 6, 7, 11, 12, 13, 15, 16, 17]


def go2_body_spheres():
    """Contact spheres for the links that are NOT feet: (Isaac body index, centre in that link's frame, radius), read off the
    collision geometry of spigym/data/robots/go2/urdf/go2.urdf — the 8 corners of the trunk box (0.3762 x 0.0935 x 0.114), the
    two head shapes, the hip cylinders (r = 0.046 at y = +-0.08), the two ends of the thigh box (0.11 x 0.0245 x 0.034 along z at
    z = -0.1065) and of the calf cylinder (r = 0.012, 0.12 long, tilted -0.21 rad about y at (0.008, 0, -0.06)).  Isaac Gym
    collides those shapes with the plane; here every sphere uses the engine's compliant foot-contact law."""
    out = []
    for sx in (-1.0, 1.0):
        for sy in (-1.0, 1.0):
            for sz in (-1.0, 1.0):
                out.append((0, (sx * 0.1881, sy * 0.04675, sz * 0.057), 0.0))
    out.append((gm.BODY_NAMES.index("Head_upper"), (0.0, 0.0, 0.0), 0.05))
    out.append((gm.BODY_NAMES.index("Head_lower"), (0.0, 0.0, 0.0), 0.047))
    ax = (float(np.sin(-0.21)), 0.0, float(np.cos(-0.21)))
    for leg in gm.LEGS:
        side = 1.0 if leg[1] == "L" else -1.0
        out.append((gm.BODY_NAMES.index(f"{leg}_hip"), (0.0, side * 0.08, 0.0), 0.046))
        for z in (-0.0515, -0.1615):
            out.append((gm.BODY_NAMES.index(f"{leg}_thigh"), (0.0, 0.0, z), 0.017))
        for e in (-0.06, 0.06):
            out.append((gm.BODY_NAMES.index(f"{leg}_calf"), (0.008 + e * ax[0], 0.0, -0.06 + e * ax[2]), 0.012))
    return out


class _RigidBodyProps:
    """The fields of gymapi.RigidBodyProperties the reference touches (mass, com, inertia)."""

    def __init__(self, mass, com, inertia):
        self.mass = float(mass)
        self.com = SimpleNamespace(x=float(com[0]), y=float(com[1]), z=float(com[2]))
        xx, yy, zz, xy, xz, yz = [float(v) for v in inertia]
        self.inertia = SimpleNamespace(x=SimpleNamespace(x=xx, y=xy, z=xz), y=SimpleNamespace(x=xy, y=yy, z=yz),
                                       z=SimpleNamespace(x=xz, y=yz, z=zz))


class _GymShim:
    """`simulator.gym` as used by scripts/eval.py:185-214 and scripts/data/common.py:105-108."""

    def __init__(self, sim: "B200Sim"):
        self._s = sim

    def get_actor_rigid_body_properties(self, env_ptr, actor):
        s, m = self._s, self._s.model
        e = int(env_ptr)
        out = []
        for name in gm.BODY_NAMES:
            if name == "base":
                p = s._params_host[e]
                out.append(_RigidBodyProps(p[0], p[1:4], p[4:10]))
            elif name.startswith("Head"):
                i, pos = m.base_lumps[0 if name == "Head_upper" else 1]
                out.append(_RigidBodyProps(i.mass, i.com, i.inertia))
            else:
                leg, part = name.split("_")
                li = gm.LEGS.index(leg)
                i = m.feet[li] if part == "foot" else m.leg_bodies[3 * li + ("hip", "thigh", "calf").index(part)]
                out.append(_RigidBodyProps(i.mass, i.com, i.inertia))
        return out

    def set_actor_rigid_body_properties(self, env_ptr, actor, props, recomputeInertia=True):
        s = self._s
        ref = s.model.body_masses_isaac_order()
        for i in range(1, len(ref)):
            if abs(props[i].mass - float(ref[i])) > 1e-6:
                raise NotImplementedError("the b200 backend parametrises the base link only (SURVEY.md §8a row 9)")
        e = int(env_ptr)
        p = props[0]
        new_mass = float(p.mass)
        row = s._params_host[e]
        if recomputeInertia and not s.inertia_keep and row[0] > 0:
            row[4:10] *= new_mass / row[0]      # quirk D15 default: the inertia tensor follows the mass
        row[0] = new_mass
        row[1:4] = [p.com.x, p.com.y, p.com.z]
        if not recomputeInertia:
            row[4:10] = [p.inertia.x.x, p.inertia.y.y, p.inertia.z.z, p.inertia.x.y, p.inertia.x.z, p.inertia.y.z]
        s._params_dirty = True
        return True

    def refresh_mass_matrix_tensors(self, sim):
        self._s._upload_params()
        return True

    def find_actor_rigid_body_handle(self, env_ptr, actor, name):
        return gm.BODY_NAMES.index(name)

    def destroy_sim(self, sim):
        self._s.close()


class B200Sim(BaseSimulator):
    PARAM_NAMES = ["mass", "comx", "comy", "comz", "inertiax", "inertiay", "inertiaz", "inertiaxy", "inertiaxz",
                   "inertiayz"]

    def __init__(self, config=None, device="cuda:0", backend=None, model: Optional[gm.Go2Model] = None):
        super().__init__(config, device)
        self.simulator_config = _get(config, "simulator.config")
        self.robot_config = _get(config, "robot")
        self.model = model or gm.go2_nominal()
        # config/simulator/b200.yaml `b200.contact`: the engine's compliant foot-contact constants; they are part of the model
        # blob the engine is created with (setup()), so they must be applied before that
        contact_cfg = _get(config, "simulator.config.b200.contact")
        if contact_cfg is not None:
            cp = self.model.contact
            for key, cast in (("kn", float), ("cn", float), ("mu", float), ("dt", float), ("veps", float), ("nsub", int)):
                v = _get(contact_cfg, key)
                if v is not None:
                    setattr(cp, key, cast(v))
        self._backend = backend
        # b200.contact.body_spheres: collide the non-foot links with the plane as well (go2_body_spheres).  Off by default:
        # the fused operator (boundary B2) models the feet only, and the two boundaries must agree on recorded data
        self.body_spheres = go2_body_spheres() if bool(_get(contact_cfg, "body_spheres", False)) else None
        self.inertia_keep = bool(_get(config, "simulator.config.b200.inertia_keep", False))
        self.strict_inertiay = bool(_get(config, "simulator.config.b200.strict_inertiay", False))
        self.params_dict = _get(config, "params_dict")   # isaacgym_active_sysid.py:34
        self.viewer = None

    # ----- configuration ---------------------------------------------------------------------------------------
    def setup(self):
        fps = _get(self.config, "simulator.config.sim.fps", None)
        self.sim_dt = 1.0 / float(fps) if fps else float(self.model.dt)
        if abs(self.sim_dt - self.model.dt) > 1e-12:
            self.model.dt = self.sim_dt
        self.device = self.sim_device
        if self._backend is None:
            from .engine import RolloutEngine   # raises SpiB200Error without libspi_b200.so or without a GPU
            self._backend = RolloutEngine(self.model, torch.device(self.device))
        self.sim = self            # opaque handle, as `simulator.sim` is for callers
        self.gym = _GymShim(self)

    def setup_terrain(self, mesh_type):
        if mesh_type not in ("plane", None):
            raise NotImplementedError(f"the b200 backend models the infinite plane only, got terrain {mesh_type!r}")
        mu = _get(self.config, "simulator.config.terrain.static_friction", None)
        if mu is not None and abs(float(mu) - self.model.contact.mu) > 1e-9:
            raise NotImplementedError("plane friction is part of the model blob; rebuild the model with contact.mu")

    # ----- assets / envs ------------------------------------------------------------------------------------------
    def load_assets(self):
        self.num_dof, self.num_bodies = 12, 19
        self.dof_names, self.body_names = list(gm.DOF_NAMES), list(gm.BODY_NAMES)
        cfg_dofs, cfg_bodies = _get(self.robot_config, "dof_names"), _get(self.robot_config, "body_names")
        if cfg_dofs is not None:   # isaacgym.py:177-180
            assert self.num_dof == len(cfg_dofs), "Number of DOFs must be equal to number of actions"
            assert self.dof_names == list(cfg_dofs), "DOF names must match the config"
        if cfg_bodies is not None:
            assert self.num_bodies == len(cfg_bodies), "Number of bodies must be equal to number of body names"
            assert self.body_names == list(cfg_bodies), "Body names must match the config"

    def create_envs(self, num_envs, env_origins, base_init_state, env_config=None):
        self.num_envs = int(num_envs)
        self.env_config = self.config
        self.env_origins = env_origins
        self.base_init_state = base_init_state
        self.envs = list(range(self.num_envs))
        self.robot_handles = [0] * self.num_envs
        nominal = gm.default_param_vector(self.model)[:10].astype(np.float64)
        self._params_host = np.tile(nominal[None], (self.num_envs, 1))
        if self.params_dict is not None:   # per-env overrides applied at creation (isaacgym_active_sysid.py:39-59)
            items = self.params_dict.items() if hasattr(self.params_dict, "items") else vars(self.params_dict).items()
            for key, spec in items:
                if "motor_model" in key:
                    continue           # motor-model parameters belong to the env, not the simulator
                name = PARAMS_DICT_KEYS.get(key)
                if name is None:
                    continue           # the reference logs unknown keys at debug level and moves on (:52-56)
                if name == "inertiay" and self.strict_inertiay:
                    continue
                values = np.asarray(_get(spec, "value"), dtype=np.float64)
                col = self.PARAM_NAMES.index(name)
                if name == "mass" and not self.inertia_keep:
                    self._params_host[:, 4:10] *= (values / self._params_host[:, 0])[:, None]
                self._params_host[:, col] = values
        self._params_dirty = True
        self._limits()

    def _limits(self):
        m, dev = self.model, self.device
        f = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device=dev)
        lo, hi = f(m.q_lower), f(m.q_upper)
        self.hard_dof_pos_limits = torch.stack([lo, hi], dim=1)
        mid, rng = (lo + hi) / 2, hi - lo
        soft = float(_get(self.config, "rewards.reward_limit.soft_dof_pos_limit", 1.0))
        term = float(_get(self.config, "termination_scales.termination_close_to_dof_pos_limit", 1.0))
        self.dof_pos_limits = torch.stack([mid - 0.5 * rng * soft, mid + 0.5 * rng * soft], dim=1)
        self.dof_pos_limits_termination = torch.stack([mid - 0.5 * rng * term, mid + 0.5 * rng * term], dim=1)
        self.dof_vel_limits = f(m.qd_limit)
        self.torque_limits = f(m.torque_limit)

    def get_dof_limits_properties(self):
        cfg = self.robot_config   # isaacgym.py:519-531: the URDF limits must agree with the YAML to 1e-5
        for key, ten in (("dof_vel_limit_list", self.dof_vel_limits), ("dof_effort_limit_list", self.torque_limits)):
            ref = _get(cfg, key)
            if ref is not None:
                assert torch.allclose(ten.cpu(), torch.tensor(list(ref), dtype=torch.float32), atol=1e-5), key
        return self.dof_pos_limits, self.dof_vel_limits, self.torque_limits

    def find_rigid_body_indice(self, body_name):
        return self.body_names.index(body_name)

    # ----- tensors --------------------------------------------------------------------------------------------------
    def prepare_sim(self):
        N, dev = self.num_envs, self.device
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self._state = z(N, gm.STATE_DIM)            # the engine-side state (what PhysX holds in the reference)
        init = self.base_init_state
        if init is not None:
            self._state[:, 0:13] = torch.as_tensor(init, dtype=torch.float32, device=dev).reshape(1, 13)
        else:
            self._state[:, 2] = 0.34
            self._state[:, 6] = 1.0
        self._state[:, 13:25] = torch.tensor(self.model.q_default, dtype=torch.float32, device=dev)
        self._torques = z(N, 12)
        self._ext_wrench = None
        self._foot_force = z(N, 4, 3)
        self._body_contact_force = z(N, 19, 3)      # non-foot contacts (b200.contact.body_spheres), world frame
        if self.body_spheres is not None:
            self._sphere_body = torch.tensor([b for b, _, _ in self.body_spheres], device=dev)
            self._sphere_offset = torch.tensor([o for _, o, _ in self.body_spheres], dtype=torch.float32, device=dev)
            self._sphere_radius = torch.tensor([r for _, _, r in self.body_spheres], dtype=torch.float32, device=dev)
        self.all_root_states = z(N, 13)
        self.robot_root_states = self.all_root_states            # one actor per env (isaacgym.py:567-571)
        self.base_quat = self.robot_root_states[..., 3:7]
        self.dof_state = z(N * 12, 2)
        self.dof_pos = self.dof_state.view(N, 12, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, 12, 2)[..., 1]
        self.contact_forces = z(N, 19, 3)
        self._rigid_body_state = z(N, 19, 13)
        self._rigid_body_pos = self._rigid_body_state[..., 0:3]
        self._rigid_body_rot = self._rigid_body_state[..., 3:7]
        self._rigid_body_vel = self._rigid_body_state[..., 7:10]
        self._rigid_body_ang_vel = self._rigid_body_state[..., 10:13]
        self._feet_idx = [self.body_names.index(f"{leg}_foot") for leg in gm.LEGS]
        self._upload_params()
        self.refresh_sim_tensors()

    def _upload_params(self):
        self._params = torch.tensor(self._params_host.astype(np.float32), device=self.device)
        self._params_dirty = False

    def refresh_sim_tensors(self):
        self.all_root_states.copy_(self._state[:, 0:13])
        self._refresh_dof()
        self.contact_forces.copy_(self._body_contact_force)
        self.contact_forces[:, self._feet_idx, :] = self._foot_force
        if hasattr(self._backend, "body_states"):                  # all 19 links by forward kinematics (CUDA)
            self._backend.body_states(self._state, out=self._rigid_body_state)
        else:
            self._rigid_body_state[:, 0, :] = self._state[:, 0:13]

    def _refresh_dof(self):
        self.dof_pos.copy_(self._state[:, 13:25])
        self.dof_vel.copy_(self._state[:, 25:37])

    # ----- control / stepping ---------------------------------------------------------------------------------------
    def apply_torques_at_dof(self, torques):
        self._torques.copy_(torques.reshape(self.num_envs, 12))

    def set_actor_root_state_tensor(self, set_env_ids, root_states):
        ids = set_env_ids.to(torch.long)
        rows = root_states.reshape(-1, 13)[ids].to(torch.float32)
        self._state[ids, 0:13] = rows
        self.all_root_states[ids] = rows

    def set_dof_state_tensor(self, set_env_ids, dof_states):
        ids = set_env_ids.to(torch.long)
        ds = dof_states.reshape(self.num_envs, 12, 2)
        self._state[ids, 13:25] = ds[ids, :, 0].to(torch.float32)
        self._state[ids, 25:37] = ds[ids, :, 1].to(torch.float32)
        if ds.data_ptr() != self.dof_state.data_ptr():
            self.dof_state.view(self.num_envs, 12, 2)[ids] = ds[ids].to(torch.float32)

    @staticmethod
    def _rotate(quat, v, inverse=False):
        """R v (or R^T v) for unit quaternions xyzw, body -> world."""
        u, w = quat[..., :3], quat[..., 3:4]
        t = 2.0 * torch.cross(u, v, dim=-1)
        return v + (-w if inverse else w) * t + torch.cross(u, t, dim=-1)

    def _wrench_from_point_forces(self, body, f, p_world):
        """World-frame forces f [N,K,3] acting at world positions p_world [N,K,3] on the Isaac bodies body [K] -> one
        [torque; force] wrench per MOVING body in its own link frame [N,13,6] (feet act on their calf, the head links on the
        base), the input of spi_b200_sim_step_ext.  self._rigid_body_state must hold the poses of the current state."""
        N, dev = self.num_envs, self._state.device
        mb = torch.tensor(MOVING_BODY_OF_ISAAC_BODY, device=dev)[body]                # [K] moving body of every point
        own = torch.tensor(ISAAC_BODY_OF_MOVING_BODY, device=dev)
        origin = self._rigid_body_state[:, own, 0:3]                                  # [N,13,3]
        quat = self._rigid_body_state[:, own, 3:7]                                    # [N,13,4]
        torque_w = torch.cross(p_world - origin[:, mb], f, dim=-1)                    # about the moving body's link origin
        fw = torch.zeros(N, 13, 3, device=dev).index_add_(1, mb, f)
        tw = torch.zeros(N, 13, 3, device=dev).index_add_(1, mb, torque_w)
        return torch.cat([self._rotate(quat, tw, inverse=True), self._rotate(quat, fw, inverse=True)], dim=-1).contiguous()

    def apply_rigid_body_force_at_pos_tensor(self, force_tensor, pos_tensor):
        """isaacgym.py:609-613: forces [N,19,3] on the 19 Isaac Gym bodies at positions [N,19,3] in ENV_SPACE (world axes,
        relative to the env origin), consumed by the NEXT physics step.  Converted here to one [torque; force] wrench per
        moving body in its own link frame (feet act on their calf, the head links on the base) for spi_b200_sim_step_ext."""
        N, dev = self.num_envs, self._state.device
        f = torch.as_tensor(force_tensor, dtype=torch.float32, device=dev).reshape(N, 19, 3)
        p = torch.as_tensor(pos_tensor, dtype=torch.float32, device=dev).reshape(N, 19, 3)
        if self.env_origins is not None:
            p = p + torch.as_tensor(self.env_origins, dtype=torch.float32, device=dev).reshape(N, 1, 3)
        if hasattr(self._backend, "body_states"):
            self._backend.body_states(self._state, out=self._rigid_body_state)       # poses of the CURRENT state
        else:
            raise NotImplementedError("external forces need a backend with body_states")
        self._ext_wrench = self._wrench_from_point_forces(torch.arange(19, device=dev), f, p)

    def _body_sphere_wrench(self):
        """Non-foot contacts (b200.contact.body_spheres): every sphere of go2_body_spheres() against the plane z = 0 with the
        engine's foot-contact law (Hunt-Crossley normal force, Coulomb-capped viscous friction: csrc/go2_ws.cuh),
        evaluated on the state at the START of the physics step and held over it through the external-wrench
        input of the engine (explicit in time: at kn = 1e4 N/m and dt = 5 ms far inside the stability limit of the lightest
        link).  Returns the wrench [N,13,6] and leaves the per-body world forces in self._body_contact_force."""
        c = self.model.contact
        self._backend.body_states(self._state, out=self._rigid_body_state)
        rb = self._rigid_body_state[:, self._sphere_body]                             # [N,K,13]
        arm = self._rotate(rb[..., 3:7], self._sphere_offset.expand(self.num_envs, -1, -1))
        p = rb[..., 0:3] + arm
        v = rb[..., 7:10] + torch.cross(rb[..., 10:13], arm, dim=-1)
        depth = self._sphere_radius - p[..., 2]
        fn = torch.clamp(c.kn * depth * (1.0 - c.cn * v[..., 2]), min=0.0)
        fn = torch.where(depth > 0, fn, torch.zeros_like(fn))
        speed = torch.sqrt(v[..., 0] ** 2 + v[..., 1] ** 2 + c.veps ** 2)
        coef = torch.minimum(torch.full_like(fn, c.dt), c.mu * fn / speed)
        coef = torch.where(depth > 0, coef, torch.zeros_like(coef))
        f = torch.stack([-coef * v[..., 0], -coef * v[..., 1], fn], dim=-1)
        self._body_contact_force.zero_().index_add_(1, self._sphere_body, f)
        return self._wrench_from_point_forces(self._sphere_body, f, p)

    def simulate_at_each_physics_step(self):
        """Advance ONE physics step (sim_dt) under the applied torques, refresh dof_state only
        (isaacgym.py:622-626)."""
        if self._params_dirty:
            self._upload_params()
        flags = gm.FLAG_INERTIA_KEEP   # the per-env rows already carry the final inertia tensor: never rescale it
        kw = {}
        if self.body_spheres is not None:
            w = self._body_sphere_wrench()
            self._ext_wrench = w if self._ext_wrench is None else self._ext_wrench + w
        if self._ext_wrench is not None:          # consumed by this step only, like Isaac Gym's force tensors
            kw["ext_wrench"], self._ext_wrench = self._ext_wrench, None
        self._backend.sim_step(self._state, self._torques, 1, params=self._params, param_names=self.PARAM_NAMES,
                               flags=flags, foot_force=self._foot_force, **kw)
        self._refresh_dof()

    # ----- viewer ---------------------------------------------------------------------------------------------------
    def setup_viewer(self):
        raise NotImplementedError("the b200 backend is headless (SURVEY.md §2 row 5)")

    def render(self, sync_frame_time=True):
        return None

    def close(self):
        if self._backend is not None and hasattr(self._backend, "close"):
            self._backend.close()


class B200ActiveSysId(B200Sim):
    """Counterpart of IsaacGymActiveSysId (isaacgym_active_sysid.py:7-98): identical to B200Sim — the per-env
    `params_dict` overrides are handled in create_envs.  spigym/run_active_sysid.py:103-106 checks the class NAME
    `IsaacGymActiveSysId`; a b200 entrypoint relaxes that check to accept this one."""
