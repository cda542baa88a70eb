"""OracleSim — a CPU stand-in for the reference's simulator plugin, backed by the CPU oracle.

TEST INFRASTRUCTURE.  It implements the subset of spigym/simulator/base_simulator/base_simulator.py
(:113-150) and of the raw Isaac Gym handle surface (scripts/eval.py:185-214) that the reference's
replay path touches, so that the reference's OWN python (evaluate_batch, apply_base_mass,
mass_sweep, LeggedRobotBase._pre_physics_step/_physics_step/_compute_torques) can be executed
unmodified on top of it by tests/golden/make_golden.py.  Physics = oracle.sim_step (fp64).

The product counterpart with the same surface is spi_active_b200.simulator.B200Sim (CUDA).
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from oracle import oracle as orc
from spi_active_b200 import go2_model as gm


class _GymShim:
    """get/set_actor_rigid_body_properties + refresh_mass_matrix_tensors + destroy_sim."""

    def __init__(self, sim: "OracleSim"):
        self._s = sim

    def get_actor_rigid_body_properties(self, env_ptr, actor):
        masses = self._s.body_masses.copy()
        masses[0] = self._s.base_mass[int(env_ptr)]
        return [SimpleNamespace(mass=float(m)) for m in masses]

    def set_actor_rigid_body_properties(self, env_ptr, actor, props, recomputeInertia=True):
        ref = self._s.body_masses
        for i in range(1, len(ref)):
            if abs(props[i].mass - float(ref[i])) > 1e-6:
                raise NotImplementedError("only the base-link mass is a per-env parameter of this backend")
        self._s.base_mass[int(env_ptr)] = np.float32(props[0].mass)
        return True

    def refresh_mass_matrix_tensors(self, sim):
        return True

    def destroy_sim(self, sim):
        return None


class OracleSim:
    def __init__(self, num_envs: int, model: gm.Go2Model | None = None, device="cpu"):
        self.model = model or gm.go2_nominal()
        self.blob = gm.build_model_blob(self.model)
        self.num_envs, self.device = int(num_envs), device
        self.num_dof, self.num_bodies = 12, 19
        self.dof_names, self.body_names = list(gm.DOF_NAMES), list(gm.BODY_NAMES)
        self.sim_dt = self.model.dt
        N = self.num_envs
        # tensors the env layer reads directly (isaacgym.py:541-577); views share storage
        self.all_root_states = torch.zeros((N, 13), dtype=torch.float32)
        self.all_root_states[:, 2] = 0.34
        self.all_root_states[:, 6] = 1.0
        self.robot_root_states = self.all_root_states
        self.base_quat = self.robot_root_states[:, 3:7]
        self.dof_state = torch.zeros((N * 12, 2), dtype=torch.float32)
        self.dof_pos = self.dof_state.view(N, 12, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, 12, 2)[..., 1]
        self.dof_pos[:] = torch.tensor(self.model.q_default, dtype=torch.float32)
        self.torques = torch.zeros((N, 12), dtype=torch.float32)
        # internal (PhysX-side) state, fp64
        self._state = np.zeros((N, 37), dtype=np.float64)
        self._pull_all()
        # raw-handle shim (scripts/eval.py:185-214)
        self.envs = list(range(N))
        self.robot_handles = [0] * N
        self.sim = object()
        self.gym = _GymShim(self)
        self.body_masses = self.model.body_masses_isaac_order()
        self.base_mass = np.full(N, self.model.base.mass, dtype=np.float32)

    # ---- tensors <-> internal state ------------------------------------------------------------
    def _pull_all(self):
        self._state[:, 0:13] = self.all_root_states.numpy().astype(np.float64)
        self._state[:, 13:25] = self.dof_pos.numpy().astype(np.float64)
        self._state[:, 25:37] = self.dof_vel.numpy().astype(np.float64)

    def set_actor_root_state_tensor(self, env_ids, root_states):
        ids = env_ids.numpy()
        self.all_root_states[env_ids] = root_states[env_ids]
        self._state[ids, 0:13] = root_states[env_ids].numpy().astype(np.float64)

    def set_dof_state_tensor(self, env_ids, dof_state):
        ids = env_ids.numpy()
        ds = dof_state.view(self.num_envs, 12, 2)
        if ds.data_ptr() != self.dof_state.data_ptr():
            self.dof_state.view(self.num_envs, 12, 2)[env_ids] = ds[env_ids]
        self._state[ids, 13:25] = ds[env_ids, :, 0].numpy().astype(np.float64)
        self._state[ids, 25:37] = ds[env_ids, :, 1].numpy().astype(np.float64)

    def refresh_sim_tensors(self):
        self.all_root_states[:] = torch.from_numpy(self._state[:, 0:13].astype(np.float32))
        self._refresh_dof()

    def _refresh_dof(self):
        self.dof_pos[:] = torch.from_numpy(self._state[:, 13:25].astype(np.float32))
        self.dof_vel[:] = torch.from_numpy(self._state[:, 25:37].astype(np.float32))

    # ---- stepping (isaacgym.py:598-599, 622-626) ----------------------------------------------------
    def apply_torques_at_dof(self, torques):
        self.torques[:] = torques.reshape(self.num_envs, 12)

    def simulate_at_each_physics_step(self):
        self._state = orc.sim_step(self.blob, self._state, self.torques.numpy().astype(np.float64), 1,
                                   params=self.base_mass[:, None], param_ids=[gm.PARAM_IDS["mass"]])
        self._refresh_dof()
