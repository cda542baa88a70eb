"""Cross-entropy-method SysID loop over candidate parameter vectors, candidates sharded over ranks.

This is the optimiser BASELINE.json config 4 names ("full inertial + motor-model CEM, 16384
candidates, H=5, sharded over 1/2/4/8 B200").  The reference ships no optimiser over the rigid-body
parameters (SURVEY.md headline fact 4): it names the parameters
(spigym/config/env/active_sysid_openloop.yaml:17-27) and scores one candidate at a time
(scripts/mass_landscape.py:123-126, scripts/mass_opt.py:136-169).  Here one iteration is

    sample (device, counter-based RNG keyed by the GLOBAL candidate index)
    -> fused rollout + cost reduction of this rank's shard          (spi_b200_eval_candidates)
    -> weighted total 10*pos + 5*quat + 1*joint                      (mass_landscape.py:32-36,162-164)
    -> all-gather of the per-candidate totals                        (NCCL over NVLink; gloo in CPU tests)
    -> elite top-k + refit of mean/std on every rank                 (spi_b200_cem_refit)

Every rank draws the SAME full population (the RNG is keyed by the global index), so only the
[C_local] cost slice crosses the wire; parameters never do.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import go2_model as gm

COST_WEIGHTS = (10.0, 5.0, 1.0)  # scripts/mass_landscape.py:32-36 / scripts/mass_opt.py:28-32

# BASELINE config 4: (mass, comx, comy, comz, Ixx, Iyy, Izz, hip_a, thigh_a, calf_a)
FULL_PARAM_NAMES = ["mass", "comx", "comy", "comz", "inertiax", "inertiay", "inertiaz",
                    "motor_model_hip_a", "motor_model_thigh_a", "motor_model_calf_a"]


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [c0, c1) of `total` candidates owned by `rank`.  `total` must
    divide evenly so that all_gather_into_tensor sees equal slices."""
    if total % world != 0:
        raise ValueError(f"candidate count {total} must be a multiple of the world size {world}")
    n = total // world
    return rank * n, (rank + 1) * n


def gather_costs(local: torch.Tensor, world: int, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All-gather the [C_local] cost slices into [C_total] in rank order.  One collective per
    iteration; works on CUDA tensors (nccl) and CPU tensors (gloo)."""
    if world == 1:
        return local
    import torch.distributed as dist
    if out is None:
        out = torch.empty((local.numel() * world,), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:  # gloo has no all_gather_into_tensor on every build: use the list form
        parts = list(out.view(world, -1).unbind(0))
        dist.all_gather(parts, local.contiguous(), group=group)
    return out


@dataclass
class CemConfig:
    names: List[str] = field(default_factory=lambda: list(FULL_PARAM_NAMES))
    mean: Sequence[float] = ()
    std: Sequence[float] = ()
    lo: Sequence[float] = ()
    hi: Sequence[float] = ()
    std_floor: Sequence[float] = ()
    elite_frac: float = 0.05
    alpha: float = 0.7
    weights: Sequence[float] = COST_WEIGHTS
    seed: int = 0
    motor_model: str = "act2tau_vec3_tanh"
    flags: int = 0


def default_full_config(model: Optional[gm.Go2Model] = None, seed: int = 0) -> CemConfig:
    """Prior of SURVEY.md §8(d) config 4: N(nominal, diag sigma), sigma = (1.5 kg, 0.03 m x3, 30 % of the
    URDF inertia x3, 3.0 N m x3), clamped to physically valid boxes."""
    nominal = gm.default_param_vector(model)
    ids = [gm.PARAM_IDS[n] for n in FULL_PARAM_NAMES]
    mean = nominal[ids].astype(np.float64)
    std = np.array([1.5, 0.03, 0.03, 0.03, 0.3 * mean[4], 0.3 * mean[5], 0.3 * mean[6], 3.0, 3.0, 3.0])
    lo = np.array([1.0, -0.15, -0.15, -0.15, 0.1 * mean[4], 0.1 * mean[5], 0.1 * mean[6], 5.0, 5.0, 5.0])
    hi = np.array([20.0, 0.15, 0.15, 0.15, 4.0 * mean[4], 4.0 * mean[5], 4.0 * mean[6], 60.0, 60.0, 60.0])
    floor = std * 1e-3
    return CemConfig(names=list(FULL_PARAM_NAMES), mean=mean.tolist(), std=std.tolist(), lo=lo.tolist(),
                     hi=hi.tolist(), std_floor=floor.tolist(), seed=seed)


class CemOptimizer:
    """Device-resident CEM state; `iterate()` runs one SysID iteration and never touches the host."""

    def __init__(self, engine, segs, cfg: CemConfig, total_candidates: int, rank: int = 0, world: int = 1,
                 group=None):
        self.eng, self.segs, self.cfg = engine, segs, cfg
        self.rank, self.world, self.group = rank, world, group
        self.C = int(total_candidates)
        self.c0, self.c1 = shard_range(self.C, rank, world)
        dev = engine.device
        f = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device=dev)
        self.mean, self.std = f(cfg.mean), f(cfg.std)
        self.lo, self.hi, self.std_floor = f(cfg.lo), f(cfg.hi), f(cfg.std_floor)
        P = self.mean.numel()
        assert len(cfg.names) == P
        self.n_elite = max(2, int(round(cfg.elite_frac * self.C)))
        # persistent buffers: nothing is allocated inside iterate()
        self.params = torch.empty((self.C, P), device=dev, dtype=torch.float32)
        self.cost3 = torch.empty((self.c1 - self.c0, 3), device=dev, dtype=torch.float32)
        self.total_local = torch.empty((self.c1 - self.c0,), device=dev, dtype=torch.float32)
        self.total = torch.empty((self.C,), device=dev, dtype=torch.float32) if world > 1 else self.total_local
        self.best = torch.empty((P + 1,), device=dev, dtype=torch.float32)
        self.iteration = 0

    def iterate(self) -> torch.Tensor:
        """One SysID iteration; returns best[P+1] = (best params, best cost) as a device tensor."""
        e, c = self.eng, self.cfg
        e.cem_sample(self.mean, self.std, self.lo, self.hi, self.C, 0, c.seed, self.iteration, out=self.params)
        e.evaluate_candidates(self.params[self.c0:self.c1], c.names, self.segs, motor_model=c.motor_model,
                              flags=c.flags, out=self.cost3)
        e.weighted_cost(self.cost3, c.weights, out=self.total_local)
        total = gather_costs(self.total_local, self.world, self.group, out=self.total if self.world > 1 else None)
        e.cem_refit(self.params, total, self.n_elite, c.alpha, self.mean, self.std, self.std_floor, out_best=self.best)
        self.iteration += 1
        return self.best


# ---- host restatement of the refit rule (used by the gloo CPU tests and by docs) -----------------------
def cem_refit_numpy(params: np.ndarray, cost: np.ndarray, n_elite: int, alpha: float, mean: np.ndarray,
                    std: np.ndarray, std_floor: Optional[np.ndarray] = None):
    """Same rule as spi_b200_cem_refit: stable rank by (cost, index), non-finite last; elite mean and
    population std; exponential smoothing; floor on std."""
    c = np.where(np.isfinite(cost), cost, np.inf).astype(np.float64)
    order = np.lexsort((np.arange(c.size), c))
    elite = params[order[:n_elite]].astype(np.float64)
    mu = elite.mean(axis=0)
    sd = np.sqrt(((elite - mu) ** 2).mean(axis=0))
    new_mean = (1 - alpha) * mean + alpha * mu
    new_std = (1 - alpha) * std + alpha * sd
    if std_floor is not None:
        new_std = np.maximum(new_std, std_floor)
    return new_mean, new_std, params[order[0]], float(cost[order[0]])
