"""Host side of the reference's two SysID entrypoints, on the fused engine.

Mirrors (same names, argument meaning, constants, output files):
  * scripts/mass_landscape.py:20-36   CONFIG_NAME / SEED / MASS_SCALE_MIN / MASS_SCALE_MAX / MASS_SAMPLES /
                                      MAX_SAFE_ENV_BATCH / COST_COEFF
  * scripts/mass_landscape.py:111-128 mass_sweep      -> ONE fused call with C = len(mass_scales) candidates
                                                         (the reference loops candidates sequentially)
  * scripts/mass_landscape.py:162-205 total cost, argmin, breakdown, landscape_results.txt schema
  * scripts/mass_opt.py:24-25,35-36   INITIAL_MASS_SCALE = 3.0, N_TRIALS = 50
  * scripts/mass_opt.py:62-76         compute_cost
  * scripts/mass_opt.py:136-169       evaluate_mass_scale
  * scripts/mass_opt.py:203-219       objective / TPE study with the first trial enqueued at 3.0 (quirk D7)
  * scripts/mass_opt.py:246-257       optimization_results.txt schema

Optuna is not installable here; `TpeSampler1D` is a small built-in Parzen-estimator sampler for the one
`mass_scale` dimension with the same ask/tell shape and the same defaults the reference relies on
(10 random startup trials, gamma = min(ceil(0.1 n), 25), 24 EI candidates).  When `optuna` IS importable the
CLI uses it, exactly like the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from pathlib import Path
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

SEED = 0
MASS_SCALE_MIN = 0.5
MASS_SCALE_MAX = 2.0
MASS_SAMPLES = 20
MAX_SAFE_ENV_BATCH = 8192
COST_COEFF = {"base_pos": 10.0, "base_quat": 5.0, "joint_pos": 1.0}
INITIAL_MASS_SCALE = 3.0
N_TRIALS = 50

# scripts/config/*.yaml: recording files per config name, in file order
DATA_FILES = {
    "all": ["go2_jump_data.npz", "go2_sine_data.npz", "go2_stand_data.npz", "go2_walk_data.npz"],
    "jump": ["go2_jump_data.npz"], "sine": ["go2_sine_data.npz"], "stand": ["go2_stand_data.npz"],
    "walk": ["go2_walk_data.npz"],
}
DATA_SUBDIR = Path("spigym") / "data" / "sysid_bag"


def load_config(config_name: str, data_root: Optional[Path] = None) -> List[Path]:
    """Data paths of a named config (scripts/mass_landscape.py:62-72).  The reference resolves them against
    the grand-parent of its checkout; here `data_root` says where `spigym/data/sysid_bag/` lives."""
    root = Path(data_root) if data_root is not None else Path.cwd()
    return [root / DATA_SUBDIR / f for f in DATA_FILES[config_name]]


def compute_cost(costs) -> float:
    """Weighted prediction cost (scripts/mass_opt.py:62-76)."""
    return (costs[0] * COST_COEFF["base_pos"] + costs[1] * COST_COEFF["base_quat"] + costs[2] * COST_COEFF["joint_pos"])


def mass_sweep(engine, segs, reference_masses: np.ndarray, mass_scales: Iterable[float], flags: int = 0) -> np.ndarray:
    """results[len(mass_scales), 3] = (base_pos, base_quat, joint_pos) mean errors per candidate base mass
    (scripts/mass_landscape.py:111-128).  One fused launch for the whole sweep."""
    import torch
    mass_scales = np.asarray(list(mass_scales), dtype=np.float32)
    base_nominal = float(reference_masses[0])
    params = torch.from_numpy((base_nominal * mass_scales.astype(np.float64)).astype(np.float32)[:, None])
    cost = engine.evaluate_candidates(params, ["mass"], segs, flags=flags)
    return cost.cpu().numpy().astype(np.float32)


def evaluate_mass_scale(mass_scale: float, engine, segs, reference_masses: np.ndarray, return_details: bool = False,
                        flags: int = 0):
    """One candidate (scripts/mass_opt.py:136-169)."""
    costs = mass_sweep(engine, segs, reference_masses, [mass_scale], flags)[0].astype(np.float64)
    total = compute_cost(costs)
    return (total, costs) if return_details else total


@dataclass
class LandscapeSummary:
    best_idx: int
    best_scale: float
    best_base_mass: float
    best_total_mass: float
    best_cost: float
    cost_percentages: np.ndarray
    total_costs: np.ndarray


def summarize_landscape(costs: np.ndarray, mass_scales: np.ndarray, base_nominal: float, total_nominal: float):
    """scripts/mass_landscape.py:161-182."""
    total_costs = (costs[:, 0] * COST_COEFF["base_pos"] + costs[:, 1] * COST_COEFF["base_quat"]
                   + costs[:, 2] * COST_COEFF["joint_pos"])
    best_idx = int(np.argmin(total_costs))
    base_masses = base_nominal * mass_scales
    contrib = np.array([costs[best_idx, 0] * COST_COEFF["base_pos"], costs[best_idx, 1] * COST_COEFF["base_quat"],
                        costs[best_idx, 2] * COST_COEFF["joint_pos"]])
    return LandscapeSummary(best_idx, float(mass_scales[best_idx]), float(base_masses[best_idx]),
                            float(base_masses[best_idx] + (total_nominal - base_nominal)), float(total_costs[best_idx]),
                            100 * contrib / contrib.sum(), total_costs)


def write_landscape_results(path: Path, config: str, horizon: int, costs: np.ndarray, mass_scales: np.ndarray,
                            base_nominal: float, total_nominal: float) -> LandscapeSummary:
    """landscape_results.txt in the reference's schema (scripts/mass_landscape.py:186-202)."""
    s = summarize_landscape(costs, mass_scales, base_nominal, total_nominal)
    base_masses = base_nominal * mass_scales
    with open(path, "w", encoding="utf-8") as f:
        f.write("Mass Landscape Results\n")
        f.write(f"{'=' * 60}\n")
        f.write(f"Config: {config}\n")
        f.write(f"Horizon: {horizon}\n")
        f.write(f"Cost coefficients: {COST_COEFF}\n")
        f.write(f"Number of samples: {len(mass_scales)}\n")
        f.write("\n")
        f.write(f"Nominal base mass: {base_nominal:.3f} kg\n")
        f.write(f"Optimal base mass: {s.best_base_mass:.3f} kg\n")
        f.write(f"Minimum total cost: {s.best_cost:.6f}\n")
        f.write("All samples:\n")
        f.write("scale,base_mass_kg,total_mass_kg,base_pos,base_quat,joint_pos,total_cost\n")
        for scale, base_mass, row, total_cost in zip(mass_scales, base_masses, costs, s.total_costs):
            total_mass = base_mass + (total_nominal - base_nominal)
            f.write(f"{scale:.4f},{base_mass:.3f},{total_mass:.3f},{row[0]:.6f},{row[1]:.6f},{row[2]:.6f},{total_cost:.6f}\n")
    return s


def write_optimization_results(path: Path, config: str, horizon: int, n_trials: int, base_nominal: float,
                               best_mass: float, best_cost: float) -> None:
    """optimization_results.txt in the reference's schema (scripts/mass_opt.py:246-257)."""
    with open(path, "w", encoding="utf-8") as f:
        f.write("Mass Optimization Results (Optuna)\n")
        f.write(f"{'=' * 60}\n")
        f.write(f"Config: {config}\n")
        f.write(f"Horizon: {horizon}\n")
        f.write(f"Cost coefficients: {COST_COEFF}\n")
        f.write(f"Number of trials: {n_trials}\n")
        f.write("\n")
        f.write(f"Nominal base mass: {base_nominal:.3f} kg\n")
        f.write(f"Optimal base mass: {best_mass:.3f} kg\n")
        f.write(f"Best cost: {best_cost:.6f}\n")


# ------------------------------------------------------------------------------------------------
# built-in 1-D TPE (stand-in for optuna.samplers.TPESampler when optuna is absent)
# ------------------------------------------------------------------------------------------------
class TpeSampler1D:
    """Tree-structured Parzen estimator over one bounded float, minimisation.

    ask() -> x; tell(x, value).  `enqueue(x)` forces the next ask (study.enqueue_trial).  Values told for
    points outside [low, high] (the reference's 3.0x first trial) take part in the good/bad split but are
    clipped into the range when they serve as kernel centres, as Optuna does for out-of-range params."""

    def __init__(self, low: float, high: float, seed: int = 0, n_startup_trials: int = 10, n_ei_candidates: int = 24):
        self.low, self.high = float(low), float(high)
        self.rng = np.random.RandomState(seed)
        self.n_startup, self.n_ei = n_startup_trials, n_ei_candidates
        self.xs: List[float] = []
        self.vals: List[float] = []
        self.queue: List[float] = []

    def enqueue(self, x: float) -> None:
        self.queue.append(float(x))

    def _parzen(self, centres: np.ndarray):
        """Truncated-Gaussian mixture with a uniform prior component; bandwidths from neighbour spacing."""
        mus = np.clip(np.asarray(centres, dtype=np.float64), self.low, self.high)
        prior_mu, prior_sigma = 0.5 * (self.low + self.high), self.high - self.low
        mus_all = np.append(mus, prior_mu)
        order = np.argsort(mus_all)
        sorted_mus = mus_all[order]
        ext = np.concatenate([[self.low], sorted_mus, [self.high]])
        sig_sorted = np.maximum(ext[1:-1] - ext[:-2], ext[2:] - ext[1:-1])
        sigmas = np.empty_like(sig_sorted)
        sigmas[order] = sig_sorted
        sigmas[-1] = prior_sigma
        n = len(mus_all)
        minsig = (self.high - self.low) / min(100.0, 1.0 + n)
        sigmas = np.clip(sigmas, minsig, self.high - self.low)
        weights = np.full(n, 1.0 / n)
        return mus_all, sigmas, weights

    def _logpdf(self, x: np.ndarray, mix) -> np.ndarray:
        mus, sigmas, w = mix
        from math import erf, sqrt
        z = (x[:, None] - mus[None]) / sigmas[None]
        cdf = lambda t: 0.5 * (1.0 + np.vectorize(erf)(t / sqrt(2.0)))
        norm = cdf((self.high - mus) / sigmas) - cdf((self.low - mus) / sigmas)
        comp = np.exp(-0.5 * z * z) / (sigmas[None] * math.sqrt(2 * math.pi) * norm[None])
        return np.log(np.maximum((comp * w[None]).sum(axis=1), 1e-300))

    def _sample(self, mix, n: int) -> np.ndarray:
        mus, sigmas, w = mix
        out = np.empty(n)
        for i in range(n):
            k = self.rng.choice(len(mus), p=w)
            while True:
                x = self.rng.normal(mus[k], sigmas[k])
                if self.low <= x <= self.high:
                    break
            out[i] = x
        return out

    def ask(self) -> float:
        if self.queue:
            return self.queue.pop(0)
        n = len(self.xs)
        if n < self.n_startup:
            return float(self.rng.uniform(self.low, self.high))
        n_good = min(int(math.ceil(0.1 * n)), 25)
        order = np.argsort(self.vals, kind="stable")
        xs = np.asarray(self.xs)
        good, bad = xs[order[:n_good]], xs[order[n_good:]]
        l, g = self._parzen(good), self._parzen(bad)
        cand = self._sample(l, self.n_ei)
        score = self._logpdf(cand, l) - self._logpdf(cand, g)
        return float(cand[int(np.argmax(score))])

    def tell(self, x: float, value: float) -> None:
        self.xs.append(float(x))
        self.vals.append(float(value) if np.isfinite(value) else float("inf"))

    @property
    def best(self) -> Tuple[float, float]:
        i = int(np.argmin(self.vals))
        return self.xs[i], self.vals[i]


def optimize_mass(objective, n_trials: int = N_TRIALS, seed: int = SEED) -> Tuple[float, float, List[Tuple[float, float]]]:
    """The study of scripts/mass_opt.py:215-219: TPE(seed), first trial enqueued at INITIAL_MASS_SCALE,
    n_trials evaluations.  Returns (best_mass_scale, best_cost, trials)."""
    try:
        import optuna  # pragma: no cover - absent in this image
        optuna.logging.set_verbosity(optuna.logging.WARNING)
        study = optuna.create_study(direction="minimize", sampler=optuna.samplers.TPESampler(seed=seed))
        study.enqueue_trial({"mass_scale": INITIAL_MASS_SCALE})
        study.optimize(lambda t: objective(t.suggest_float("mass_scale", MASS_SCALE_MIN, MASS_SCALE_MAX)),
                       n_trials=n_trials, show_progress_bar=False)
        trials = [(t.params["mass_scale"], t.value) for t in study.trials]
        return study.best_params["mass_scale"], study.best_value, trials
    except ImportError:
        pass
    sampler = TpeSampler1D(MASS_SCALE_MIN, MASS_SCALE_MAX, seed=seed)
    sampler.enqueue(INITIAL_MASS_SCALE)
    trials = []
    for _ in range(n_trials):
        x = sampler.ask()
        v = float(objective(x))
        sampler.tell(x, v)
        trials.append((x, v))
    bx, bv = sampler.best
    return bx, bv, trials
