#!/usr/bin/env python
"""Mass parameter optimisation — same CLI and outputs as the reference's scripts/mass_opt.py (:38-59, :172-259):
50 TPE trials over mass_scale in [0.5, 2.0], first trial enqueued at 3.0, best re-evaluated for the breakdown."""
from __future__ import annotations

import argparse
from datetime import datetime

import numpy as np

from _common import add_common_args, build, landscape


def main() -> None:
    args = add_common_args(argparse.ArgumentParser(description="Mass parameter optimization")).parse_args()
    engine, segs, num_samples, ref_masses = build(args)
    timestamp = datetime.now().strftime("%Y%m%d-%H%M%S")
    run_dir = args.project_dir / f"mass_opt_{args.config}" / f"{timestamp}_h{args.horizon}"
    run_dir.mkdir(parents=True, exist_ok=True)
    base_nominal, total_nominal = float(ref_masses[0]), float(ref_masses.sum())
    print(f"Optimizing: {args.config}, {landscape.N_TRIALS} trials, horizon={args.horizon}")
    print(f"Nominal: {base_nominal:.3f} kg, Output: {run_dir}\n")
    objective = lambda mass_scale: landscape.evaluate_mass_scale(mass_scale, engine, segs, ref_masses)
    best_mass_scale, best_cost, trials = landscape.optimize_mass(objective, landscape.N_TRIALS, landscape.SEED)
    _, best_costs = landscape.evaluate_mass_scale(best_mass_scale, engine, segs, ref_masses, return_details=True)
    contrib = np.array([best_costs[0] * landscape.COST_COEFF["base_pos"], best_costs[1] * landscape.COST_COEFF["base_quat"],
                        best_costs[2] * landscape.COST_COEFF["joint_pos"]])
    pct = 100 * contrib / contrib.sum()
    best_mass = base_nominal * best_mass_scale
    print(f"\nOptimal: scale={best_mass_scale:.4f}, base={best_mass:.3f} kg, cost={best_cost:.6f}")
    print(f"Breakdown: pos {pct[0]:.0f}%, quat {pct[1]:.0f}%, joint {pct[2]:.0f}%")
    results_path = run_dir / "optimization_results.txt"
    landscape.write_optimization_results(results_path, args.config, args.horizon, len(trials), base_nominal, best_mass, best_cost)
    print(f"Results: {results_path}")


if __name__ == "__main__":
    main()
