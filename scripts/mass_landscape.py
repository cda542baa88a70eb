#!/usr/bin/env python
"""Mass sensitivity replay scan — same CLI and outputs as the reference's scripts/mass_landscape.py (:39-59, :131-205),
one fused GPU launch for the whole sweep instead of a sequential loop over candidates."""
from __future__ import annotations

import argparse
from datetime import datetime

import numpy as np

from _common import add_common_args, build, landscape


def main() -> None:
    args = add_common_args(argparse.ArgumentParser(description="Mass sensitivity replay scan")).parse_args()
    engine, segs, num_samples, ref_masses = build(args)
    timestamp = datetime.now().strftime("%Y%m%d-%H%M%S")
    run_dir = args.project_dir / f"{timestamp}_{args.config}_h{args.horizon}"
    run_dir.mkdir(parents=True, exist_ok=True)
    base_nominal, total_nominal = float(ref_masses[0]), float(ref_masses.sum())
    mass_scales = np.linspace(landscape.MASS_SCALE_MIN, landscape.MASS_SCALE_MAX, landscape.MASS_SAMPLES)
    print(f"Mass landscape: {args.config}, {landscape.MASS_SAMPLES} samples, horizon={args.horizon}")
    print(f"Nominal: {base_nominal:.3f} kg, Output: {run_dir}\n")
    costs = landscape.mass_sweep(engine, segs, ref_masses, mass_scales)
    results_path = run_dir / "landscape_results.txt"
    s = landscape.write_landscape_results(results_path, args.config, args.horizon, costs, mass_scales, base_nominal, total_nominal)
    print(f"\nOptimal: scale={s.best_scale:.4f}, base={s.best_base_mass:.3f} kg, cost={s.best_cost:.6f}")
    print(f"Breakdown: pos {s.cost_percentages[0]:.0f}%, quat {s.cost_percentages[1]:.0f}%, joint {s.cost_percentages[2]:.0f}%")
    try:  # the png of mass_landscape.py:75-108 needs matplotlib, which is optional here
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        fig, axes = plt.subplots(4, 1, figsize=(8, 11), sharex=True)
        masses = base_nominal * mass_scales
        for idx, label in enumerate(["Base position L2", "Base quaternion L2", "Joint position L2"]):
            axes[idx].plot(masses, costs[:, idx], marker="o", linewidth=1.5); axes[idx].set_ylabel(label)
        axes[3].plot(masses, s.total_costs, marker="o", linewidth=1.5, color="red"); axes[3].set_ylabel("Total Cost (Weighted)")
        axes[-1].set_xlabel("Base link mass (kg)")
        fig.savefig(run_dir / f"mass_sensitivity_h{args.horizon}.png", dpi=150)
    except ImportError:
        pass
    print(f"Results: {results_path}")


if __name__ == "__main__":
    main()
