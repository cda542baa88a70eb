"""The one physics result the reference publishes (README.md:153-163, assets/optimization_results.png): the cost-vs-base-mass
bowl of `mass_opt.py --config all --horizon 5` and its identified mass 7.006 kg / cost 0.028378 (ground truth 6.921 kg).

This tool re-creates that experiment on the CPU oracle (fp64) for a grid of the engine's OWN modelling choices — the contact
stiffness kn, the Hunt-Crossley damping cn, the integrator sub-steps nsub and the SPI_FLAG_INERTIA_KEEP reading of
`recomputeInertia=True` — each time recording the four trajectories with that model (sim-to-sim, like the reference) and
scoring the seven published masses with the README's own masking (strict_reference, env_batch = 1730: only the 244 `jump`
windows before the first file boundary count, divided by 1726 — scripts/eval.py:279-280, 304-309).  Writes
profiles/readme_bowl.json; DESIGN.md §2 quotes it.  Test infrastructure (imports oracle/)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np

from oracle import oracle as orc
from spi_active_b200 import go2_model as gm, landscape, recorders
from spi_active_b200.dataset import concat_windows, pack_segments, to_device, window_recording

PUB_MASS = np.array([3.6, 4.8, 7.0, 10.2, 12.7, 13.7, 20.8])
PUB_COST = np.array([0.046, 0.036, 0.0284, 0.0366, 0.0475, 0.0516, 0.076])
README_BEST_MASS, README_BEST_COST, README_TRUE_MASS = 7.006, 0.028378, 6.921
W = np.array([10.0, 5.0, 1.0])


def spearman(a, b):
    ra, rb = np.argsort(np.argsort(a)), np.argsort(np.argsort(b))
    n = len(ra)
    return 1.0 - 6.0 * float(((ra - rb) ** 2).sum()) / (n * (n * n - 1))


def strict_dataset(model, strict=True):
    blob = gm.build_model_blob(model)

    def fn(init, actions):
        return orc.rollout_states(blob, np.array([[model.base.mass]], np.float32), [gm.PARAM_IDS["mass"]], init[None],
                                  actions[None])[0, 0]
    wins = [window_recording(recorders.record(n, fn, model), 5) for n in recorders.CONFIG_FILES["all"]]
    S, ds = concat_windows(wins)
    segs = pack_segments(to_device(ds, "cpu"), env_batch=S, strict_reference=strict)
    return blob, segs


def total_cost(blob, segs, masses, flags=0):
    cost, st = orc.eval_candidates(blob, np.asarray(masses, np.float32)[:, None], [gm.PARAM_IDS["mass"]],
                                   segs.seg_init.numpy(), segs.seg_actions.numpy(), segs.seg_target.numpy(),
                                   segs.seg_gains.numpy(), segs.seg_mask.numpy(), flags=flags,
                                   cost_denominator=segs.cost_denominator)
    return cost @ W


def one(kn, cn, nsub, keep, strict=True, tpe=False):
    model = gm.go2_nominal(gm.ContactParams(kn=kn, cn=cn, nsub=nsub))
    blob, segs = strict_dataset(model, strict)
    flags = gm.FLAG_INERTIA_KEEP if keep else 0
    tot = total_cost(blob, segs, PUB_MASS, flags)
    grid = np.linspace(0.5, 2.0, 151) * model.base.mass
    gtot = total_cost(blob, segs, grid, flags)
    row = dict(kn=kn, cn=cn, nsub=nsub, inertia_keep=bool(keep), strict_mask=bool(strict),
               cost_at_published_masses=[round(float(x), 5) for x in tot],
               ratio_to_published=[round(float(x), 3) for x in tot / PUB_COST],
               max_rel_err=round(float(np.abs(tot / PUB_COST - 1).max()), 3), spearman=round(spearman(tot, PUB_COST), 3),
               argmin_mass_151pt=round(float(grid[int(np.argmin(gtot))]), 3), min_cost=round(float(gtot.min()), 5))
    if tpe:
        obj = lambda s: float(total_cost(blob, segs, [s * model.base.mass], flags)[0])
        bs, bv, trials = landscape.optimize_mass(obj)
        row["tpe_50_trials"] = dict(best_mass=round(bs * model.base.mass, 3), best_cost=round(bv, 6),
                                    first_trial=[round(trials[0][0] * model.base.mass, 2), round(trials[0][1], 5)])
    return row


def main():
    rows = [one(10000.0, 0.5, 2, False, tpe=True)]
    for kn in (5000.0, 20000.0):
        rows.append(one(kn, 0.5, 2, False))
    for cn in (0.25, 1.0):
        rows.append(one(10000.0, cn, 2, False))
    for nsub in (1, 4):
        rows.append(one(10000.0, 0.5, nsub, False))
    rows.append(one(10000.0, 0.5, 2, True))
    rows.append(one(10000.0, 0.5, 2, False, strict=False))
    out = dict(published=dict(masses=PUB_MASS.tolist(), costs=PUB_COST.tolist(), best_mass=README_BEST_MASS,
                              best_cost=README_BEST_COST, true_mass=README_TRUE_MASS,
                              source="README.md:153-163, assets/optimization_results.png (read off the plot, +-0.001)"),
               rows=rows)
    (ROOT / "profiles" / "readme_bowl.json").write_text(json.dumps(out, indent=1))
    for r in rows:
        print(r)


if __name__ == "__main__":
    main()
