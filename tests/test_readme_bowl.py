"""The reference's one published physics result (README.md:153-163 + assets/optimization_results.png): `mass_opt.py --config
all --horizon 5` identifies a base mass of 7.006 kg (ground truth 6.921 kg, +1.23 %) at a weighted cost of 0.028378, and the
plot gives seven (mass, cost) points of the bowl, read off at +-0.001:

    3.6 kg 0.046 | 4.8 kg 0.036 | 7.0 kg 0.0284 | 10.2 kg 0.0366 | 12.7 kg 0.0475 | 13.7 kg 0.0516 | 20.8 kg 0.076 (the 3.0x trial)

Isaac Gym itself cannot run here, so this is the only EXTERNAL anchor the rigid-body / contact model has.  The experiment is
re-created sim-to-sim exactly as the README does it: four recordings made with the nominal model, windows of H = 5, ONE chunk
of env_batch = 1730, and the reference's literal masking (scripts/eval.py:279-280: everything at or after the first file
boundary of the chunk is masked, so only the 244 `jump` windows count, divided by 1726 — `strict_reference=True`).

Stated tolerances (profiles/readme_bowl.json has the sensitivity table, tools/readme_bowl.py makes it):
  * pointwise: within 15 % of every published point (measured: -11 % ... +5 %);
  * rank: Spearman >= 0.96 against the published ranks.  The only inversion is the pair 3.6 kg / 12.7 kg, which the plot has
    3 % apart (0.046 vs 0.0475) at a read-off error of +-2 %; every other pair ranks as published;
  * identified mass: argmin within 5 % of the README's 7.006 kg (measured 6.71-6.75 kg: -4 % of 7.006, -3 % of the 6.921 kg truth;
    the README run itself sits +1.2 % off the truth — the bias comes from the one-step action shift of load_dataset, SURVEY D18,
    and is masking-dependent: with every window counted the argmin is the recorded mass exactly, next test);
  * minimum cost within 15 % of 0.028378 (measured 0.0251).
"""
import numpy as np
import pytest
import torch

from spi_active_b200 import go2_model as gm
from spi_active_b200 import landscape
from spi_active_b200.dataset import pack_segments, to_device

import synth

PUB_MASS = np.array([3.6, 4.8, 7.0, 10.2, 12.7, 13.7, 20.8])
PUB_COST = np.array([0.046, 0.036, 0.0284, 0.0366, 0.0475, 0.0516, 0.076])
README_BEST_MASS, README_BEST_COST, README_TRUE_MASS = 7.006, 0.028378, 6.921
W = np.array([landscape.COST_COEFF[k] for k in ("base_pos", "base_quat", "joint_pos")])


def spearman(a, b):
    ra, rb = np.argsort(np.argsort(a)), np.argsort(np.argsort(b))
    n = len(ra)
    return 1.0 - 6.0 * float(((ra - rb) ** 2).sum()) / (n * (n * n - 1))


@pytest.fixture(scope="module")
def strict_segs():
    S, ds = synth.dataset("all", 5)
    assert S == 1730
    segs = pack_segments(to_device(ds, "cpu"), env_batch=1730, strict_reference=True)
    assert int(segs.seg_mask.sum()) == 244 and segs.cost_denominator == 1726.0          # the README's effective sample
    return ds, segs


def oracle_total(orc, blob, segs, masses, flags=0):
    cost, st = orc.eval_candidates(blob, np.asarray(masses, np.float32)[:, None], [gm.PARAM_IDS["mass"]],
                                   segs.seg_init.numpy(), segs.seg_actions.numpy(), segs.seg_target.numpy(),
                                   segs.seg_gains.numpy(), segs.seg_mask.numpy(), flags=flags,
                                   cost_denominator=segs.cost_denominator)
    assert st.sum() == 0
    return cost @ W


def check_against_readme(tot_pub, grid_mass, grid_tot):
    rel = tot_pub / PUB_COST - 1.0
    assert np.abs(rel).max() < 0.15, rel
    assert spearman(tot_pub, PUB_COST) >= 0.96
    order_ours, order_pub = np.argsort(tot_pub), np.argsort(PUB_COST)
    swapped = {int(i) for i in np.nonzero(order_ours != order_pub)[0]}
    assert swapped <= {3, 4} and {PUB_MASS[order_pub[3]], PUB_MASS[order_pub[4]]} == {3.6, 12.7}      # the near-tie only
    best = float(grid_mass[int(np.argmin(grid_tot))])
    assert abs(best - README_BEST_MASS) / README_BEST_MASS < 0.05, best
    assert abs(best - README_TRUE_MASS) / README_TRUE_MASS < 0.035, best
    assert abs(float(grid_tot.min()) - README_BEST_COST) / README_BEST_COST < 0.15
    return best


def test_oracle_reproduces_the_published_bowl(oracle_lib, blob, strict_segs):
    ds, segs = strict_segs
    tot = oracle_total(oracle_lib, blob, segs, PUB_MASS)
    grid = np.linspace(0.5, 2.0, 151) * README_TRUE_MASS
    best = check_against_readme(tot, grid, oracle_total(oracle_lib, blob, segs, grid))
    assert 6.6 < best < 6.85
    # SPI_FLAG_INERTIA_KEEP (the other reading of recomputeInertia=True, D15) ranks all seven points as published
    tot_keep = oracle_total(oracle_lib, blob, segs, PUB_MASS, flags=gm.FLAG_INERTIA_KEEP)
    assert spearman(tot_keep, PUB_COST) == 1.0 and np.abs(tot_keep / PUB_COST - 1.0).max() < 0.15


def test_unmasked_sweep_identifies_the_recorded_mass_exactly(oracle_lib, blob, strict_segs):
    """Every window counted (the default, non-literal mask): the argmin of a sweep containing scale 1.0 is the recorded mass, and
    the costs are 3-6x the README's — the published numbers can only come from the literal chunk mask."""
    ds, _ = strict_segs
    segs = pack_segments(to_device(ds, "cpu"))
    scales = np.sort(np.append(np.linspace(0.5, 2.0, 41), 1.0))
    tot = oracle_total(oracle_lib, blob, segs, scales * README_TRUE_MASS)
    assert scales[int(np.argmin(tot))] == 1.0
    ratio = oracle_total(oracle_lib, blob, segs, PUB_MASS) / PUB_COST
    assert ratio.min() > 3.0 and ratio.max() < 6.5


def test_mass_opt_study_on_the_oracle_lands_on_the_bowl_minimum(oracle_lib, blob, strict_segs):
    """scripts/mass_opt.py:203-219: 50 TPE trials, the first enqueued at 3.0x (the 20.8 kg point of the plot)."""
    ds, segs = strict_segs
    trials = []

    def objective(scale):
        v = float(oracle_total(oracle_lib, blob, segs, [scale * README_TRUE_MASS])[0])
        trials.append((scale, v))
        return v
    best_scale, best_cost, _ = landscape.optimize_mass(objective)
    assert len(trials) == landscape.N_TRIALS and trials[0][0] == landscape.INITIAL_MASS_SCALE
    assert abs(trials[0][1] - PUB_COST[-1]) / PUB_COST[-1] < 0.10                        # the 3.0x trial: 0.0714 vs 0.076
    best_mass = best_scale * README_TRUE_MASS
    assert abs(best_mass - README_BEST_MASS) / README_BEST_MASS < 0.05, best_mass
    assert abs(best_cost - README_BEST_COST) / README_BEST_COST < 0.15


@pytest.mark.gpu
def test_cuda_engine_reproduces_the_published_bowl(engine, oracle_lib, blob, strict_segs):
    """The same experiment through the C-ABI (`spi_b200_eval_candidates`), against the README and against the fp64 oracle:
    the 20-point sweep of scripts/mass_landscape.py ranks exactly like the oracle (Spearman >= 0.99, BASELINE north_star)."""
    ds, segs_cpu = strict_segs
    segs = segs_cpu.to(engine.device)
    f = lambda m: torch.from_numpy(np.asarray(m, np.float32))[:, None]
    tot = engine.evaluate_candidates(f(PUB_MASS), ["mass"], segs).cpu().numpy().astype(np.float64) @ W
    grid = np.linspace(0.5, 2.0, 151) * README_TRUE_MASS
    gtot = engine.evaluate_candidates(f(grid), ["mass"], segs).cpu().numpy().astype(np.float64) @ W
    check_against_readme(tot, grid, gtot)
    np.testing.assert_allclose(tot, oracle_total(oracle_lib, blob, segs_cpu, PUB_MASS), rtol=2e-4)
    sweep = np.linspace(landscape.MASS_SCALE_MIN, landscape.MASS_SCALE_MAX, landscape.MASS_SAMPLES) * README_TRUE_MASS
    cuda = landscape.mass_sweep(engine, segs, np.array([README_TRUE_MASS]), sweep / README_TRUE_MASS).astype(np.float64) @ W
    ref = oracle_total(oracle_lib, blob, segs_cpu, sweep)
    assert spearman(cuda, ref) >= 0.99 and int(np.argmin(cuda)) == int(np.argmin(ref))
    np.testing.assert_allclose(cuda, ref, rtol=2e-4)
    # the study itself on the GPU: 51 single-candidate launches, identified mass vs the README's
    best_scale, best_cost, trials = landscape.optimize_mass(
        lambda s: landscape.evaluate_mass_scale(s, engine, segs, np.array([README_TRUE_MASS])))
    assert abs(best_scale * README_TRUE_MASS - README_BEST_MASS) / README_BEST_MASS < 0.05
    assert abs(best_cost - README_BEST_COST) / README_BEST_COST < 0.15
