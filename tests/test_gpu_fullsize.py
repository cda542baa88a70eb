"""Full-size checks (BASELINE.json configs 2 and 4: S = 1730 windows of the `all` dataset, H = 5, 4096 ten-parameter
candidates per GPU, a 16384-candidate population) through size-independent properties — the oracle needs ~30 s per
4096 x 1730 evaluation on 16 cores, so at these sizes it only checks a random sample of the candidates.

Properties (all of the CUDA path, through the C-ABI):
  * candidates are independent: permuting the rows of `params` permutes the costs bit-exactly;
  * sharding: any contiguous split of the candidates over ranks gives bit-identical costs (what weak scaling relies on);
  * the masked mean over segments is the count-weighted mean of the means of any segment partition (fp32 tolerance);
  * a random sample of the 4096 candidates matches the fp64 oracle at the usual tolerance;
  * the counter-based population of a rank's shard is a bit-exact slice of the full 16384 population;
  * device elite selection / refit equals the host restatement on the full population;
  * sim-to-sim identifiability: on data recorded at the nominal mass the landscape's argmin is the nominal scale, with a
    Spearman rank correlation >= 0.99 between the CUDA landscape and the fp64 oracle's (the README's published bowl:
    tests/test_readme_bowl.py).
"""
import numpy as np
import pytest
import torch

from spi_active_b200 import cem
from spi_active_b200 import go2_model as gm
from spi_active_b200.dataset import SegmentBatch, pack_segments, to_device

import synth

pytestmark = pytest.mark.gpu

C_FULL, POP = 4096, 16384


@pytest.fixture(scope="module")
def full(engine):
    S, ds = synth.dataset("all", 5)
    assert S == 1730
    segs = pack_segments(to_device(ds, engine.device))
    cfg = cem.default_full_config(engine.model)
    f = lambda a: torch.tensor(np.asarray(a, np.float32), device=engine.device)
    pop = engine.cem_sample(f(cfg.mean), f(cfg.std), f(cfg.lo), f(cfg.hi), POP, 0, cfg.seed, 0)
    return S, ds, segs, cfg, pop


def _eval(engine, cfg, params, segs, **kw):
    return engine.evaluate_candidates(params, cfg.names, segs, motor_model=cfg.motor_model, flags=cfg.flags, **kw)


def test_full_size_permutation_and_sharding_are_bit_exact(engine, full):
    S, ds, segs, cfg, pop = full
    params = pop[:C_FULL].contiguous()
    cost, status = _eval(engine, cfg, params, segs, return_status=True)
    cost = cost.clone()
    assert int(status.sum()) == 0 and bool(torch.isfinite(cost).all())
    perm = torch.randperm(C_FULL, generator=torch.Generator().manual_seed(0)).to(engine.device)
    cost_p = _eval(engine, cfg, params[perm].contiguous(), segs)
    assert torch.equal(cost_p, cost[perm])
    for world in (2, 8):
        parts = [_eval(engine, cfg, params[a:b].contiguous(), segs).clone()
                 for a, b in (cem.shard_range(C_FULL, r, world) for r in range(world))]
        assert torch.equal(torch.cat(parts), cost)


def test_full_size_segment_partition(engine, full):
    S, ds, segs, cfg, pop = full
    params = pop[:256].contiguous()
    whole = _eval(engine, cfg, params, segs).cpu().numpy().astype(np.float64)
    acc, n_tot = 0.0, 0.0
    for a, b in ((0, 577), (577, 1200), (1200, 1730)):          # ragged, not multiples of 32
        mask = segs.seg_mask[a:b]
        n = float(mask.sum())
        part = SegmentBatch(segs.seg_init[a:b], segs.seg_actions[a:b], segs.seg_target[a:b], segs.seg_gains[a:b], mask, n)
        acc = acc + n * _eval(engine, cfg, params, part).cpu().numpy().astype(np.float64)
        n_tot += n
    assert n_tot == segs.cost_denominator
    np.testing.assert_allclose(acc / n_tot, whole, rtol=2e-6)


def test_full_size_sample_matches_oracle(engine, oracle_lib, blob, full):
    S, ds, segs, cfg, pop = full
    cost = _eval(engine, cfg, pop[:C_FULL].contiguous(), segs).cpu().numpy()
    pick = np.random.default_rng(1).choice(C_FULL, 8, replace=False)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    ref, st = oracle_lib.eval_candidates(blob, pop[:C_FULL].cpu().numpy()[pick], [gm.PARAM_IDS[n] for n in cfg.names],
                                         init, act, tgt, gains, mask, motor_model=gm.MOTOR_MODELS[cfg.motor_model],
                                         flags=cfg.flags, cost_denominator=denom)
    assert st.sum() == 0
    np.testing.assert_allclose(cost[pick], ref, rtol=5e-5, atol=2e-6)


@pytest.mark.parametrize("case", ["ties_at_threshold", "mostly_inf", "all_equal", "n_elite_1", "negative_and_nan"])
def test_elite_selection_edge_cases(engine, case):
    """spi_b200_cem_refit's O(C) radix select against the host rule (stable order by (cost, index), NaN / inf last) where the
    old O(C^2) rank kernel and a naive select differ: ties straddling the elite threshold, more diverged candidates than
    non-elites, all-equal costs, a single elite, negative costs and NaN."""
    rng = np.random.default_rng(11)
    C, P = 5000, 4
    params = rng.standard_normal((C, P)).astype(np.float32)
    cost = rng.random(C).astype(np.float32)
    n_elite = 250
    if case == "ties_at_threshold":
        cost = np.round(cost * 40).astype(np.float32) / 40            # ~125 candidates per distinct value
    elif case == "mostly_inf":
        cost[rng.choice(C, C - 100, replace=False)] = np.inf          # the threshold key is +inf itself
    elif case == "all_equal":
        cost[:] = 0.5
    elif case == "n_elite_1":
        n_elite = 1
    elif case == "negative_and_nan":
        cost = (cost - 0.5).astype(np.float32); cost[::7] = np.nan; cost[3] = -np.inf
    f = lambda a: torch.tensor(np.asarray(a, np.float32), device=engine.device)
    mean, std = f(np.zeros(P)), f(np.ones(P))
    best = engine.cem_refit(f(params), f(cost), n_elite, 0.7, mean, std, f(np.full(P, 1e-6)))
    m_ref, s_ref, b_ref, c_ref = cem.cem_refit_numpy(params, cost, n_elite, 0.7, np.zeros(P), np.ones(P), np.full(P, 1e-6))
    np.testing.assert_allclose(mean.cpu().numpy(), m_ref, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(std.cpu().numpy(), s_ref, rtol=2e-4, atol=2e-6)
    np.testing.assert_array_equal(best.cpu().numpy()[:-1], b_ref)
    assert float(best[-1]) == c_ref or (np.isnan(c_ref) and np.isnan(float(best[-1])))


def test_population_shards_are_slices_of_the_full_population(engine, full):
    S, ds, segs, cfg, pop = full
    f = lambda a: torch.tensor(np.asarray(a, np.float32), device=engine.device)
    for world in (2, 8):
        for r in (0, world - 1):
            c0, c1 = cem.shard_range(POP, r, world)
            part = engine.cem_sample(f(cfg.mean), f(cfg.std), f(cfg.lo), f(cfg.hi), c1 - c0, c0, cfg.seed, 0)
            assert torch.equal(part, pop[c0:c1])
    other = engine.cem_sample(f(cfg.mean), f(cfg.std), f(cfg.lo), f(cfg.hi), 64, 0, cfg.seed, 1)
    assert not torch.equal(other, pop[:64])                       # the iteration is part of the counter
    p = pop.cpu().numpy().astype(np.float64)
    assert (p >= np.asarray(cfg.lo) - 1e-6).all() and (p <= np.asarray(cfg.hi) + 1e-6).all()
    z = (p - np.asarray(cfg.mean)) / np.asarray(cfg.std)
    assert abs(z[:, 1:4].mean()) < 0.02 and abs(z[:, 1:4].std() - 1.0) < 0.02      # unclipped columns: N(0, 1)


def test_full_population_refit_matches_host_rule(engine, full):
    S, ds, segs, cfg, pop = full
    rng = np.random.default_rng(3)
    cost = rng.random(POP).astype(np.float32)
    cost[rng.choice(POP, 50, replace=False)] = np.inf                 # diverged candidates sort last
    cost[100:110] = cost[100]                                          # ties resolve by index
    n_elite = max(2, int(round(cfg.elite_frac * POP)))
    f = lambda a: torch.tensor(np.asarray(a, np.float32), device=engine.device)
    mean, std = f(cfg.mean), f(cfg.std)
    best = engine.cem_refit(pop, torch.from_numpy(cost).to(engine.device), n_elite, cfg.alpha, mean, std, f(cfg.std_floor))
    m_ref, s_ref, b_ref, c_ref = cem.cem_refit_numpy(pop.cpu().numpy(), cost, n_elite, cfg.alpha, np.asarray(cfg.mean),
                                                     np.asarray(cfg.std), np.asarray(cfg.std_floor))
    np.testing.assert_allclose(mean.cpu().numpy(), m_ref, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(std.cpu().numpy(), s_ref, rtol=2e-4, atol=1e-7)
    np.testing.assert_array_equal(best.cpu().numpy()[:-1], b_ref)
    assert float(best[-1]) == c_ref


def test_sim_to_sim_landscape_identifies_the_recorded_mass(engine, oracle_lib, blob, full):
    """scripts/mass_landscape.py on data recorded at the URDF mass, every window counted: the argmin of a sweep that contains
    scale 1.0 is scale 1.0, and the CUDA landscape ranks the 42 candidates like the fp64 ORACLE's (Spearman >= 0.99, BASELINE
    north_star; the comparison with the README's published bowl is tests/test_readme_bowl.py)."""
    S, ds, segs, cfg, pop = full
    scales = np.sort(np.append(np.linspace(0.5, 2.0, 41), 1.0))
    masses = (scales * engine.model.base.mass).astype(np.float32)
    w = np.array(cem.COST_WEIGHTS)
    tot = engine.evaluate_candidates(torch.from_numpy(masses)[:, None], ["mass"], segs).cpu().numpy().astype(np.float64) @ w
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    ref, st = oracle_lib.eval_candidates(blob, masses[:, None], [gm.PARAM_IDS["mass"]], init, act, tgt, gains, mask,
                                         cost_denominator=denom)
    ref = ref @ w
    assert st.sum() == 0
    assert scales[int(np.argmin(tot))] == 1.0 and scales[int(np.argmin(ref))] == 1.0
    ra, rb = np.argsort(np.argsort(tot)), np.argsort(np.argsort(ref))
    rho = 1.0 - 6.0 * ((ra - rb) ** 2).sum() / (len(ra) * (len(ra) ** 2 - 1))
    assert rho >= 0.99
    np.testing.assert_allclose(tot, ref, rtol=2e-4)
