"""spi_b200_fim_contract (tcgen05 tensor-core J J^T over whole rollouts) against an fp64 restatement of
active_sysid_openloop.py:402-426 summed over the steps as active_sysid.py:567-590 does, and against the per-step
CUDA-core kernel spi_b200_fim_reward.  Tolerance: 3xTF32 products are fp32-accurate (error term lo*lo ~ 2^-22), so the
contraction must match fp64 to 5e-6 relative to the largest entry of each env's matrix — 100x tighter than plain TF32
(2^-11) could ever meet."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from spi_active_b200.engine import RolloutEngine
    return RolloutEngine()


def reference(hist, live, delta):
    h = hist.astype(np.float64)
    J = (h[:, :, 0:1, :] - h[:, :, 1:, :]) / delta                 # [T,M,P,25]
    if live is not None:
        J = J * live.astype(np.float64)[:, :, None, None]
    return np.einsum("tmpd,tmqd->mpq", J, J)


def make(T, M, P, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    main = rng.standard_normal((T, M, 1, 25))
    hist = main + 0.05 * scale * rng.standard_normal((T, M, P + 1, 25))
    hist[:, :, 0:1, :] = main
    live = rng.random((T, M)) > 0.2
    return hist.astype(np.float32), live


@pytest.mark.parametrize("T,M,P", [(1, 1, 1), (2, 8, 10), (7, 11, 10), (64, 19, 16), (301, 40, 10), (1248, 9, 3)])
def test_fim_contract_matches_fp64(eng, T, M, P):
    hist, live = make(T, M, P, seed=T * 1000 + M * 10 + P)
    delta = 0.1
    ref = reference(hist, live, delta)
    jtj, trace = eng.fim_contract(torch.from_numpy(hist), delta, live=torch.from_numpy(live))
    torch.cuda.synchronize()
    jtj, trace = jtj.cpu().numpy(), trace.cpu().numpy()
    scale = np.abs(ref).reshape(M, -1).max(axis=1)[:, None, None] + 1e-30
    err = np.abs(jtj - ref) / scale
    assert err.max() < 5e-6, (err.max(), np.unravel_index(err.argmax(), err.shape), jtj[0, :2, :2], ref[0, :2, :2])
    np.testing.assert_allclose(trace, np.trace(ref, axis1=1, axis2=2), rtol=5e-6)
    # symmetric by construction of the 3 products (hi*hi + hi*lo + lo*hi)
    np.testing.assert_allclose(jtj, jtj.transpose(0, 2, 1), rtol=0, atol=2e-6 * scale.max())


def test_fim_contract_no_mask_and_accumulate(eng):
    hist, _ = make(33, 24, 10, seed=5)
    delta = 0.1
    ref = reference(hist, None, delta)
    h = torch.from_numpy(hist).cuda()
    jtj, trace = eng.fim_contract(h[:20], delta)
    eng.fim_contract(h[20:], delta, out_JtJ=jtj, out_trace=trace, accumulate=True)
    torch.cuda.synchronize()
    scale = np.abs(ref).max()
    assert np.abs(jtj.cpu().numpy() - ref).max() / scale < 5e-6
    np.testing.assert_allclose(trace.cpu().numpy(), np.trace(ref, axis1=1, axis2=2), rtol=5e-6)


def test_fim_contract_matches_per_step_kernel_and_ignores_dead_nans(eng):
    """Same numbers as T launches of the CUDA-core per-step kernel; rows of dead groups may hold NaN/Inf (diverged
    physics) and must not leak into the sum."""
    T, M, P = 40, 16, 10
    hist, live = make(T, M, P, seed=9)
    delta = 0.1
    hist_bad = hist.copy()
    hist_bad[~live] = np.nan
    h = torch.from_numpy(hist).cuda()
    acc = torch.zeros(M, P, P, device="cuda")
    for t in range(T):
        j, _ = eng.fim_reward(h[t], delta)
        acc += j * torch.from_numpy(live[t]).cuda()[:, None, None]
    jtj, trace = eng.fim_contract(torch.from_numpy(hist_bad), delta, live=torch.from_numpy(live))
    torch.cuda.synchronize()
    assert torch.isfinite(jtj).all() and torch.isfinite(trace).all()
    scale = float(acc.abs().max())
    assert float((jtj - acc).abs().max()) / scale < 1e-5


def test_fim_contract_argument_errors(eng):
    from spi_active_b200._lib import SpiB200Error
    hist = torch.zeros(2, 3, 18, 25)          # P = 17 > 16 slots
    with pytest.raises(SpiB200Error):
        eng.fim_contract(hist, 0.1)
    with pytest.raises(SpiB200Error):
        eng.fim_contract(torch.zeros(2, 3, 4, 25), 0.0)
