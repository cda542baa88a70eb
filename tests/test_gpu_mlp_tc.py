"""spi_b200_policy_forward (tcgen05 3xTF32 actor MLP) against an fp64 evaluation of the reference's actor
(agents/modules/modules.py:47-63: Linear-ELU x3 + Linear; 900-512-256-128-12 per config/algo/ppo.yaml:32-40) and against
torch's fp32 evaluation on the GPU.  Tolerance: fp32 grade — the kernel must be as close to fp64 as torch's own fp32
GEMMs are (a few 1e-6 of the output scale); plain TF32 would sit at ~1e-3."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mlp64(x, ws, bs):
    h = x.astype(np.float64)
    for i, (w, b) in enumerate(zip(ws, bs)):
        h = h @ w.astype(np.float64).T + b.astype(np.float64)
        if i < len(ws) - 1:
            h = np.where(h > 0, h, np.expm1(np.minimum(h, 0.0)))
    return h


def _make(dims, seed, bias_scale=0.1):
    g = torch.Generator().manual_seed(seed)
    ws = [torch.randn(dims[i + 1], dims[i], generator=g) / np.sqrt(dims[i]) for i in range(4)]
    bs = [torch.randn(dims[i + 1], generator=g) * bias_scale for i in range(4)]
    return ws, bs


@pytest.mark.parametrize("dims,M", [((900, 512, 256, 128, 12), 11264), ((900, 512, 256, 128, 12), 77),
                                    ((45, 128, 128, 128, 3), 300), ((64, 256, 384, 128, 16), 129)])
def test_policy_forward_is_fp32_grade(dims, M):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from spi_active_b200.engine import TensorCorePolicy
    dev = torch.device("cuda:0")
    ws, bs = _make(dims, seed=sum(dims) + M)
    x = torch.randn(M, dims[0], generator=torch.Generator().manual_seed(1)) * 1.5
    x[:, :7] *= 30.0                       # a few large entries (clipped observations reach +-100)
    pol = TensorCorePolicy(ws, bs, dev)
    y = pol(x.to(dev))
    torch.cuda.synchronize()
    y = y.cpu().numpy()
    ref = _mlp64(x.numpy(), [w.numpy() for w in ws], [b.numpy() for b in bs])
    scale = np.abs(ref).max()
    err = np.abs(y - ref).max() / scale
    # torch fp32 on the same device, same inputs: the yardstick
    h = x.to(dev)
    for i in range(4):
        h = torch.addmm(bs[i].to(dev), h, ws[i].to(dev).t())
        if i < 3:
            h = torch.nn.functional.elu(h)
    err32 = np.abs(h.cpu().numpy() - ref).max() / scale
    assert err < 5e-6, (err, err32)
    assert err < 10 * err32 + 1e-6, (err, err32)


def test_policy_rejects_unsupported_shapes():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from spi_active_b200._lib import SpiB200Error
    from spi_active_b200.engine import TensorCorePolicy
    ws, bs = _make((900, 500, 256, 128, 12), 0)
    with pytest.raises(SpiB200Error):
        TensorCorePolicy(ws, bs, torch.device("cuda:0"))
    ws, bs = _make((900, 512, 256, 64, 12), 0)
    with pytest.raises(SpiB200Error):
        TensorCorePolicy(ws, bs, torch.device("cuda:0"))


def test_policy_forward_ring_permutes_the_first_layer():
    """spi_b200_policy_forward_ring: with col_map[r] a permutation of the input columns (and a few unused elements), the
    forward on the ring-ordered input at head position r equals the plain forward on the un-permuted input."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from spi_active_b200.engine import TensorCorePolicy
    dev = torch.device("cuda:0")
    dims, M, n_rot = (900, 512, 256, 128, 12), 300, 4
    ws, bs = _make(dims, seed=11)
    pol = TensorCorePolicy(ws, bs, dev)
    rng = np.random.default_rng(5)
    col_map = np.stack([rng.permutation(900) for _ in range(n_rot)]).astype(np.int32)
    col_map[2, :7] = -1                                    # rotation 2 ignores 7 ring elements (their columns get no input)
    pol.enable_ring(col_map)
    x = torch.randn(M, 900, generator=torch.Generator().manual_seed(3)) * 1.5
    rot = torch.zeros(1, dtype=torch.int32, device=dev)
    for r in range(n_rot):
        ring = torch.zeros(M, 900)
        valid = col_map[r] >= 0
        ring[:, valid] = x[:, col_map[r][valid]]           # ring element k carries input column col_map[r, k]
        ring[:, ~valid] = 123.0                            # must be ignored
        x_eff = torch.zeros(M, 900)
        x_eff[:, col_map[r][valid]] = x[:, col_map[r][valid]]
        hi, lo = pol.alloc_input(M)
        pol.split_input(ring.to(dev).contiguous(), hi, lo)
        rot.fill_(r)
        y = pol.forward_ring(hi, lo, M, rot)
        torch.cuda.synchronize()
        ref = _mlp64(x_eff.numpy(), [w.numpy() for w in ws], [b.numpy() for b in bs])
        err = np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err < 5e-6, (r, err)
