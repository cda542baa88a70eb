"""spi_b200_policy_forward (tcgen05 fp16-pair actor MLP: hi.hi + hi.lo + lo.hi, fp32-grade) against an fp64 evaluation of the reference's actor
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
    ws, bs = _make((900, 512, 256, 128, 12), 0)
    ws[1][3, 5] = 300.0                                     # outside the fp16-pair range of the weights (+-255)
    with pytest.raises(SpiB200Error):
        TensorCorePolicy(ws, bs, torch.device("cuda:0"))
    ws[1][3, 5] = float("nan")
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


@pytest.mark.parametrize("knob", ["SPI_B200_MLP_PAIR", "SPI_B200_MLP_WIDE"])
def test_alternative_tile_shapes_keep_parity(knob):
    """The measured-but-not-default tile shapes of the actor kernel (cta_group::2 CTA pairs on 256 x 256 tiles; single-CTA
    128 x 256 tiles) are selected by an environment variable read once per process, so they are exercised in a child
    process: same fp32-grade parity as the default path."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import os, subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np, torch\n"
        "from test_gpu_mlp_tc import _make, _mlp64\n"
        "from spi_active_b200.engine import TensorCorePolicy\n"
        "dims, M = (900, 512, 256, 128, 12), 2048\n"
        "ws, bs = _make(dims, seed=7)\n"
        "x = torch.randn(M, 900, generator=torch.Generator().manual_seed(2)) * 1.5\n"
        "y = TensorCorePolicy(ws, bs, torch.device('cuda:0'))(x.cuda()).cpu().numpy()\n"
        "ref = _mlp64(x.numpy(), [w.numpy() for w in ws], [b.numpy() for b in bs])\n"
        "err = float(np.abs(y - ref).max() / np.abs(ref).max())\n"
        "print('ERR', err); assert err < 5e-6, err\n") % (str(root), str(root / "tests"))
    env = dict(os.environ, **{knob: "1"})
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ERR" in res.stdout, res.stdout + res.stderr


def test_policy_ring_argument_errors():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from spi_active_b200._lib import SpiB200Error
    from spi_active_b200.engine import TensorCorePolicy
    dev = torch.device("cuda:0")
    ws, bs = _make((900, 512, 256, 128, 12), 0)
    pol = TensorCorePolicy(ws, bs, dev)
    hi, lo = pol.alloc_input(128)
    rot = torch.zeros(1, dtype=torch.int32, device=dev)
    with pytest.raises(SpiB200Error):                       # no ring copies yet
        pol.forward_ring(hi, lo, 128, rot)
    bad = np.tile(np.arange(900, dtype=np.int32), (2, 1))
    bad[1, 5] = 900                                         # column index outside [-1, in)
    with pytest.raises(SpiB200Error):
        pol.enable_ring(bad)
    pol.enable_ring(np.tile(np.arange(900, dtype=np.int32), (2, 1)))
    rot.fill_(7)                                            # out-of-range head positions are clamped on the device, not read out of bounds
    y = pol.forward_ring(hi, lo, 128, rot)
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
