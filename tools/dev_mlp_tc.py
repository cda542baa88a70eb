"""dev helper: tcgen05 3xTF32 actor MLP vs cuBLAS fp32 / TF32 at the config-5 shape (11264 x 900-512-256-128-12)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from spi_active_b200.engine import TensorCorePolicy
from spi_active_b200.active import PolicyMLP
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 11264
pm = PolicyMLP.random(dev, seed=0, gain=1.0)
x = torch.randn(M, 900, device=dev)
pol = TensorCorePolicy(pm.weights, pm.biases, dev)
xh, xl = pol.alloc_input(M)
pol.split_input(x, xh, xl)
out = torch.empty(M, 12, device=dev)
ref64 = x.double()
for i in range(4):
    ref64 = ref64 @ pm.weights[i].double().t() + pm.biases[i].double()
    if i < 3: ref64 = torch.nn.functional.elu(ref64)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
y = pol.forward_split(xh, xl, M, out)
torch.cuda.synchronize()
sc = ref64.abs().max()
print("tc  err", float((y.double() - ref64).abs().max() / sc), "ms", timeit(lambda: pol.forward_split(xh, xl, M, out)))
print("f32 err", float((pm(x).double() - ref64).abs().max() / sc), "ms", timeit(lambda: pm(x)))
torch.backends.cuda.matmul.allow_tf32 = True
print("tf32 err", float((pm(x).double() - ref64).abs().max() / sc), "ms", timeit(lambda: pm(x)))
print("flop", 2 * M * (900*512 + 512*256 + 256*128 + 128*12) / 1e9, "GFLOP")
