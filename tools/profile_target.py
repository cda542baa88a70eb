"""ncu target: a few launches of the fused rollout kernel at the bench shape (C candidates x S=1730 x H=5,
10 parameters, act2tau_vec3_tanh).  Usage (on the GPU box, see profiles/README.md):

    ncu --set full --clock-control none --import-source on -k regex:rollout -s 2 -c 1 \
        -o gpurun_out/prof_rollout python tools/profile_target.py [C]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from spi_active_b200 import cem, recorders  # noqa: E402
from spi_active_b200.dataset import pack_segments, to_device  # noqa: E402
from spi_active_b200.engine import RolloutEngine  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
eng = RolloutEngine()
S, ds = bench.build_dataset(recorders.engine_rollout_fn(eng), eng.model)
segs = pack_segments(to_device(ds, eng.device))
cfg = cem.default_full_config(eng.model)
opt = cem.CemOptimizer(eng, segs, cfg, C)
for _ in range(reps):
    opt.iterate()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
opt.iterate()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"C={C} S={S} iteration {ms:.3f} ms  -> {C * S * 5 / ms * 1e3:.4e} candidate-env steps/s")
