"""ncu target for the active-exploration path (BASELINE config 5 shape): a few closed-loop control steps WITHOUT the
CUDA graph (so every kernel of a step shows up in the launch list) and one tensor-core FIM contraction.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_active.csv \
        python tools/profile_active.py [M] [steps]
"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from spi_active_b200 import active as act
from spi_active_b200.engine import RolloutEngine

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
eng = RolloutEngine()
cfg = act.ActiveConfig(exploration_params=list(act.ActiveExploration.PARAM_ORDER), fim_mode="tensor", fim_chunk=64)
ex = act.ActiveExploration(eng, act.PolicyMLP.random(eng.device, seed=0, gain=0.3), M, cfg)
rng = np.random.default_rng(0)
r = np.asarray(act.COMMAND_RANGES)
vals = rng.uniform(r[act.COMMAND_SAMPLING_IDXS, 0], r[act.COMMAND_SAMPLING_IDXS, 1], (M, 5, 3)).astype(np.float32)
cmds = torch.from_numpy(np.stack([act.expand_commands(act.commands_constant(v, 250)) for v in vals]))
out = ex.evaluate_policy(cmds, total_steps=steps, use_cuda_graph=False)
torch.cuda.synchronize()
print("steps", out["steps"], "reward mean", out["total_reward"].mean())
