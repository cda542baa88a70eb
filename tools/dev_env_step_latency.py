"""dev helper: latency of one spi_b200_env_step launch (the physics of a closed-loop control step) against the number of
physics steps it contains: slope = cost of a physics step, intercept = per-launch set-up (parameters, inertia, state I/O)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from spi_active_b200 import active as act, go2_model as gm
from spi_active_b200.engine import RolloutEngine

eng = RolloutEngine()
cfg = act.ActiveConfig(exploration_params=list(act.ActiveExploration.PARAM_ORDER))
for M in (341, 1024):
    ex = act.ActiveExploration(eng, act.PolicyMLP.random(eng.device, seed=0, gain=0.3), M, cfg)
    s0 = act.ActiveExploration.initial_main_states(M, ex.model, cfg).repeat_interleave(ex.param_dim + 1, 0).to(eng.device)
    a = torch.zeros(ex.num_envs, 12, device=eng.device)
    for with_params in (True, False):
        row = []
        for dec in (1, 2, 4, 8, 16):
            st = s0.clone()
            def run():
                eng.env_step(st, a, params=ex.params if with_params else None, param_names=ex.param_names if with_params else (),
                             decimation=dec, motor_model=cfg.motor_model, flags=gm.FLAG_HIP_HALF, zero_action_mask=ex.done)
            g = torch.cuda.CUDAGraph()
            run(); torch.cuda.synchronize()
            with torch.cuda.graph(g):
                run()
            st.copy_(s0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(50):
                g.replay()
            e1.record(); torch.cuda.synchronize()
            row.append(e0.elapsed_time(e1) / 50 * 1e3)
        print(f"envs={ex.num_envs} params={with_params}: us per launch at 1/2/4/8/16 physics steps:", np.round(row, 1))
