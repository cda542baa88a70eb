"""dev helper: time evaluate_policy of the active-exploration path at BASELINE config-5 shape
(M main envs x (P + 1), 1248 closed-loop steps, random-init 900-512-256-128-12 policy)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np, torch
from spi_active_b200 import active as act
from spi_active_b200.engine import RolloutEngine

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1250
eng = RolloutEngine()
mode = sys.argv[3] if len(sys.argv) > 3 else "auto"
cfg = act.ActiveConfig(exploration_params=list(act.ActiveExploration.PARAM_ORDER), fim_mode=mode)
ex = act.ActiveExploration(eng, act.PolicyMLP.random(eng.device, seed=0, gain=0.3), M, cfg)
rng = np.random.default_rng(0)
r = np.asarray(act.COMMAND_RANGES)
vals = rng.uniform(r[act.COMMAND_SAMPLING_IDXS, 0], r[act.COMMAND_SAMPLING_IDXS, 1], (M, 5, 3)).astype(np.float32)
cmds = torch.from_numpy(np.stack([act.expand_commands(act.commands_constant(v, 250)) for v in vals])).pin_memory()
for graph in (True, False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = ex.evaluate_policy(cmds, total_steps=steps, use_cuda_graph=graph)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if graph:   # second call: graph already captured
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = ex.evaluate_policy(cmds, total_steps=steps, use_cuda_graph=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    rew = out["total_reward"][::ex.param_dim + 1]
    print(f"fim={ex.fim_mode} graph={graph} M={M} P={ex.param_dim} envs={ex.num_envs} steps={out['steps']}: {dt:.3f} s  -> "
          f"{ex.num_envs * out['steps'] / dt:.3e} env-steps/s; reward mean {rew.mean():.4g} alive {(rew > 0).mean():.2f}")

# two independent pipelines on two streams (PipelinedExploration)
for n_pipe in (2, 3, 4, 6):
    pipe = act.PipelinedExploration(eng, act.PolicyMLP.random(eng.device, seed=0, gain=0.3), M, cfg, n_pipelines=n_pipe)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = pipe.evaluate_policy(cmds, total_steps=steps, use_cuda_graph=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    # is the host the limit?  time of the launch loop alone (advance_rollout calls return before the GPU is done)
    for i, (sub, sl) in enumerate(zip(pipe.subs, pipe.slices)):
        with torch.cuda.stream(pipe.streams[i]):
            nst = sub.begin_rollout(cmds[sl], steps, True, act.ActiveExploration.initial_main_states(M, pipe.model, pipe.cfg)[sl])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(nst):
        for i, sub in enumerate(pipe.subs):
            with torch.cuda.stream(pipe.streams[i]):
                sub.advance_rollout()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    for i, sub in enumerate(pipe.subs):
        with torch.cuda.stream(pipe.streams[i]):
            sub.finish_rollout()
    torch.cuda.synchronize()
    print(f"   launch loop {t_host:.3f} s of {t_all:.3f} s")
    rew = out["total_reward"][::pipe.param_dim + 1]
    print(f"pipelines={n_pipe} graph=True M={M} envs={pipe.num_envs} steps={out['steps']}: {dt:.3f} s  -> "
          f"{pipe.num_envs * out['steps'] / dt:.3e} env-steps/s; reward mean {rew.mean():.4g} alive {(rew > 0).mean():.2f}")
