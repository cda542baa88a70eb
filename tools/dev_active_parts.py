"""dev helper: which part of the closed-loop control step costs what once the pipelines overlap?

Captures, per pipeline, a CUDA graph of ONE part of the step (actor forward / physics / post-step / the whole step) and replays it
round-robin over the pipelines' streams, like PipelinedExploration does with the whole step: seconds per replay round for each part
alone.  If the parts overlapped perfectly the whole step would cost max(parts); if not at all, their sum.

    python tools/dev_active_parts.py [n_pipelines] [M]
"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from spi_active_b200 import active as act
from spi_active_b200 import go2_model as gm
from spi_active_b200.engine import RolloutEngine

n_pipe = int(sys.argv[1]) if len(sys.argv) > 1 else 3
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
REPS = 600
eng = RolloutEngine()
cfg = act.ActiveConfig(exploration_params=list(act.ActiveExploration.PARAM_ORDER))
pipe = act.PipelinedExploration(eng, act.PolicyMLP.random(eng.device, seed=0, gain=0.3), M, cfg, n_pipelines=n_pipe)
rng = np.random.default_rng(0)
r = np.asarray(act.COMMAND_RANGES)
vals = rng.uniform(r[act.COMMAND_SAMPLING_IDXS, 0], r[act.COMMAND_SAMPLING_IDXS, 1], (M, 5, 3)).astype(np.float32)
cmds = torch.from_numpy(np.stack([act.expand_commands(act.commands_constant(v, 250)) for v in vals])).pin_memory()
pipe.evaluate_policy(cmds, total_steps=64)          # warm-up: buffers, graphs, attributes


def part_fn(sub, part):
    def actor():
        return sub.tc_policy.forward_ring(sub.obs_hi, sub.obs_lo, sub.num_envs, sub.ctrl[3:4], out=sub.raw_actions)

    def physics():
        sub.backend.env_step(sub.state, sub.raw_actions, params=sub.params, param_names=sub.param_names,
                             motor_model=cfg.motor_model, flags=gm.FLAG_HIP_HALF, zero_action_mask=sub.done)

    def post():
        c = sub.cfg
        sub.backend.active_post_step(sub.state, sub.raw_actions, sub.done, sub.main_commands, sub.commands, sub.actions,
                                     sub.gait_indices, sub.clock, None, None, None, sub.hist, sub.live_hist, sub.dead_steps,
                                     sub.schedule, sub.counter, sub.ctrl, sub.dt, c.action_clip, act.CLIP_OBSERVATIONS,
                                     act.TERMINATION_GRAVITY, sub.model.q_default, obs_hi=sub.obs_hi, obs_lo=sub.obs_lo,
                                     ring_slots=act.RING_SLOTS, fim_jtj=sub.jtj, fim_trace=sub.trace_acc,
                                     fim_delta=float(c.delta_param))
    n_small = 30 * 32      # 30 CTAs instead of 117: the same launch latency on a quarter of the SMs

    def physics_small():
        sub.backend.env_step(sub.state[:n_small], sub.raw_actions[:n_small], params=sub.params[:n_small], param_names=sub.param_names,
                             motor_model=cfg.motor_model, flags=gm.FLAG_HIP_HALF, zero_action_mask=sub.done[:n_small])
    if part == "actor+physics30":
        return lambda: (actor(), physics_small())
    if part == "physics30":
        return physics_small
    return {"actor": actor, "physics": physics, "post": post, "all": sub._policy_step,
            "actor+physics": lambda: (actor(), physics()), "physics+post": lambda: (physics(), post())}[part]


init = act.ActiveExploration.initial_main_states(M, pipe.model, pipe.cfg)
for part in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("actor", "physics", "post", "actor+physics", "physics+post", "all")):
    graphs = []
    for i, (sub, sl) in enumerate(zip(pipe.subs, pipe.slices)):
        with torch.cuda.stream(pipe.streams[i]):
            sub.begin_rollout(cmds[sl], 1250, True, init[sl])
            fn = part_fn(sub, part)
            fn(); fn()
            g = torch.cuda.CUDAGraph()
            pipe.streams[i].synchronize()
            with torch.cuda.graph(g, stream=pipe.streams[i]):
                fn()
            graphs.append(g)
            sub.counter.zero_()
    torch.cuda.synchronize()
    for rep in range(2):
        for sub in pipe.subs:
            sub.counter.zero_()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(REPS):
            for i, g in enumerate(graphs):
                with torch.cuda.stream(pipe.streams[i]):
                    g.replay()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"pipelines={n_pipe} part={part:14s}: {dt / REPS * 1e6:7.1f} us per round of {n_pipe} replays (host loop {t_host / REPS * 1e6:6.1f} us)")
