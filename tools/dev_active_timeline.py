"""dev helper: GPU timeline of the pipelined closed-loop step.  Each pipeline captures K control steps into one CUDA graph with
EXTERNAL events (cudaEventRecordExternal nodes) between the parts of a step (actor | physics | tick + post-step); the graphs are
replayed round-robin like PipelinedExploration does and the events of the last replay are read back against a common origin.
(The event nodes turn the programmatic edges between the kernels into full dependencies: the timeline is ~2 % slower than the
product's step.)     python tools/dev_active_timeline.py [n_pipelines] [K]
"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from spi_active_b200 import active as act
from spi_active_b200 import go2_model as gm
from spi_active_b200.engine import RolloutEngine

n_pipe = int(sys.argv[1]) if len(sys.argv) > 1 else 3
K = int(sys.argv[2]) if len(sys.argv) > 2 else 24
M = 1024
eng = RolloutEngine()
cfg = act.ActiveConfig(exploration_params=list(act.ActiveExploration.PARAM_ORDER))
pipe = act.PipelinedExploration(eng, act.PolicyMLP.random(eng.device, seed=0, gain=0.3), M, cfg, n_pipelines=n_pipe)
rng = np.random.default_rng(0)
r = np.asarray(act.COMMAND_RANGES)
vals = rng.uniform(r[act.COMMAND_SAMPLING_IDXS, 0], r[act.COMMAND_SAMPLING_IDXS, 1], (M, 5, 3)).astype(np.float32)
cmds = torch.from_numpy(np.stack([act.expand_commands(act.commands_constant(v, 250)) for v in vals])).pin_memory()
pipe.evaluate_policy(cmds, total_steps=64)
init = act.ActiveExploration.initial_main_states(M, pipe.model, pipe.cfg)
ext = lambda: torch.cuda.Event(enable_timing=True, external=True)
origin = ext()
graphs, events = [], []
for i, (sub, sl) in enumerate(zip(pipe.subs, pipe.slices)):
    st = pipe.streams[i]
    with torch.cuda.stream(st):
        sub.begin_rollout(cmds[sl], 1250, True, init[sl])
        c = sub.cfg

        def parts():
            raw = sub.tc_policy.forward_ring(sub.obs_hi, sub.obs_lo, sub.num_envs, sub.ctrl[3:4], out=sub.raw_actions)
            yield
            sub.backend.env_step(sub.state, raw, params=sub.params, param_names=sub.param_names, motor_model=c.motor_model,
                                 flags=gm.FLAG_HIP_HALF, zero_action_mask=sub.done)
            yield
            sub.backend.active_post_step(sub.state, raw, sub.done, sub.main_commands, sub.commands, sub.actions,
                                         sub.gait_indices, sub.clock, None, None, None, sub.hist, sub.live_hist, sub.dead_steps,
                                         sub.schedule, sub.counter, sub.ctrl, sub.dt, c.action_clip, act.CLIP_OBSERVATIONS,
                                         act.TERMINATION_GRAVITY, sub.model.q_default, obs_hi=sub.obs_hi, obs_lo=sub.obs_lo,
                                         ring_slots=act.RING_SLOTS, fim_jtj=sub.jtj, fim_trace=sub.trace_acc,
                                         fim_delta=float(c.delta_param))
            yield
        for _ in parts():
            pass
        st.synchronize()
        ev = [[ext() for _ in range(4)] for _ in range(K)]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            if i == 0:
                origin.record(st)
            for k in range(K):
                ev[k][0].record(st)
                for j, _ in enumerate(parts()):
                    ev[k][j + 1].record(st)
        graphs.append(g); events.append(ev)
torch.cuda.synchronize()
for rep in range(6):
    for sub in pipe.subs:
        sub.counter.zero_()
    torch.cuda.synchronize()
    for i, g in enumerate(graphs):
        with torch.cuda.stream(pipe.streams[i]):
            g.replay()
    torch.cuda.synchronize()
t = np.array([[[origin.elapsed_time(e) * 1e3 for e in step] for step in ev] for ev in events])     # [pipe, K, 4] us
print(f"{n_pipe} pipelines x {K} steps in one replay each; times in us from the origin; the second half of the steps is analysed")
t = t[:, K // 2:, :]; K = t.shape[1]
for i in range(n_pipe):
    for k in range(min(K, 4)):
        a, b, c_, d = t[i, k]
        print(f"pipe {i} step {k}: start {a:7.1f} | actor {b - a:6.1f} | physics {c_ - b:6.1f} | tick+post {d - c_:6.1f} | step {d - a:6.1f}")
span = t[:, :, 3].max() - t[:, :, 0].min()
print(f"all {n_pipe * K} pipeline-steps in {span:.1f} us -> {span / K:.1f} us per control step of the whole population")
print("mean part durations (us): actor %.1f physics %.1f tick+post %.1f" % tuple((t[:, :, j + 1] - t[:, :, j]).mean() for j in range(3)))
# busy intervals: how much of the span has 1, 2, 3 parts in flight, and which
marks = []
for i in range(n_pipe):
    for k in range(K):
        for j, name in enumerate(("A", "P", "S")):
            marks.append((t[i, k, j], +1, name)); marks.append((t[i, k, j + 1], -1, name))
marks.sort()
cur = {"A": 0, "P": 0, "S": 0}; last = marks[0][0]; acc = {}
for tm, d, name in marks:
    key = "".join(n * cur[n] for n in "APS") or "-"
    acc[key] = acc.get(key, 0.0) + (tm - last); last = tm
    cur[name] += d
tot = sum(acc.values())
print("share of the span by what is in flight (A actor, P physics, S post-step):",
      {k: round(v / tot, 3) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])})
