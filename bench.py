#!/usr/bin/env python
"""bench.py — headline benchmark of the SysID hot path (BASELINE.json metric: candidate-env steps/s and
seconds per SysID iteration, Go2, H = 5, 4096 candidates per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one full SysID (CEM) iteration over the synthetic `all` dataset (S = 1730 windows of the
four recorded-shape trajectories, H = 5): device-side sampling of the 10-parameter candidates ->
fused rollout + cost kernel -> weighted cost -> all-gather of the costs (N > 1) -> elite refit.
Weak scaling: 4096 candidates per GPU.  The timed region holds every kernel of the iteration plus an
L2 flush (256 MiB memset) between iterations.

JSON keys beyond the base contract: `roofline` (FP32 FMA pipe is the bound of this path — see
DESIGN.md §6; the HBM figure is reported alongside), `cpu_baseline` (the CPU oracle in fp32 on all
host cores, bounded sample), `e2e` (the same metric through the host-pointer C-ABI call
spi_b200_eval_candidates_host: H2D of candidates + dataset and D2H of the costs inside the timed
region), `gpu_launches`, `clocks`.

`--workload active` (optional, NOT the headline): the active-exploration rollout of BASELINE config 5 — one "step" = one
evaluate_policy call of 1024 command trajectories per GPU x 11 envs x 1249 closed-loop control steps; reports env steps/s.

`--impl reference`: the reference's own implementation of this path is Isaac Gym (closed source, not
installable: DESIGN.md §5), so this arm times the CPU restatement (oracle/, fp32, all host threads)
on bounded samples of the same workload.  This and the cpu_baseline leg are the only places bench.py
touches oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "candidate_env_steps_per_s"
UNIT = "candidate-env steps/s"
HORIZON = 5
CANDIDATES_PER_GPU = 4096
DATA_CONFIG = "all"


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0, period_ms: int = 100):
        self.rows, self.proc, self.thread = [], None, None
        self.gpu_index, self.period_ms = gpu_index, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i",
                 str(self.gpu_index), "-lms", str(self.period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(self.period_ms / 1000.0 * 1.5)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_max, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.3] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); sm_max.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(sm_max) if sm_max else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def build_dataset(rollout_fn, model):
    """The four recorded-shape trajectories of scripts/config/all.yaml (jump, sine, stand, walk), windowed
    with H = 5 exactly like scripts/eval.py:101-171 -> numpy dataset dict, S = 1730."""
    from spi_active_b200 import recorders
    from spi_active_b200.dataset import concat_windows, window_recording
    wins = [window_recording(recorders.record(n, rollout_fn, model), HORIZON) for n in recorders.CONFIG_FILES[DATA_CONFIG]]
    return concat_windows(wins)


def pack_numpy(ds):
    init = np.concatenate([ds["init_base_pos"], ds["init_base_ori"], ds["init_base_lin_vel"], ds["init_base_ang_vel"],
                           ds["init_joint_pos"], ds["init_joint_vel"]], axis=1).astype(np.float32)
    tgt = np.concatenate([ds["target_base_pos"], ds["target_base_ori"], ds["target_joint_pos"]], axis=1).astype(np.float32)
    gains = np.concatenate([ds["pd_gain_kp"], ds["pd_gain_kd"]], axis=1).astype(np.float32)
    mask = (~ds["motion_ends"]).astype(np.uint8)
    return init, ds["action_sequences"].astype(np.float32), tgt, gains, mask, float(mask.sum())


def roofline_constants():
    p = ROOT / "profiles" / "roofline.json"
    rec = json.loads(p.read_text())
    return rec


def ncu_executed_flops_per_rollout():
    """FP32 operations the rollout kernel EXECUTES per (candidate, segment) rollout — 2 x FFMA + FMUL + FADD thread-level counts
    of the committed ncu capture divided by its rollouts (grid x 32) — or None.  The oracle's op count (the `achieved` figure)
    includes work the kernel never does (zeros of the joint-origin sparsity, the constant calf projection, divisions)."""
    p = ROOT / "profiles" / "ncu_rollout_summary.json"
    try:
        rec = json.loads(p.read_text())
        rollouts = rec.get("rollouts_per_launch") or rec["grid_size"] * 32.0 * 1730.0 / (55.0 * 32.0)   # (older captures: padded grid)
        return rec["executed_fp32_flop_per_launch"] / rollouts
    except Exception:
        return None


def ncu_traffic():
    """Per-launch DRAM bytes of the rollout kernel from the committed `ncu --set full` summary, or None."""
    p = ROOT / "profiles" / "ncu_rollout_summary.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
# CPU arms (oracle): cpu_baseline leg and --impl reference
# ------------------------------------------------------------------------------------------------
def oracle_dataset():
    from oracle import oracle as orc
    from spi_active_b200 import go2_model as gm
    model = gm.go2_nominal()
    blob = gm.build_model_blob(model)

    def fn(init, actions):
        st = orc.rollout_states(blob, np.array([[model.base.mass]], np.float32), [gm.PARAM_IDS["mass"]], init[None],
                                actions[None], precision=32)
        return st[0, 0]
    return blob, build_dataset(fn, model)


def cpu_population(cfg, n, seed=0):
    rng = np.random.default_rng(seed)
    p = np.asarray(cfg.mean) + rng.standard_normal((n, len(cfg.mean))) * np.asarray(cfg.std)
    return np.clip(p, cfg.lo, cfg.hi).astype(np.float32)


def time_oracle(blob, packed, cfg, n_cand, threads=0, seed=0):
    from oracle import oracle as orc
    from spi_active_b200 import go2_model as gm
    init, act, tgt, gains, mask, denom = packed
    params = cpu_population(cfg, n_cand, seed)
    ids = [gm.PARAM_IDS[n] for n in cfg.names]
    t0 = time.perf_counter()
    cost, status = orc.eval_candidates(blob, params, ids, init, act, tgt, gains, mask, decimation=4,
                                       motor_model=gm.MOTOR_MODELS[cfg.motor_model], flags=cfg.flags,
                                       cost_denominator=denom, precision=32, n_threads=threads)
    dt = time.perf_counter() - t0
    assert np.isfinite(cost[status == 0]).all()
    return dt, n_cand * init.shape[0] * HORIZON


def calibrate_oracle(blob, packed, cfg, target_s):
    """Pick a candidate count whose oracle run takes about target_s seconds on this host."""
    dt, units = time_oracle(blob, packed, cfg, 4)
    rate = units / dt
    S = packed[0].shape[0]
    return max(4, int(rate * target_s / (S * HORIZON)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    from spi_active_b200 import cem
    orc.build()
    cores = orc.num_threads()
    blob, (S, ds) = oracle_dataset()
    packed = pack_numpy(ds)
    cfg = cem.default_full_config()
    # each step = a bounded sample of the N x 4096-candidate population; sized so K + W steps end in minutes
    budget_s = float(os.environ.get("SPI_BENCH_REFERENCE_BUDGET_S", "120"))
    n_cand = calibrate_oracle(blob, packed, cfg, budget_s / max(1, args.steps + args.warmup))
    n_cand = min(n_cand, CANDIDATES_PER_GPU * args.gpus)
    for w in range(args.warmup):
        time_oracle(blob, packed, cfg, n_cand, seed=w)
    total_t, total_u = 0.0, 0
    for k in range(args.steps):
        dt, u = time_oracle(blob, packed, cfg, n_cand, seed=100 + k)
        total_t += dt; total_u += u
    value = total_u / total_t
    sample = f"{n_cand} of {CANDIDATES_PER_GPU * args.gpus} candidates x S={S} x H={HORIZON} per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, S),
        "reference_note": "Isaac Gym is closed source and not installable; this arm is the CPU restatement (oracle/, fp32) "
                          "on all host threads",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(n_gpus, S):
    return {"workload": f"10-parameter Go2 inertial + motor-model CEM iteration (BASELINE config 4 shape at "
                        f"{CANDIDATES_PER_GPU} candidates per GPU), dataset `{DATA_CONFIG}` S={S}, H={HORIZON}, "
                        f"decimation 4",
            "candidates_per_gpu": CANDIDATES_PER_GPU, "candidates_total": CANDIDATES_PER_GPU * n_gpus,
            "segments": S, "horizon": HORIZON, "params": 10, "motor_model": "act2tau_vec3_tanh",
            "parallelism": f"candidates sharded over {n_gpus} GPU(s), dataset replicated",
            "cache": "L2 flushed (256 MiB memset) between timed iterations; the 0.97 MB dataset is L2-resident "
                     "by design within an iteration (compute-bound path)"}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from spi_active_b200 import cem, recorders
    from spi_active_b200.dataset import pack_segments, to_device
    from spi_active_b200.engine import RolloutEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the rollout engine "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    dev = torch.device(f"cuda:{local_rank}")

    eng = RolloutEngine(device=dev)
    S, ds = build_dataset(recorders.engine_rollout_fn(eng), eng.model)
    segs = pack_segments(to_device(ds, dev))
    packed = pack_numpy(ds)
    cfg = cem.default_full_config(eng.model)
    C_total = CANDIDATES_PER_GPU * world
    opt = cem.CemOptimizer(eng, segs, cfg, C_total, rank, world)
    C_local = opt.c1 - opt.c0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_tf, _ = eng.fp32_peak(8192)

    def step():
        flush.zero_()
        opt.iterate()

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    eng.timing_read(reset=True)
    eng.timing_enable(True)
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    eng.timing_enable(False)
    kern_ms, kern_n = eng.timing_read(reset=True)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    t = torch.tensor([ms_total, kern_ms / max(kern_n, 1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kern_ms_per_launch = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    value = C_total * S * HORIZON / (ms_per_step * 1e-3)
    best = opt.best.cpu().numpy()

    # ---- e2e: host buffers through the C-ABI host entry point, copies inside the timed region ----------
    host_params = opt.params[opt.c0:opt.c1].cpu().numpy().copy()
    init, act, tgt, gains, mask, denom = packed
    for _ in range(2):
        eng.evaluate_candidates_host(host_params, cfg.names, init, act, tgt, gains, mask, motor_model=cfg.motor_model,
                                     flags=cfg.flags, cost_denominator=denom)
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        cost_h, status_h = eng.evaluate_candidates_host(host_params, cfg.names, init, act, tgt, gains, mask,
                                                        motor_model=cfg.motor_model, flags=cfg.flags,
                                                        cost_denominator=denom)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - te0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = C_total * S * HORIZON / (float(te[0]) / args.steps)
    h2d = host_params.nbytes + init.nbytes + act.nbytes + tgt.nbytes + gains.nbytes + mask.nbytes
    d2h = cost_h.nbytes + status_h.nbytes

    # ---- the other BASELINE configs, bounded (all ranks take part: config 4 strong scaling and config 5 are sharded) ----
    def max_over_ranks(x):
        tt = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0])
    extra = None
    if not args.no_extras:
        try:
            extra = run_extras(eng, segs, cfg, S, rank, world, barrier, max_over_ranks)
        except Exception as exc:            # the headline line must survive a failure of a secondary record
            extra = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if world > 1:
            _rank0_done_wait(world)
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (rollout_kernel) ------------------------------------------------
    rc = roofline_constants()
    flops_per_rollout = rc["full10_vec3_tanh"]["total"]
    alg_flops = flops_per_rollout * C_local * S
    achieved_tf = alg_flops / (kern_ms_per_launch * 1e-3) / 1e12
    alg_bytes = rc["bytes_per_segment"] * S + C_local * (4 * 10 + 12)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = {
        "bound": "fp32", "kernel": "rollout_ws_kernel", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf, "traffic": ncu_traffic(),
        "peak_source": "FP32 FFMA microbenchmark (spi_b200_fp32_peak) measured in this run on this GPU; "
                       "MEASURED_PEAKS.json holds no FP32 figure",
        "kernel_ms_per_launch": kern_ms_per_launch, "kernel_share_of_step": kern_ms_per_launch / ms_per_step,
        "algorithmic_flops_per_launch": alg_flops, "flops_per_rollout": flops_per_rollout,
        "executed": (lambda fr: None if fr is None else {
            "flops_per_rollout": fr, "tflops": fr * C_local * S / (kern_ms_per_launch * 1e-3) / 1e12,
            "frac": fr * C_local * S / (kern_ms_per_launch * 1e-3) / 1e12 / peak_tf,
            "note": "hardware-style fraction: FP32 operations the kernel executes (ncu thread-level FFMA x 2 + FMUL + FADD, "
                    "profiles/ncu_rollout_summary.json) over the same FFMA peak; `frac` above counts the oracle's operations"
        })(ncu_executed_flops_per_rollout()),
        "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                "achieved_gbs": alg_bytes / (kern_ms_per_launch * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                "frac": alg_bytes / (kern_ms_per_launch * 1e-3) / 1e9 / hbm_peak},
    }

    # ---- cpu_baseline: the CPU oracle (fp32, all host cores) on a bounded sample --------------------------
    from oracle import oracle as orc
    orc.build()
    blob = eng.blob
    n_cand = calibrate_oracle(blob, packed, cfg, 12.0)
    dt, units = time_oracle(blob, packed, cfg, n_cand)
    cpu = {"value": units / dt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
           "sample": f"{n_cand} candidates x S={S} x H={HORIZON} = {units} env steps in {dt:.1f} s (oracle fp32, "
                     f"std::thread over (candidate, segment))"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "s_per_sysid_iteration": ms_per_step * 1e-3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, S),
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "api": "RolloutEngine.evaluate_candidates_host -> spi_b200_eval_candidates_host"},
        "gpu_launches": int(launches), "clocks": clocks,
        "best": {"cost": float(best[-1]), "base_mass_kg": float(best[0])},
    }
    if extra is not None:
        line["extra"] = extra
    emit(line)
    if world > 1:
        _rank0_done_signal()
        dist.destroy_process_group()
    return 0


# Rank 0 times the CPU oracle (cpu_baseline) after the GPU work; the other ranks must not spin in an NCCL barrier meanwhile
# (a spinning rank costs a host core each and depressed the 8-GPU cpu_baseline by 20 % in round 1): they sleep on the
# c10d TCP store instead.
def _rank0_done_signal():
    import torch.distributed as dist
    try:
        dist.distributed_c10d._get_default_store().set("spi_bench_rank0_done", "1")
    except Exception:
        pass


def _rank0_done_wait(world):
    import datetime

    import torch.distributed as dist
    try:
        dist.distributed_c10d._get_default_store().wait(["spi_bench_rank0_done"], datetime.timedelta(seconds=600))
    except Exception:
        pass


# ------------------------------------------------------------------------------------------------
# bounded secondary records of the default run (`extra`): the other BASELINE.json configs, timed the same way
# ------------------------------------------------------------------------------------------------
def run_extras(eng, segs, cfg, S, rank, world, barrier, max_over_ranks):
    """-> dict for the JSON line's `extra` key.  Every record is measured with CUDA events / a synchronised wall clock after
    its own warm-up, max over ranks, and is small enough that the default run still ends within a minute or two:
      config4_strong  BASELINE config 4 as written: 16384 candidates IN TOTAL split over the N ranks (strong scaling),
                      ms per CEM iteration (sample -> rollout -> cost -> all-gather -> select / refit)
      config2         scripts/mass_landscape.py --config all --horizon 5: the 20-candidate base-mass sweep, ms per sweep (rank 0)
      config3         scripts/mass_opt.py: the 50-trial TPE study + the final re-evaluation = 51 sequential single-candidate
                      launches, seconds per study incl. the host-side sampler (rank 0)
      active          BASELINE config 5: one evaluate_policy of 1024 command trajectories per GPU x 11 envs x 1249 control
                      steps (3 pipelines), seconds per rollout and env steps/s, plus the kernel times of one un-graphed
                      control step (actor / physics / post-step) and of the Fisher contraction
    """
    import torch

    from spi_active_b200 import active as act, cem, landscape
    out = {}
    dev = eng.device

    def timed(fn, n, warm=1):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / n

    # ---- config 4, strong scaling -----------------------------------------------------------------------------------
    C4 = 16384
    opt = cem.CemOptimizer(eng, segs, cfg, C4, rank, world)
    ms = timed(opt.iterate, 3)
    out["config4_strong"] = {"candidates_total": C4, "candidates_per_gpu": C4 // world, "n_gpus": world,
                             "ms_per_cem_iteration": ms, "candidate_env_steps_per_s": C4 * S * HORIZON / (ms * 1e-3),
                             "scaling": "strong"}
    del opt
    if rank == 0:
        # ---- config 2: the 20-point landscape ---------------------------------------------------------------------------
        scales = np.linspace(landscape.MASS_SCALE_MIN, landscape.MASS_SCALE_MAX, landscape.MASS_SAMPLES)
        ref_masses = eng.model.body_masses_isaac_order()
        ms2 = timed(lambda: landscape.mass_sweep(eng, segs, ref_masses, scales), 5, warm=2) if world == 1 else None
        if world > 1:       # the other ranks are parked at a barrier: no collective inside, plain local timing
            for _ in range(2):
                landscape.mass_sweep(eng, segs, ref_masses, scales)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(5):
                landscape.mass_sweep(eng, segs, ref_masses, scales)
            torch.cuda.synchronize(); ms2 = (time.perf_counter() - t0) / 5 * 1e3
        out["config2"] = {"candidates": int(landscape.MASS_SAMPLES), "segments": S, "ms_per_sweep": ms2,
                          "candidate_env_steps_per_s": landscape.MASS_SAMPLES * S * HORIZON / (ms2 * 1e-3),
                          "note": "includes the D2H of the [20,3] costs, as the CLI does"}
        # ---- config 3: the TPE study ---------------------------------------------------------------------------------------
        obj = lambda sc: landscape.evaluate_mass_scale(sc, eng, segs, ref_masses)
        landscape.optimize_mass(obj, n_trials=5)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        bs, bv, trials = landscape.optimize_mass(obj)
        landscape.evaluate_mass_scale(bs, eng, segs, ref_masses, return_details=True)      # mass_opt.py:226-228
        torch.cuda.synchronize(); s3 = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(51):
            obj(1.0)
        torch.cuda.synchronize(); s3_eval = time.perf_counter() - t0
        out["config3"] = {"trials": len(trials) + 1, "s_per_study": s3, "s_in_the_51_evaluations": s3_eval,
                          "best_base_mass_kg": bs * float(ref_masses[0]), "best_cost": bv,
                          "note": "51 sequential C = 1 launches (55 CTAs each: latency-bound by design of the study); the "
                                  "rest of the time is the host-side TPE sampler"}
    barrier()
    # ---- config 5: active exploration ------------------------------------------------------------------------------------------
    M, T = 1024, 1250
    acfg = act.ActiveConfig(exploration_params=list(act.ActiveExploration.PARAM_ORDER), seed=rank)
    ex = act.PipelinedExploration(eng, act.PolicyMLP.random(dev, seed=0, gain=0.3), M, acfg, n_pipelines=3)
    rng = np.random.default_rng(rank)
    r = np.asarray(act.COMMAND_RANGES)
    vals = rng.uniform(r[act.COMMAND_SAMPLING_IDXS, 0], r[act.COMMAND_SAMPLING_IDXS, 1], (M, 5, 3)).astype(np.float32)
    cmds = torch.from_numpy(np.stack([act.expand_commands(act.commands_constant(v, 250)) for v in vals])).pin_memory()
    res = {}

    def rollout():
        res["out"] = ex.evaluate_policy(cmds, total_steps=T)
    for _ in range(3):          # graph capture + allocator / clock warm-up (one warm-up rollout left 0.14 - 0.18 s of run-to-run spread)
        rollout()
    barrier(); t0 = time.perf_counter()
    for _ in range(5):
        rollout()
    barrier()
    s5 = max_over_ranks(time.perf_counter() - t0) / 5
    rec = {"main_envs_per_gpu": M, "envs_per_gpu": ex.num_envs, "control_steps": int(res["out"]["steps"]), "n_gpus": world,
           "s_per_rollout": s5, "env_steps_per_s": world * ex.num_envs * res["out"]["steps"] / s5,
           "reward_mean": float(res["out"]["total_reward"][::ex.param_dim + 1].mean())}
    # kernel times of one control step of the un-cut population, launched eagerly (no graph, one stream): CUDA events
    one = act.ActiveExploration(eng, act.PolicyMLP.random(dev, seed=0, gain=0.3), M, acfg)
    one.reset_all(cmds, total_steps=64)
    if one.step_impl == "fused" and one.tc_policy is not None and getattr(one, "ring", False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = np.zeros(3)
        n_rep = 20
        for k in range(n_rep + 3):
            one.step_idx += 1
            ev[0].record()
            raw = one.tc_policy.forward_ring(one.obs_hi, one.obs_lo, one.num_envs, one.ctrl[3:4], out=one.raw_actions)
            ev[1].record()
            one.backend.env_step(one.state, raw, params=one.params, param_names=one.param_names,
                                 motor_model=acfg.motor_model, flags=1, zero_action_mask=one.done)
            ev[2].record()
            one.backend.active_post_step(one.state, raw, one.done, one.main_commands, one.commands, one.actions,
                                         one.gait_indices, one.clock, None, None, None, one.hist, one.live_hist,
                                         one.dead_steps, one.schedule, one.counter, one.ctrl, one.dt, acfg.action_clip,
                                         act.CLIP_OBSERVATIONS, act.TERMINATION_GRAVITY, one.model.q_default,
                                         obs_hi=one.obs_hi, obs_lo=one.obs_lo, ring_slots=act.RING_SLOTS,
                                         fim_jtj=one.jtj if one.fim_mode == "fused" else None,
                                         fim_trace=one.trace_acc if one.fim_mode == "fused" else None,
                                         fim_delta=float(acfg.delta_param))
            ev[3].record()
            torch.cuda.synchronize()
            if k >= 3:
                acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        # the alternative Fisher path (fim_mode="tensor"): tcgen05 contraction of a 64-step history of the same shape
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hist64 = torch.randn(64, M, one.param_dim + 1, 25, device=dev)
        live64 = torch.ones(64, M, dtype=torch.uint8, device=dev)
        for _ in range(3):
            one.backend.fim_contract(hist64, float(acfg.delta_param), live=live64)
        e0.record()
        for _ in range(10):
            one.backend.fim_contract(hist64, float(acfg.delta_param), live=live64)
        e1.record(); torch.cuda.synchronize()
        fim_ms = e0.elapsed_time(e1) / 10
        flops = 2.0 * 3 * one.num_envs * sum(a * b for a, b in zip((928, 512, 256), (512, 256, 128)))
        rec["step_kernels_ms"] = {"actor_mlp": acc[0] / n_rep, "physics_env_step": acc[1] / n_rep,
                                  "post_step": acc[2] / n_rep,
                                  "fisher": "fused into post_step (fim_mode=%s)" % one.fim_mode,
                                  "fim_contract_tcgen05_per_64_steps": fim_ms,
                                  "fim_contract_tcgen05_gbs": hist64.numel() * 4 / (fim_ms * 1e-3) / 1e9,
                                  "actor_issued_f16_tflops": flops / (acc[0] / n_rep * 1e-3) / 1e12,
                                  "note": "eager single-stream launches incl. launch gaps; inside the captured, 3-way pipelined "
                                          "rollout the kernels overlap"}
    out["active"] = rec
    return out


class _StdoutGuard:
    """Everything the process (python or C libraries such as NCCL's version banner) writes to fd 1 goes to stderr;
    the one JSON line of the contract is written to the real stdout with emit()."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


_GUARD = None


def emit(obj):
    line = json.dumps(obj)
    if _GUARD is not None:
        _GUARD.emit(line)
    else:
        print(line, flush=True)


# ------------------------------------------------------------------------------------------------
# optional second workload (NOT the headline metric): the active-exploration rollout of BASELINE config 5
# ------------------------------------------------------------------------------------------------
def run_active(args):
    """`--workload active`: a "step" is one evaluate_policy call — 1024 command trajectories per GPU x (1 + 10) envs x
    1249 closed-loop control steps (actor MLP on the tensor cores, physics, fused post-step, Fisher contraction), the trials
    cut into 3 pipelines (active.PipelinedExploration).  Reports env-steps/s; same barrier / max-over-ranks timing."""
    import torch
    import torch.distributed as dist

    from spi_active_b200 import active as act
    from spi_active_b200.engine import RolloutEngine
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    eng = RolloutEngine(device=torch.device(f"cuda:{local_rank}"))
    M, T = 1024, 1250
    cfg = act.ActiveConfig(exploration_params=list(act.ActiveExploration.PARAM_ORDER), seed=rank)
    ex = act.PipelinedExploration(eng, act.PolicyMLP.random(eng.device, seed=0, gain=0.3), M, cfg, n_pipelines=3)
    rng = np.random.default_rng(rank)
    r = np.asarray(act.COMMAND_RANGES)
    vals = rng.uniform(r[act.COMMAND_SAMPLING_IDXS, 0], r[act.COMMAND_SAMPLING_IDXS, 1], (M, 5, 3)).astype(np.float32)
    cmds = torch.from_numpy(np.stack([act.expand_commands(act.commands_constant(v, 250)) for v in vals])).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 1)):
        out = ex.evaluate_policy(cmds, total_steps=T)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = ex.evaluate_policy(cmds, total_steps=T)
    barrier()
    t = torch.tensor([time.perf_counter() - t0], device=eng.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s_per = float(t[0]) / args.steps
    if rank == 0:
        rew = out["total_reward"][::ex.param_dim + 1]
        emit({"metric": "active_env_steps_per_s", "value": world * ex.num_envs * out["steps"] / s_per, "unit": "env steps/s",
              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s_per,
              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (actor: fp16 pairs, fp32-grade)",
              "data": "synthetic",
              "config": {"workload": "active-exploration rollout (BASELINE config 5): 1024 command trajectories per GPU x "
                                     "(1 + 10) envs x 1249 closed-loop control steps, 900-512-256-128-12 actor, 3 pipelines",
                         "main_envs_per_gpu": M, "envs_per_gpu": ex.num_envs, "control_steps": out["steps"]},
              "reward_mean": float(rew.mean()), "note": "secondary workload; the headline metric is the default run"})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    global _GUARD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the bounded `extra` records (other BASELINE configs)")
    ap.add_argument("--workload", default="sysid", choices=["sysid", "active"],
                    help="sysid (default) = the headline metric; active = the config-5 exploration rollout (secondary)")
    args = ap.parse_args()
    if not (args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl != "reference"):
        _GUARD = _StdoutGuard()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself one rank per GPU
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), str(Path(__file__).resolve()),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        if args.workload != "sysid":
            cmd += ["--workload", args.workload]
        return subprocess.call(cmd)
    if args.workload == "active":
        return run_active(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
