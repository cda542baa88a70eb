#!/usr/bin/env python
"""Active SysID command search on the B200 engine — the counterpart of the reference's spigym/run_active_sysid.py
(:48-159) + agents/sysid/active_sysid.py:165-242: load a policy checkpoint, build (1 main + P aux) env groups, run
`iterations` rounds of M command trajectories through evaluate_policy (FIM reward), save best_commands.npz.

    python scripts/run_active_sysid.py --checkpoint logs/.../model_1000.pt --num-envs 1024 \
        --exploration-params mass comx comy comz inertiax inertiay inertiaz motor_model_hip_a motor_model_thigh_a motor_model_calf_a

Without --checkpoint a random-init actor of the same architecture is used (there is no trained policy in the
reference checkout; the numbers are then throughput / plumbing numbers only).

Multi-GPU: launch with `python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ...`;
--num-envs is then the number of main envs PER GPU, the N * num_envs trials of an iteration are sharded over the ranks
and the per-trial rewards / Fisher blocks are all-gathered with NCCL (active.optimize_commands)."""
from __future__ import annotations

import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from spi_active_b200 import active as act  # noqa: E402
from spi_active_b200.engine import RolloutEngine  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser(description="Active SysID excitation-trajectory search")
    ap.add_argument("--checkpoint", type=Path, default=None, help="PPO checkpoint with `actor_model_state_dict`")
    ap.add_argument("--num-envs", type=int, default=1024, help="number of MAIN envs (+num_envs of the reference)")
    ap.add_argument("--exploration-params", nargs="+", default=["mass"])
    ap.add_argument("--delta-param", type=float, default=0.1)
    ap.add_argument("--ksync-steps", type=int, default=5)
    ap.add_argument("--motor-model", default="act2tau_vec3_tanh")
    ap.add_argument("--iterations", type=int, default=5)
    ap.add_argument("--rollout-length", type=float, default=25.0)
    ap.add_argument("--horizon-length", type=float, default=5.0)
    ap.add_argument("--command-sampling-mode", default="constant", choices=["constant", "polynomial", "bezier"],
                    help="config/algo/active_sysid.yaml:44-50")
    ap.add_argument("--poly-degree", type=int, default=3)
    ap.add_argument("--num-bezier-points", type=int, default=4)
    ap.add_argument("--log-dir", type=Path, default=Path("logs/active_sysid"))
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--fim-mode", default="auto", choices=["auto", "step", "tensor"],
                    help="per-step CUDA-core J J^T or deferred tensor-core contraction (spi_b200_fim_contract)")
    ap.add_argument("--pipelines", type=int, default=3,
                    help="independent sub-populations whose control steps are interleaved on their own CUDA streams "
                         "(active.PipelinedExploration; 1 = one explorer)")
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    eng = RolloutEngine()
    policy = (act.PolicyMLP.from_checkpoint(args.checkpoint, eng.device) if args.checkpoint
              else act.PolicyMLP.random(eng.device, seed=args.seed))
    cfg = act.ActiveConfig(exploration_params=args.exploration_params, delta_param=args.delta_param,
                           ksync_steps=args.ksync_steps, motor_model=args.motor_model,
                           rollout_length=args.rollout_length, seed=args.seed, fim_mode=args.fim_mode)
    ex = (act.PipelinedExploration(eng, policy, args.num_envs, cfg, n_pipelines=args.pipelines)
          if args.pipelines > 1 and args.num_envs >= 64 * args.pipelines else act.ActiveExploration(eng, policy, args.num_envs, cfg))
    if rank == 0:
        print(f"Active SysID: {world} GPU(s) x {args.num_envs} main envs x (1 + {ex.param_dim}) = {world * ex.num_envs} envs, "
              f"{ex.total_steps} steps/rollout, FIM mode {ex.fim_mode}")
    t0 = time.perf_counter()
    res = act.optimize_commands(ex, args.iterations, args.rollout_length, args.horizon_length, args.seed,
                                rank=rank, world=world, mode=args.command_sampling_mode, poly_degree=args.poly_degree,
                                num_bezier_points=args.num_bezier_points)
    dt = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    args.log_dir.mkdir(parents=True, exist_ok=True)
    np.savez(args.log_dir / "best_commands.npz", best_commands=res["best_commands"])     # active_sysid.py:223-227
    print(f"Best value: {res['best_value']:.6g}; per-iteration best: {[f'{v:.4g}' for v in res['history']]}")
    print(f"{args.iterations} iterations in {dt:.2f} s ({dt / args.iterations:.2f} s per iteration); "
          f"saved {args.log_dir / 'best_commands.npz'}")


if __name__ == "__main__":
    main()
