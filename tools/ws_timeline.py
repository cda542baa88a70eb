"""dev helper (not a test): where do the cycles of a rollout_ws_kernel sub-step go?

Builds a -DSPI_WS_PROFILE copy of the library (clock64 stamps around every barrier of a window of CTAs in the middle of the
grid), runs the bench shape once and prints, per sub-step round: the legs' phase 1 / publish / wait [A] / wait [B1] / phase 2 /
wait [B2], the base role's wait [A] / solve / advance / bias, the critical path (last leg at [A] -> legs released from [B1]), and
how many of the CTAs resident on an SM are in phase 1 at the same time (the lock-step question of DESIGN.md 4.2).

    python tools/ws_timeline.py build        # here (no GPU): tools/_build/libspi_b200_prof.so
    python tools/ws_timeline.py run [C]      # on the GPU box
"""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PROF_LIB = ROOT / "tools" / "_build" / "libspi_b200_prof.so"
SLOTS, WARPS = 256, 5


def build(extra=()):
    sys.path.insert(0, str(ROOT))
    from spi_active_b200 import _lib
    PROF_LIB.parent.mkdir(exist_ok=True)
    cmd = [_lib._nvcc(), *_lib.NVCC_FLAGS, "-DSPI_WS_PROFILE", *extra, "-o", str(PROF_LIB), *map(str, _lib.SOURCES)]
    subprocess.run(cmd, check=True)
    print("built", PROF_LIB)


def run(Ccand=1024):
    os.environ["SPI_B200_LIB"] = str(PROF_LIB)
    sys.path.insert(0, str(ROOT))
    import numpy as np
    import torch
    import bench
    from spi_active_b200 import cem, recorders
    from spi_active_b200.dataset import pack_segments, to_device
    from spi_active_b200.engine import RolloutEngine

    eng = RolloutEngine()
    S, ds = bench.build_dataset(recorders.engine_rollout_fn(eng), eng.model)
    segs = pack_segments(to_device(ds, eng.device))
    cfg = cem.default_full_config(eng.model)
    f = lambda a: torch.tensor(np.asarray(a, np.float32), device=eng.device)
    params = eng.cem_sample(f(cfg.mean), f(cfg.std), f(cfg.lo), f(cfg.hi), Ccand, 0, cfg.seed, 0)
    n_cta = Ccand * ((S + 31) // 32)
    nblk = 148 * 4 * 3
    blk0 = n_cta // 2
    buf = torch.zeros((nblk, WARPS, SLOTS), dtype=torch.int64, device=eng.device)
    for _ in range(2):
        eng.evaluate_candidates(params, cfg.names, segs, motor_model=cfg.motor_model)
    eng.lib.spi_b200_debug_ws_prof.restype = C.c_int
    eng.lib.spi_b200_debug_ws_prof.argtypes = [C.c_void_p, C.c_int, C.c_int]
    eng.lib.spi_b200_debug_ws_prof(buf.data_ptr(), blk0, nblk)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.evaluate_candidates(params, cfg.names, segs, motor_model=cfg.motor_model)
    e1.record(); torch.cuda.synchronize()
    print(f"C={Ccand} n_cta={n_cta} ms={e0.elapsed_time(e1):.3f} (instrumented)")
    eng.lib.spi_b200_debug_ws_prof(None, 0, 0)
    t = buf.cpu().numpy()
    out = ROOT / "gpurun_out"; out.mkdir(exist_ok=True)
    np.save(out / "ws_timeline.npy", t)
    analyse(t)


def analyse(t):
    import numpy as np
    nblk = t.shape[0]
    smid = t[:, 0, 0]
    # stamps (rollout_ws.cuh, every one predicated on the data of the event):
    #   leg : 0 state arrived ([B2] released)  1 phase 1 computed  2 published (odd legs: partner arrived + added)
    #         3 a0 arrived ([B1] released)     4 phase 2 done
    #   base: 0 torques + bias done  1 pair sums arrived ([A] released)  2 solved  3 [B1] released  4 advanced
    leg = t[:, :4, 1:1 + 200].reshape(nblk, 4, 40, 5).astype(np.float64)
    base = t[:, 4, 1:1 + 200].reshape(nblk, 40, 5).astype(np.float64)
    ok = (leg[:, :, :, 0] > 0).all(axis=(1, 2)) & (base[:, :, 0] > 0).all(axis=1)
    leg, base, smid = leg[ok], base[ok], smid[ok]
    print(f"{ok.sum()} of {nblk} CTAs complete")
    mid = slice(4, 36)
    def m(x): return f"{np.mean(x):7.0f} (p10 {np.percentile(x, 10):6.0f}, p50 {np.percentile(x, 50):6.0f}, p90 {np.percentile(x, 90):6.0f})"
    rnd = leg[:, :, 1:, 0] - leg[:, :, :-1, 0]
    print("round (state arrived -> state arrived)           ", m(rnd[:, :, mid]))
    print("leg: phase 1 compute                             ", m((leg[..., 1] - leg[..., 0])[:, :, mid]))
    print("leg even: publish                                ", m((leg[:, 0::2, :, 2] - leg[:, 0::2, :, 1])[:, :, mid]))
    print("leg odd: wait for the partner + add + publish    ", m((leg[:, 1::2, :, 2] - leg[:, 1::2, :, 1])[:, :, mid]))
    print("leg: published -> a0 arrived ([A] + solve + [B1])", m((leg[..., 3] - leg[..., 2])[:, :, mid]))
    print("leg: phase 2 compute                             ", m((leg[..., 4] - leg[..., 3])[:, :, mid]))
    print("leg: phase 2 done -> next state arrived ([B2])   ", m((leg[:, :, 1:, 0] - leg[:, :, :-1, 4])[:, :, mid]))
    last_pub = leg[..., 2].max(axis=1)
    first_pub = leg[..., 2].min(axis=1)
    print("slowest - fastest leg published                  ", m((last_pub - first_pub)[:, mid]))
    print("base: bias done -> pair sums arrived (wait [A])  ", m((base[..., 1] - base[..., 0])[:, mid]))
    print("CRITICAL last leg published -> base has the sums ", m((base[..., 1] - last_pub)[:, mid]))
    print("CRITICAL base: solve                             ", m((base[..., 2] - base[..., 1])[:, mid]))
    print("CRITICAL solved -> first leg has a0              ", m((leg[..., 3].min(axis=1) - base[..., 2])[:, mid]))
    print("         solved -> last leg has a0               ", m((leg[..., 3].max(axis=1) - base[..., 2])[:, mid]))
    print("base: solved -> [B1] released                    ", m((base[..., 3] - base[..., 2])[:, mid]))
    print("base: advance                                    ", m((base[..., 4] - base[..., 3])[:, mid]))
    print("base: advanced -> torques + bias of the next done", m((base[:, 1:, 0] - base[:, :-1, 4])[:, mid]))
    print("last leg phase 2 done -> first leg next state    ", m((leg[:, :, 1:, 0].min(axis=1) - leg[:, :, :-1, 4].max(axis=1))[:, mid]))
    print("base advanced - last leg phase 2 done (>0: legs wait for the base)", m((base[:, :, 4] - leg[..., 4].max(axis=1))[:, mid]))
    # lock-step: for every SM, sample times; count CTAs whose leg 0 is inside [stamp 0, stamp 2) = phase 1
    hist = np.zeros(8)
    resident = np.zeros(8)
    for sm in np.unique(smid):
        idx = np.where(smid == sm)[0]
        if len(idx) < 4:
            continue
        L = leg[idx]                                   # [n, 4, 40, 6]
        t0, t1 = L[:, 0, 0, 0].max(), L[:, 0, -1, 4].min()
        # the window in which at least 4 CTAs of this SM were recorded concurrently is what we can judge
        ts = np.linspace(L[:, 0, 0, 0].min(), L[:, 0, -1, 4].max(), 4000)
        for tt in ts:
            alive = (L[:, 0, 0, 0] <= tt) & (L[:, 0, -1, 4] > tt)
            na = int(alive.sum())
            if na != 4:
                continue
            inph1 = 0
            for c in np.where(alive)[0]:
                k = np.searchsorted(L[c, 0, :, 0], tt, side="right") - 1
                if k >= 0 and tt < L[c, 0, k, 2]:
                    inph1 += 1
            hist[inph1] += 1
    if hist.sum():
        print("time share with k of the 4 resident CTAs in phase 1 (leg 0), k = 0..4:", np.round(hist[:5] / hist.sum(), 3))
        p = (hist[:5] / hist.sum() * np.arange(5)).sum() / 4
        from math import comb
        print("   independent CTAs with the same duty cycle would give:", np.round([comb(4, k) * p ** k * (1 - p) ** (4 - k) for k in range(5)], 3))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "run":
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 1024)
    else:
        import numpy as np
        analyse(np.load(sys.argv[2]))
