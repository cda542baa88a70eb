"""Freeze the algorithmic FLOP figure of the roofline into profiles/roofline.json.

Runs the CPU oracle once per sampled segment on its op-counting scalar type: every add, mul, div,
sqrt, transcendental and compare of the oracle's scalar code for ONE (candidate, segment) rollout
counts 1 (an FMA therefore counts 2).  The figure is independent of how the CUDA kernel is
parallelised and excludes any recomputation the kernel adds.  Test/measurement infrastructure: this is
the only thing the script uses oracle/ for.

    python tests/tools/count_flops.py            # writes profiles/roofline.json
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import oracle as orc  # noqa: E402
from spi_active_b200 import cem, go2_model as gm  # noqa: E402
import synth  # noqa: E402


def main():
    blob = gm.build_model_blob()
    S, ds = synth.dataset("all", 5)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    cfg = cem.default_full_config()
    ids = [gm.PARAM_IDS[n] for n in cfg.names]
    params = np.asarray(cfg.mean, np.float32)
    out = {}
    for label, motor, names, p in (("full10_vec3_tanh", 3, ids, params), ("mass_only_none", 0, [0], params[:1])):
        tot = []
        keys = None
        for s in range(0, S, 10):
            c = orc.count_flops(blob, p, names, init[s], act[s], tgt[s], decimation=4, motor_model=motor)
            keys = list(c.keys())
            tot.append([c[k] for k in keys])
        mean = np.mean(np.asarray(tot, dtype=np.float64), axis=0)
        rec = {k: float(v) for k, v in zip(keys, mean)}
        nsub = int(blob[gm.BLOB["NSUB"]])
        rec["substeps_per_rollout"] = 5 * 4 * nsub
        rec["flops_per_substep"] = rec["total"] / rec["substeps_per_rollout"]
        rec["flops_per_env_step"] = rec["total"] / 5
        out[label] = rec
    out["definition"] = ("mean over every 10th segment of the `all` dataset (S=1730, H=5, decimation 4, "
                         "nsub sub-steps) of the oracle's op count for one (candidate, segment) rollout; "
                         "add=mul=div=sqrt=transcendental=compare=1")
    out["bytes_per_segment"] = 37 * 4 + 5 * 12 * 4 + 19 * 4 + 24 * 4 + 1
    out["bytes_per_candidate_in"] = "4*P"
    out["bytes_per_candidate_out"] = 12
    path = ROOT / "profiles" / "roofline.json"
    path.write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
