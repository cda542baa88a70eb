"""Summarise an `ncu --set full` report (+ optional launch-list csv) into profiles/: a JSON of the metrics the
roofline uses and a short markdown table of stall reasons.  Run in the build container (ncu is installed, no GPU
needed):

    python tools/ncu_summary.py gpurun_out/prof_rollout.ncu-rep --tag r1_b --launches gpurun_out/launches.csv
"""
import argparse
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEYS = {
    "gpu__time_duration.sum": "duration",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__grid_size": "grid_size", "launch__block_size": "block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions_executed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_utilisation_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_inst_pct_of_peak",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_cycles_active_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_inst_pct_of_peak",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_inst_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
    "smsp__warps_active.avg.per_cycle_active": "active_warps_per_scheduler",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "smsp__sass_inst_executed_op_local_ld.sum": "local_loads", "smsp__sass_inst_executed_op_local_st.sum": "local_stores",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum": "thread_ffma", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum": "thread_fmul",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum": "thread_fadd",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3,
         "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--launches")
    ap.add_argument("--row", type=int, default=0, help="which captured launch of the report")
    ap.add_argument("--set-roofline-traffic", action="store_true",
                    help="also write profiles/ncu_rollout_summary.json (bench.py reads `traffic` from it)")
    ap.add_argument("--rollouts", type=float, default=None,
                    help="(candidate, segment) rollouts of the captured launch (C x S), written next to the traffic figure")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + a.row]
    out = {"report": Path(a.report).name, "kernel": vals[hdr.index("Kernel Name")]}
    stalls = {}
    for i, h in enumerate(hdr):
        v = vals[i].replace(",", "")
        if h in KEYS:
            try:
                x = float(v)
            except ValueError:
                continue
            u = units[i]
            if KEYS[h].startswith("dram_bytes") or KEYS[h] == "duration":
                x *= SCALE.get(u, 1.0)
                u = "byte" if "bytes" in KEYS[h] else "s"
            out[KEYS[h]] = x
            out.setdefault("units", {})[KEYS[h]] = u
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v)
    out["dram_bytes_per_launch"] = out.get("dram_bytes_read", 0.0) + out.get("dram_bytes_write", 0.0)
    # FP32 operations the kernel actually EXECUTED (thread level, predicated-on): 2 x FFMA + FMUL + FADD per elapsed cycle x cycles
    per_cycle = {}
    for op in ("ffma", "fmul", "fadd"):
        key = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"
        if key in hdr:
            per_cycle[op] = float(vals[hdr.index(key)].replace(",", ""))
    if len(per_cycle) == 3 and "duration" in out and "sm_clock" in out:
        cycles = out["duration"] * out["sm_clock"] * 1e9
        out["executed_fp32"] = {"thread_ops_per_cycle": per_cycle,
                                "flop_per_launch": (2 * per_cycle["ffma"] + per_cycle["fmul"] + per_cycle["fadd"]) * cycles,
                                "note": "2 x FFMA + FMUL + FADD (MUFU / FMNMX not counted)"}
    out["stall_cycles_per_issued_instruction"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    if a.launches:
        lr = [r for r in csv.reader(open(a.launches)) if len(r) > 10]
        h = lr[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
        agg = {}
        for r in lr[1:]:
            name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            agg.setdefault(name, [0, 0.0])
            agg[name][0] += 1; agg[name][1] += float(r[vi]) * 1e-9
        tot = sum(v[1] for v in agg.values())
        out["launch_list"] = {k: {"launches": v[0], "seconds": v[1], "share": v[1] / tot} for k, v in
                              sorted(agg.items(), key=lambda kv: -kv[1][1])}
    dst = ROOT / "profiles" / f"ncu_{a.tag}.json"
    dst.write_text(json.dumps(out, indent=1) + "\n")
    if a.set_roofline_traffic:
        (ROOT / "profiles" / "ncu_rollout_summary.json").write_text(json.dumps(
            {"source": dst.name, "kernel": out["kernel"], "grid_size": out.get("grid_size"),
             "dram_bytes_per_launch": out["dram_bytes_per_launch"], "rollouts_per_launch": a.rollouts,
             "executed_fp32_flop_per_launch": out.get("executed_fp32", {}).get("flop_per_launch"),
             "note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the rollout kernel at the grid "
                     "size above (ncu --set full); the dataset and candidate rows are L2-resident, so DRAM traffic hardly grows "
                     "with the candidate count"}, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
