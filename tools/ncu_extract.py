"""Turn an `ncu --set full` report into the small JSON that bench.py's roofline block reads (profiles/<capture>.json).

  python tools/ncu_extract.py gpurun_out/prof_r02a.ncu-rep --name r02a --cycles 4096 --stats gpurun_out/bench_r02a.json \
         -o profiles/r02a_ncu.json

Run in the authoring container (ncu reads reports without a GPU).  Per kernel (launches of the same kernel are averaged):
device time, DRAM bytes read / written, L1/shared data-pipe wavefronts and their share of peak, issue / FMA / ALU pipe
utilisation, fp32 flops executed (FFMA counted as 2), registers, occupancy, barrier stalls.  `--stats` is the bench line
of the same command (work counters per step), so that per-unit figures can be derived; the git commit of the tree the
capture was taken from is recorded as `commit` and printed by bench.py as `ncu_capture_commit`.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "ns": 1e-6,
         "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}

WANT = {
    "time_ms": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "l1_wavefronts_per_sm": "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg",
    "l1_shared_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1_data_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "alu_pipe_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "tensor_pipe_pct": "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "barrier_stall_per_issue": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "short_scoreboard_per_issue": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "long_scoreboard_per_issue": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "sm_cycles": "sm__cycles_elapsed.avg",
    "ffma_per_cycle": "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "fadd_per_cycle": "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
    "fmul_per_cycle": "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
    "thread_inst": "thread_inst_executed",
    "warp_inst": "smsp__inst_executed.sum",
    "grid": "Grid Size",
    "block": "Block Size",
}


def short_name(k):
    m = re.search(r"(k_[A-Za-z0-9_]+)", k)
    return m.group(1) if m else k.split("(")[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--name", required=True)
    ap.add_argument("--cycles", type=int, required=True, help="cycles per full-batch launch in the captured command")
    ap.add_argument("--stats", help="bench JSON line of the same command (work_per_step)")
    ap.add_argument("--steps-total", type=int, default=1, help="decode steps the captured command ran (steps + warm-up): launches per step = captured launches / this")
    ap.add_argument("--min-grid", type=int, default=0, help="ignore launches with fewer CTAs (e.g. chunked front-end launches)")
    ap.add_argument("--note", default="")
    ap.add_argument("-o", "--out", required=True)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = csv.reader(io.StringIO(raw))
    head, units = next(rd), next(rd)
    col = {h: i for i, h in enumerate(head)}
    kernels = {}
    for row in rd:
        name = short_name(row[col["Kernel Name"]])
        d = {}
        for key, metric in WANT.items():
            if metric not in col:
                continue
            v, u = row[col[metric]], units[col[metric]]
            if key in ("grid", "block"):
                d[key] = v
                continue
            try:
                x = float(v)
            except ValueError:
                continue
            base = u.split("/")[0]
            d[key] = x * SCALE.get(base, 1.0)
        kernels.setdefault(name, []).append(d)
    out = {}
    for name, rows in kernels.items():
        agg = {"launches_captured": len(rows), "launches_per_step": len(rows) / a.steps_total, "grid": rows[0].get("grid"), "block": rows[0].get("block")}
        for key in WANT:
            if key in ("grid", "block"):
                continue
            vals = [r[key] for r in rows if key in r]
            if vals:
                agg[key] = sum(vals) / len(vals)
        if "ffma_per_cycle" in agg and "sm_cycles" in agg:
            agg["fp32_flop"] = (2 * agg.get("ffma_per_cycle", 0) + agg.get("fadd_per_cycle", 0) + agg.get("fmul_per_cycle", 0)) * agg["sm_cycles"]
        agg["dram_bytes"] = agg.get("dram_read_bytes", 0) + agg.get("dram_write_bytes", 0)
        out[name] = agg
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    dirty = bool(subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "--", "pyft8_b200/csrc", "include"], capture_output=True, text=True).stdout.strip())
    doc = {"capture": a.name, "report": os.path.basename(a.report), "commit": commit + ("+dirty" if dirty else ""),
           "cycles_per_launch": a.cycles, "note": a.note, "kernels": out}
    if a.stats and os.path.exists(a.stats):
        for line in open(a.stats):
            line = line.strip()
            if line.startswith("{"):
                j = json.loads(line)
                doc["work_per_step"] = j.get("work_per_step")
                doc["bench_value"] = j.get("value")
    json.dump(doc, open(a.out, "w"), indent=1)
    for k, v in sorted(out.items(), key=lambda kv: -kv[1].get("time_ms", 0)):
        print(f"{k:24s} {v.get('time_ms', 0):9.3f} ms  dram {v['dram_bytes'] / 1e9:7.3f} GB  l1 {v.get('l1_data_pipe_pct', 0):5.1f}%  "
              f"issue {v.get('issue_active_pct', 0):5.1f}%  fma {v.get('fma_pipe_pct', 0):5.1f}%  alu {v.get('alu_pipe_pct', 0):5.1f}%  regs {v.get('regs', 0):.0f}")


if __name__ == "__main__":
    main()
