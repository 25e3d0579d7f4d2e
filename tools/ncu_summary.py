#!/usr/bin/env python
"""Development aid: condense an `ncu --set full` report of interp_kernel into the CSV kept under profiles/, and refresh
profiles/step_traffic.json (the DRAM bytes per launch that bench.py reports as roofline.traffic).

    python tools/ncu_summary.py gpurun_out/prof_step.ncu-rep profiles/r1_ncu_step_v8.csv "header comment"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = (
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
    "launch__registers_per_thread", "sm__cycles_elapsed.max", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
)


def main():
    rep, out, comment = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    row = next(r for r in rows[2:] if "interp_kernel" in " ".join(r))
    d = {h: (u, v) for h, u, v in zip(hdr, units, row)}
    lines = [f"# {comment}", f"Kernel Name,,{d['Kernel Name'][1]}"]
    for k in sorted(d):
        if k in KEEP or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            lines.append(f"{k},{d[k][0]},{d[k][1]}")
    open(out, "w").write("\n".join(lines) + "\n")

    def to_bytes(key):
        u, v = d[key]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return float(v.replace(",", "")) * scale
    traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    tj = os.path.join(ROOT, "profiles", "step_traffic.json")
    t = json.load(open(tj)) if os.path.exists(tj) else {}
    t["7b"] = {"dram_bytes_per_launch": traffic, "source": f"{os.path.relpath(out, ROOT)} (ncu --set full, one interp_kernel launch at pos 6)"}
    json.dump(t, open(tj, "w"), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    main()
