#!/usr/bin/env python
"""Summarise an `ncu --set full` capture (.ncu-rep) as JSON: per kernel the duration, issue
and pipe utilisation, shared-memory wavefronts, DRAM bytes, registers, occupancy limits and
the stall reasons per issue.

    python tools/ncu_summary.py gpurun_out/r02_v10_c3.ncu-rep "note" > profiles/..._summary.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic",
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
REASONS = ["wait", "short_scoreboard", "long_scoreboard", "barrier", "not_selected", "math_pipe_throttle",
           "mio_throttle", "no_instruction", "branch_resolving", "dispatch_stall", "lg_throttle"]


def main():
    rep = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {"source": "ncu --set full --import-source on --clock-control none; " + note, "kernels": {}}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(K3dParams")[0].replace("void <unnamed>::", "")
        k = {}
        for key in KEYS + [STALLS % s for s in REASONS]:
            if key in d and d[key] != "":
                k[key] = {"value": d[key], "unit": units[hdr.index(key)]}
        out["kernels"][name] = k
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
