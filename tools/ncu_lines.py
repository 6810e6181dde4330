#!/usr/bin/env python
"""Per-source-line view of an ncu --set full --import-source on capture.

    ncu -i REP --page source --csv --print-source cuda,sass --kernel-name regex:solve > x.csv
    python tools/ncu_lines.py x.csv [nodes] [top]

Prints, for the source lines of k3d_kernels.cu with the most warp instructions: instructions
executed (per node when `nodes` is given), stall samples and the dominant stall reasons.
"""
import csv
import sys


def main():
    path = sys.argv[1]
    nodes = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    rows = list(csv.reader(open(path)))
    hdr = None
    lines = []
    ix_s = 0
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            ix_s = r.index("# Samples")
            continue
        if hdr is None or not r or r[0] in ("", "File Path", "Function Name"):
            continue
        if not r[0].isdigit() or r[ix_s] == "-":
            continue
        lines.append(r)
    ix = {n: i for i, n in enumerate(hdr)}
    # the header has "Source" twice; stall columns appear once before the "(Not Issued)" ones
    stall_cols = [(n, i) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    tot_inst = sum(float(r[ix["Instructions Executed"]]) for r in lines)
    tot_samp = sum(float(r[ix["# Samples"]]) for r in lines)
    print("total warp instructions %.4g (%.1f per node), samples %d" % (tot_inst, tot_inst / nodes, tot_samp))
    agg = {}
    for _, i in stall_cols:
        pass
    stall_tot = {n: sum(float(r[i]) for r in lines) for n, i in stall_cols}
    print("stall samples:", ", ".join("%s %.1f%%" % (n[6:], 100 * v / max(tot_samp, 1))
                                      for n, v in sorted(stall_tot.items(), key=lambda kv: -kv[1])[:8]))
    lines.sort(key=lambda r: -float(r[ix["# Samples"]]))
    print("%5s %9s %7s  %-40s %s" % ("line", "inst/node", "samp%", "top stalls", "source"))
    for r in lines[:top]:
        inst = float(r[ix["Instructions Executed"]])
        samp = float(r[ix["# Samples"]])
        st = sorted(((float(r[i]), n[6:]) for n, i in stall_cols), reverse=True)[:3]
        print("%5s %9.1f %6.2f%%  %-40s %s" % (r[0], inst / nodes, 100 * samp / max(tot_samp, 1),
                                             " ".join("%s:%d" % (n, v) for v, n in st), r[1].strip()[:90]))


if __name__ == "__main__":
    main()
