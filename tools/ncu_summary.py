#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here with `ncu -i`): headline metrics, pipe utilisation, warp-stall
sample breakdown and (with --source) the hottest SASS/source lines.  Used to write profiles/*.txt."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.per_cycle_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "sass__thread_inst_executed_true_per_opcode", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    out = []
    for r in rows[2:]:
        out.append(dict(zip(hdr, r)))
    return hdr, rows[1], out


def main():
    rep = sys.argv[1]
    hdr, units, kernels = raw(rep)
    for k in kernels:
        print("kernel:", k.get("Kernel Name"), "grid", k.get("Grid Size"), "block", k.get("Block Size"))
        for key in KEYS:
            if key in k:
                print("  %-75s %s %s" % (key, k[key], units[hdr.index(key)]))
        st = {h: float(v.replace(",", "")) for h, v in k.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and v not in ("", "n/a")}
        tot = sum(st.values()) or 1
        print("  warp stall samples (%d):" % tot)
        for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
            print("    %-28s %5.1f %%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot))
    if "--source" in sys.argv:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        h = rows[0]
        print(h)
        si = [i for i, x in enumerate(h) if "Sampling" in x and "All" in x]
        if si:
            si = si[0]
            body = [r for r in rows[1:] if len(r) > si and r[si].replace(",", "").isdigit()]
            body.sort(key=lambda r: -int(r[si].replace(",", "")))
            for r in body[:40]:
                print(r[si], "|", " | ".join(r[:3]))


if __name__ == "__main__":
    main()
