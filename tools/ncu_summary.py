#!/usr/bin/env python3
"""Summaries of ncu artefacts brought back in gpurun_out/ (run here, no GPU needed):

    python tools/ncu_summary.py launches gpurun_out/ev_launches_tstar.csv      # per-kernel launch count / time / share
    python tools/ncu_summary.py rep gpurun_out/ev_schur_f200.ncu-rep           # headline metrics of every captured launch
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("duration_us", "gpu__time_duration.sum"),
    ("dram_read_bytes", "dram__bytes_read.sum"),
    ("dram_write_bytes", "dram__bytes_write.sum"),
    ("fp64_tensor_path_pct_of_peak", "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed"),
    ("fp64_pipe_cycles_active_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("issue_active_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("warp_instructions", "smsp__inst_executed.sum"),
    ("registers_per_thread", "launch__registers_per_thread"),
    ("dyn_smem_bytes", "launch__shared_mem_per_block_dynamic"),
    ("grid", "launch__grid_size"),
    ("shared_ld_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("shared_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("l2_hit_rate_pct", "lts__t_sector_hit_rate.pct"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6,
        "nsecond": 1e-3, "ns": 1e-3}


def launches(path):
    lines = [ln for ln in open(path, errors="replace") if not ln.startswith("==")]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# %s: %d launches, %.1f us of kernel time (cold-cache, serialised: shares, not absolutes)" % (path, sum(a[0] for a in agg.values()), tot))
    print("# %8s %12s %7s  kernel" % ("launches", "total_us", "share"))
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %8d %12.1f %6.1f%%  %s" % (n, t, 100 * t / tot, name[:170]))


def rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none --import-source on: %s" % path)
    for r in rows[2:]:
        print("kernel: %s" % r[col["Kernel Name"]][:200])
        for label, key in KEYS:
            if key in col:
                v = r[col[key]]
                try:
                    v = float(v.replace(",", "")) * UNIT.get(units[col[key]], 1.0)
                except ValueError:
                    pass
                print("  %-34s %s" % (label, ("%.6g" % v) if isinstance(v, float) else v))
        st = []
        for h in hdr:
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
            if m:
                try:
                    st.append((float(r[col[h]]), m.group(1)))
                except ValueError:
                    pass
        print("  stall reasons per issue (top 6):    " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:6]))


if __name__ == "__main__":
    {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2])
