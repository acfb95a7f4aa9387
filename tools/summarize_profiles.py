#!/usr/bin/env python3
"""Turn the ncu exports under gpurun_out/ into the small text summaries committed under profiles/.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > gpurun_out/X_raw.csv      (done here if missing)
    python tools/summarize_profiles.py gpurun_out/X.ncu-rep profiles/r1_X.txt
    python tools/summarize_profiles.py --launches gpurun_out/launches.csv profiles/r1_launches.txt
"""
import csv
import os
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def raw(rep):
    out = rep.replace(".ncu-rep", "_raw.csv")
    if not os.path.exists(rep):
        return out          # exported on the GPU box already (the reports themselves are too large to bring back)
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(rep):
        with open(out, "w") as fh:
            subprocess.check_call(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=fh, stderr=subprocess.DEVNULL)
    return out


def kernels(rep, dst):
    rows = list(csv.reader(open(raw(rep))))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none, source {os.path.basename(rep)} (cold-cache, serialised replays)\n")
        for r in rows[2:]:
            fh.write(f"\n== {r[hdr.index('Kernel Name')]}\n")
            for k in KEYS:
                if k in hdr:
                    fh.write(f"{k:82s} {r[hdr.index(k)]:>22s} {units[hdr.index(k)]}\n")
            rd, wr = float(r[hdr.index('dram__bytes_read.sum')]), float(r[hdr.index('dram__bytes_write.sum')])
            u = units[hdr.index('dram__bytes_read.sum')]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
            t = float(r[hdr.index('gpu__time_duration.sum')])
            tu = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}[units[hdr.index('gpu__time_duration.sum')]]
            fh.write(f"{'derived: DRAM traffic (read+write) / duration':82s} {(rd + wr) * scale / (t * tu) / 1e9:22.1f} GB/s\n")
            def sectors(k):
                try:
                    return float(r[hdr.index(k)]) * 32.0 if k in hdr else None
                except ValueError:
                    return None
            tex, fab = sectors("lts__t_sectors_srcunit_tex.sum"), sectors("lts__t_sectors_srcunit_ltcfabric.sum")
            if tex is not None:
                # (lts__t_bytes is not in the --set full export of this ncu; sectors are 32 bytes.  srcunit_tex = what the SMs
                # moved through L2; ltcfabric = the same requests crossing between the two L2 partitions, on top)
                fh.write(f"{'derived: SM<->L2 traffic (lts__t_sectors_srcunit_tex x 32 B)':82s} {tex / 1e9:22.3f} GB\n")
                fh.write(f"{'derived: SM<->L2 traffic / duration':82s} {tex / (t * tu) / 1e9:22.1f} GB/s\n")
            if fab is not None:
                fh.write(f"{'derived: L2 partition fabric traffic (lts__t_sectors_srcunit_ltcfabric x 32 B) / duration':82s} {fab / (t * tu) / 1e9:22.1f} GB/s\n")
            if "lts__t_bytes.sum" in hdr:
                lb = float(r[hdr.index("lts__t_bytes.sum")]) * {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("lts__t_bytes.sum")]]
                fh.write(f"{'derived: L2 traffic (lts__t_bytes) / duration':82s} {lb / (t * tu) / 1e9:22.1f} GB/s\n")


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r and r[0].isdigit()]
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        name, val = r[4], float(r[-1])
        name = name.split("(")[0]
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1; d[1] += val; total += val
    with open(dst, "w") as fh:
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, source {os.path.basename(src)}\n")
        fh.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        fh.write(f"{'kernel':70s} {'launches':>9s} {'total [ms]':>12s} {'share':>8s}\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{name[:70]:70s} {n:9d} {t / 1e6:12.3f} {t / total:8.1%}\n")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernels(sys.argv[1], sys.argv[2])
