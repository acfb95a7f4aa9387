#!/usr/bin/env python3
"""Turn this round's gpurun_out/ artefacts into the files committed under profiles/ (r2_*)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import summarize_profiles as sp

SCALE = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
TSCALE = {"s": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9}


def copy(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
        print("copied", dst)


def main():
    global LANES
    try:
        # (the ncu captures run bench.py with its defaults: the lane count of the newest default-run record)
        newest = max((f for f in ("r2_bench_n1_3dB.json", "r2_bench_quick.json") if os.path.exists(os.path.join(G, f))),
                     key=lambda f: os.path.getmtime(os.path.join(G, f)))
        LANES = json.loads(open(os.path.join(G, newest)).read().strip().splitlines()[-1])["config"]["decoder_lanes"]
    except Exception:
        LANES = 4096
    copy("r2_bench_n1_3dB.json", "r2_bench_n1_3dB.json")
    copy("r2_bench_reference_arm.json", "r2_bench_reference_arm.json")
    copy("r2_membench.txt", "r2_membench.txt")
    for n in (2, 4, 8):
        copy(f"r2_bench_n{n}.json", f"r2_bench_n{n}.json")
    copy("r2_multi_gpu_record.txt", "r2_multi_gpu_record.txt")
    if os.path.exists(os.path.join(G, "r2_launches.csv")):
        sp.launches(os.path.join(G, "r2_launches.csv"), os.path.join(P, "r2_bench_launches.txt"))
        print("wrote r2_bench_launches.txt")
    entries = []
    for rep, dst in (("r2_bench_chain_full.ncu-rep", "r2_bench_chain_ncu_full.txt"),
                     ("r2_irregular_full.ncu-rep", "r2_irregular_decoder_ncu_full.txt")):
        path = os.path.join(G, rep)
        if not os.path.exists(path) and not os.path.exists(path.replace(".ncu-rep", "_raw.csv")):
            continue
        sp.kernels(path, os.path.join(P, dst))
        print("wrote", dst)
        rows = list(csv.reader(open(sp.raw(path))))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            if "k_fused" not in name:
                continue
            g = lambda k: float(r[hdr.index(k)]) * SCALE[units[hdr.index(k)]]
            rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
            sec = lambda k: float(r[hdr.index(k)]) * 32.0          # L2 sectors are 32 bytes
            # what the SMs moved through L2 (srcunit_tex) and what crossed between the two L2 partitions on top of it
            l2, l2_fabric = sec("lts__t_sectors_srcunit_tex.sum"), sec("lts__t_sectors_srcunit_ltcfabric.sum")
            l2_rd, l2_wr = sec("lts__t_sectors_srcunit_tex_op_read.sum"), sec("lts__t_sectors_srcunit_tex_op_write.sum")
            t = float(r[hdr.index("gpu__time_duration.sum")]) * TSCALE[units[hdr.index("gpu__time_duration.sum")]]
            entries.append({
                "source": f"profiles/{dst} (ncu --set full --clock-control none -k regex:k_fused... python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-extras)",
                "kernel": name.split("(")[0].replace("void qr::", ""),
                "config": {"n": 64800, "frames": 4096, "max_iterations": 50, "precision": "fp32", "schedule": "fused", "lanes": LANES},
                "frame_iterations": 4096 * 50,
                "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                "l2_bytes_per_launch": l2, "l2_read_bytes": l2_rd, "l2_write_bytes": l2_wr,
                "l2_partition_fabric_bytes": l2_fabric, "l2_peak_GBps": 10000.0,
                "l2_peak_source": "profiles/r2_membench.txt: gathers of 128-byte rows out of a 32 MB L2-resident set, 16 warps/SM",
                "duration_ms_under_ncu": t * 1e3})
    if entries:
        json.dump({"entries": entries[-1:]}, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
        print("wrote r2_traffic.json", entries[-1]["dram_bytes_per_launch"] / 1e9, "GB DRAM,", entries[-1]["l2_bytes_per_launch"] / 1e9, "GB L2")


if __name__ == "__main__":
    main()
