"""Condense a tools/ncu_summary.py text (one block per launch) into one table row per launch.

  python tools/ncu_table.py gpurun_out/r02_ncu_full_cfg4_conv.txt --out profiles/r02_ncu_full_cfg4_conv_summary.txt [--header FILE]
"""
import argparse
import re

ap = argparse.ArgumentParser()
ap.add_argument("src")
ap.add_argument("--out", required=True)
ap.add_argument("--header", default="")
a = ap.parse_args()
rows, cur = [], None
for line in open(a.src):
    if line.startswith("=="):
        cur = {}
        rows.append(cur)
    elif cur is not None:
        m = re.match(r"\s+kernel: (.*)", line)
        if m:
            cur["kernel"] = re.sub(r"\(CUtensorMap.*", "", m.group(1)).replace("void ", "").replace("bsl::", "")
            continue
        p = line.split()
        if len(p) >= 2:
            cur[p[0]] = (p[1], p[2] if len(p) > 2 else "")


def val(r, k, scale=None):
    if k not in r:
        return float("nan")
    v, u = r[k]
    v = float(v.replace(",", ""))
    if scale == "us":
        v *= {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1)
    if scale == "MB":
        v *= {"Mbyte": 1, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1)
    return v


out = []
if a.header:
    out.append(open(a.header).read().rstrip())
out.append(f"{'#':>3s} {'kernel':58s} {'us':>8s} {'grid':>5s} {'tensor%':>8s} {'smem-tc%':>8s} {'L2%':>6s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'dram TB/s':>9s}")
for i, r in enumerate(rows):
    us = val(r, "gpu__time_duration.sum", "us")
    rd, wr = val(r, "dram__bytes_read.sum", "MB"), val(r, "dram__bytes_write.sum", "MB")
    out.append(f"{i:3d} {r.get('kernel', '?')[:58]:58s} {us:8.1f} {val(r, 'launch__grid_size'):5.0f} "
               f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):8.1f} "
               f"{val(r, 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):8.1f} "
               f"{val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {rd:11.1f} {wr:11.1f} "
               f"{(rd + wr) / us if us else float('nan'):9.2f}")
open(a.out, "w").write("\n".join(out) + "\n")
print("\n".join(out))
