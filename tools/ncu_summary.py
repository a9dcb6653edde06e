"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into small text files under profiles/.

  python tools/ncu_summary.py gpurun_out/full_dec1_1_wgrad.ncu-rep [...] --out profiles/r01_ncu_full.txt
  python tools/ncu_summary.py --launches gpurun_out/launches.csv --out profiles/r01_ncu_launches.txt
"""
import argparse
import collections
import csv
import io
import subprocess

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
]


def full(paths, out):
    buf = io.StringIO()
    for p in paths:
        r = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            buf.write(f"== {p}: unreadable\n")
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(units, vals)))
            buf.write(f"== {p}\n   kernel: {d['Kernel Name'][1]}\n")
            for k in KEYS:
                if k in d:
                    buf.write(f"   {k:82s} {d[k][1]:>16s} {d[k][0]}\n")
    open(out, "w").write(buf.getvalue())
    print(buf.getvalue())


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none: {sum(a[0] for a in agg.values())} launches, "
             f"{tot / 1e3:.2f} ms total (serialised, cold-cache; shares are what matter)",
             f"# {'kernel':60s} {'launches':>8s} {'us total':>12s} {'share':>7s} {'us/launch':>10s}"]
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"  {name[:60]:60s} {n:8d} {us:12.1f} {100 * us / tot:6.1f}% {us / n:10.1f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


ap = argparse.ArgumentParser()
ap.add_argument("reps", nargs="*")
ap.add_argument("--launches")
ap.add_argument("--out", required=True)
a = ap.parse_args()
if a.launches:
    launches(a.launches, a.out)
else:
    full(a.reps, a.out)
