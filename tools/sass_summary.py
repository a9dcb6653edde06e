"""SASS evidence that the conv path is tcgen05 / TMEM / TMA: per-function counts of UTCHMMA (tcgen05.mma), LDTM
(tcgen05.ld), UTMALDG / UTMASTG (TMA load / store), UTCBAR (tcgen05.commit) in the built library, no mma.sync (HMMA).

  python tools/sass_summary.py [--out profiles/r02_sass_summary.txt]     # CPU only: cuobjdump reads the cubin
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=os.path.join(ROOT, "boxsegliver_b200", "libbsl_b200.so"))
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_sass_summary.txt"))
a = ap.parse_args()
txt = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
pats = {"UTCHMMA": r"\bUTCHMMA", "UTCHMMA.2CTA": r"\bUTCHMMA\.2CTA", "UTMALDG.2CTA": r"\bUTMALDG\S*\.2CTA", "UTCQMMA/UTCIMMA": r"\bUTC[QI]MMA", "LDTM": r"\bLDTM", "UTMALDG": r"\bUTMALDG",
        "UTMASTG": r"\bUTMASTG", "UTCBAR": r"\bUTCBAR", "SYNCS": r"\bSYNCS", "HMMA(mma.sync)": r"\bHMMA",
        "STG.E.256": r"STG\.E\.(ENL2\.)?256"}
tot, rows = collections.Counter(), []
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    c = {k: len(re.findall(p, f)) for k, p in pats.items()}
    for k, v in c.items():
        tot[k] += v
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
        rows.append((name, c))
out = [f"SASS instruction counts of {os.path.relpath(a.lib, ROOT)} (cuobjdump -sass, sm_100a cubin); tools/sass_summary.py",
       "", f"totals over {len(funcs) - 1} functions: {dict(tot)}", "",
       "%-112s %8s %6s %6s %8s %8s %7s" % ("function", "UTCHMMA", ".2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR")]
for name, c in sorted(rows, key=lambda r: -r[1]["UTCHMMA"]):
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(CUtensorMap_st.*", "", dem)
    out.append("%-112s %8d %6d %6d %8d %8d %7d" % (dem[:112], c["UTCHMMA"], c["UTCHMMA.2CTA"], c["LDTM"], c["UTMALDG"], c["UTMASTG"],
                                                   c["UTCBAR"]))
with open(a.out, "w") as fh:
    fh.write("\n".join(out) + "\n")
print("\n".join(out[:8]))
print(f"... {len(rows)} tensor-core functions -> {a.out}")
