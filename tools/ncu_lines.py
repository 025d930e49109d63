"""Aggregate an .ncu-rep source page by CUDA source line (needs -lineinfo): warp-instructions executed, avg lanes,
stall samples.  Usage: ncu_lines.py report.ncu-rep [top_n] [local_source_for_text]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = sys.argv[3] if len(sys.argv) > 3 else "odelib_b200/csrc/odl_kernels.cuh"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
text = open(src).read().splitlines()
cur, hdr, rec = None, None, []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
    elif r and r[0] == "Line No":
        hdr = r
        iE, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    elif hdr and r and r[0].isdigit():
        try:
            rec.append((cur.split("/")[-1], int(r[0]), int(r[iE]), int(r[iT]), int(r[iS] or 0)))
        except ValueError:
            pass
tot = sum(x[2] for x in rec) or 1
tots = sum(x[4] for x in rec) or 1
print(f"total warp-instructions {tot}, samples {tots}")
# by region of the kernels file
for f, ln, e, t, s in sorted(rec, key=lambda x: -x[2])[:top]:
    line = text[ln - 1].strip()[:110] if f.endswith("odl_kernels.cuh") and ln <= len(text) else ""
    print(f"{f:18s}:{ln:5d} {e / tot * 100:6.2f}% inst {s / tots * 100:6.2f}% smpl {t / max(e, 1):5.1f} lanes | {line}")
if len(sys.argv) > 4:
    # ranges "name:lo-hi,..." -> share per range
    for item in sys.argv[4].split(","):
        name, rng = item.split(":")
        lo, hi = map(int, rng.split("-"))
        sel = [x for x in rec if x[0].endswith("odl_kernels.cuh") and lo <= x[1] <= hi]
        e = sum(x[2] for x in sel)
        print(f"range {name:24s} {lo}-{hi}: {e / tot * 100:6.2f}% inst, {sum(x[4] for x in sel) / tots * 100:6.2f}% samples, "
              f"{sum(x[3] for x in sel) / max(e, 1):5.1f} lanes")
