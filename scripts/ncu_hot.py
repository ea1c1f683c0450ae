"""Where a kernel's samples fall, per CUDA source line, from an .ncu-rep captured with --import-source on.
usage: python scripts/ncu_hot.py REPORT.ncu-rep [top-n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]
i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0] not in ("", "Line No")]      # per-line aggregate rows
num = lambda x: int(x) if x.isdigit() else 0
tot = sum(num(r[i_s]) for r in data) or 1
tex = sum(num(r[i_e]) for r in data) or 1
print("total samples", tot, "warp instructions", tex)
best = sorted(data, key=lambda r: -num(r[i_s]))[:topn]
for r in sorted(best, key=lambda r: int(r[0])):
    print(f'{r[0]:>5} {100 * num(r[i_s]) / tot:5.1f}% smp {100 * num(r[i_e]) / tex:5.1f}% ins  {r[1][:120]}')
