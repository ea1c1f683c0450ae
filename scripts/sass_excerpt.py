"""Writes profiles/<tag>_sass_excerpts.txt: per hot kernel the SASS mnemonic histogram and the lines that show the
async-copy (LDGSTS / UTMALDG), tensor (UTCHMMA / LDTM), packed-fp32 (FFMA2 / FADD2) and mbarrier (SYNCS) paths.
usage: python scripts/sass_excerpt.py [tag]   (build container: cuobjdump reads the in-tree library)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r7"
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "chrysalis_b200", "libchrysalis_b200.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
want = {"moran_lattice_kernel (TMA producer, default)": "moran_lattice_kernelILb1ELi2EE", "moran_lattice_kernel (CSR-fed)": "moran_lattice_kernelILb1ELi1EE", "moran_lattice_perm_kernel (bulk-copy producer, default)": "moran_lattice_perm_kernelILb1EE",
        "gram_tcgen05_super_kernel (default)": "gram_tcgen05_super_kernel", "gram_tcgen05_kernel (variant 2)": "gram_tcgen05_kernelE", "moran_tile_kernel (GEARY=0,SHIFT=1,MAPPED=0)": "moran_tile_kernelILb0ELb1ELb0E"}
pat = re.compile(r"/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)")
out = ["SASS excerpts of the hot kernels in chrysalis_b200/libchrysalis_b200.so (cuobjdump -sass, sm_100a; mnemonic histograms + the lines that",
       "show the async-copy / tensor / packed-fp32 paths).  Regenerate with scripts/sass_excerpt.py after a build.", ""]
for label, key in want.items():
    f = [x for x in funcs if key in x.split("\n")[0]]
    if not f:
        out.append(f"== {label}: not found")
        continue
    ops, lines = collections.Counter(), []
    for ln in f[0].split("\n"):
        m = pat.search(ln)
        if m:
            op = m.group(2)
            ops[op.split(".")[0]] += 1
            if re.match(r"(LDGSTS|UTMALDG|UTCHMMA|UTCQMMA|LDTM|STTM|FFMA2|FADD2|SYNCS|UBLKCP|LDS\.|UTCBAR|ARRIVES)", op):
                lines.append(ln.strip()[:150])
    out.append(f"== {label}  [{f[0].split(chr(10))[0][:100]}]")
    out.append("   instruction counts: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(28)))
    seen = collections.Counter()
    for ln in lines:
        op = pat.search(ln).group(2)
        if seen[op] < 2:
            out.append("   " + ln)
            seen[op] += 1
    out.append("")
open(os.path.join(ROOT, "profiles", f"{tag}_sass_excerpts.txt"), "w").write("\n".join(out))
print("\n".join(out))
