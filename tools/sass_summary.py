#!/usr/bin/env python
"""Per-kernel SASS opcode counts of the product objects (math_atlas_b200/csrc/obj/*.o): the evidence that the kernels use
what DESIGN.md says they use (DMMA / UTMALDG / UTCHMMA / LDTM / SYNCS ...).  python tools/sass_summary.py > profiles/r02_sass_opcodes.txt"""
import os, re, subprocess, sys, collections
OBJ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "math_atlas_b200", "csrc", "obj")
KEY = ["DMMA", "DFMA", "FFMA", "HMMA", "UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "LDS", "STS", "LDG", "STG", "LDSM",
       "REDUX", "SHFL", "BAR", "MEMBAR", "FENCE", "ATOM", "RED", "UCGABAR", "CCTL", "ERRBAR"]
print("# kernel (demangled, shortened) | registers | opcode counts (only non-zero of: " + " ".join(KEY) + ")")
for f in sorted(os.listdir(OBJ)):
    if not f.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
    cur, counts = None, collections.OrderedDict()
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1); counts[cur] = collections.Counter(); continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur:
            op = m.group(1)
            for k in KEY:
                if op == k or op.startswith(k + ".") or op.startswith(k):
                    counts[cur][k] += 1
                    break
    names = list(counts)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    print(f"\n## {f}")
    for n, d in zip(names, dem):
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"\(.*", "", d)
        d = d.replace("void ", "").replace("ab::", "")
        c = counts[n]
        if not any(c.values()):
            continue
        print(f"{d[:86]:86s} | {regs.get(n, '?'):>3s} | " + " ".join(f"{k}={c[k]}" for k in KEY if c[k]))
