#!/usr/bin/env python
"""Per-kernel SASS statistics of a built library: instruction count, FP64 / TMA / spill instruction counts."""
import re, subprocess, sys
so = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name = None; st = {}
for ln in txt.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1); st[name] = dict(n=0, f64=0, tma=0, spill=0, shfl=0, lds=0, ldg=0, imad=0, bar=0)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", ln)
    if m and name:
        ins = m.group(1)
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        d = st[name]; d["n"] += 1
        if re.match(r"D(FMA|ADD|MUL|SETP)", op): d["f64"] += 1
        if op.startswith("UTMALDG") or op.startswith("UBLKCP"): d["tma"] += 1
        if op.startswith("STL") or op.startswith("LDL"): d["spill"] += 1
        if op.startswith("SHFL"): d["shfl"] += 1
        if op.startswith("LDS") or op.startswith("STS"): d["lds"] += 1
        if op.startswith("LDG") or op.startswith("STG"): d["ldg"] += 1
        if op.startswith("IMAD") or op.startswith("IADD") or op.startswith("LEA"): d["imad"] += 1
for k, d in st.items():
    if pat in k:
        out = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        print(f"{out[:80]:80s} instr {d['n']:6d} f64 {d['f64']:5d} shfl {d['shfl']:4d} lds/sts {d['lds']:4d} ldg/stg {d['ldg']:4d} int {d['imad']:5d} tma {d['tma']:2d} spill {d['spill']:4d}")
