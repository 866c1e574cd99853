"""SASS opcode histogram of the shipped library (runs here: cuobjdump needs no GPU) -> profiles/<tag>_sass_histogram.txt

    python scripts/sass_histogram.py [tag]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "torch-em_b200", "libb200em.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "HMMA", "IMMA", "ATOMG", "RED", "ATOMS",
         "FFMA2", "FMUL2", "FADD2", "CCTL", "REDUX"]
tot = collections.Counter()
per = []
cur = None
i = -1
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        i += 1
        cur = collections.Counter()
        per.append((names[i], cur))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                tot[w] += 1
                cur[w] += 1
                break
out = [f"# SASS opcode histogram of torch-em_b200/libb200em.so (cuobjdump -sass, sm_100a), whole library: {len(per)} kernels", ""]
out += [f"{w:12s} {tot[w]}" for w in WATCH]
out += ["", "# per kernel: tcgen05 MMA (UTCHMMA*), TMEM loads (LDTM*), TMA tensor loads (UTMALDG*), bulk copies (UBLKCP*), tcgen05.commit (UTCBAR*), packed fp32x2 math (FFMA2/FMUL2/FADD2)", ""]
for name, c in sorted(per):
    if any(c[w] for w in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "FFMA2")):
        out.append(f"{name[:118]:118s} UTCHMMA {c['UTCHMMA']:4d}  LDTM {c['LDTM']:3d}  UTMALDG {c['UTMALDG']:3d}  UBLKCP {c['UBLKCP']:3d}  UTCBAR {c['UTCBAR']:3d}  FFMA2 {c['FFMA2']:4d}")
path = os.path.join(ROOT, "profiles", f"{tag}_sass_histogram.txt")
open(path, "w").write("\n".join(out) + "\n")
print(path, dict(tot))
