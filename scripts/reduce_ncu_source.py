"""Reduce an `ncu --page source --csv --print-source sass` dump to the rows that matter for shared-memory / stall analysis.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv
    python scripts/reduce_ncu_source.py src.csv [kernel-name substring] > profiles/rNN_src_<kernel>.csv

Keeps, per kernel, every instruction that touches shared memory (LDS / STS / LDGSTS / ATOMS / LDSM) plus the 25 instructions with the
most stall samples, with the columns: SASS, executions, stall samples, shared-memory wavefronts (actual / ideal / excessive, n-way).
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
out = csv.writer(sys.stdout)
for k in kernels:
    if want not in k["name"]:
        continue
    h = k["hdr"]
    ix = {c: i for i, c in enumerate(h)}
    rs = [r for r in k["rows"] if len(r) >= len(h)]
    col = lambda r, c: r[ix[c]] if c in ix else ""
    tot_s = sum(int(col(r, "Warp Stall Sampling (All Samples)") or 0) for r in rs)
    tot_w = sum(int(col(r, "L1 Wavefronts Shared") or 0) for r in rs)
    tot_i = sum(int(col(r, "L1 Wavefronts Shared Ideal") or 0) for r in rs)
    out.writerow(["# kernel", k["name"]])
    out.writerow(["# totals", f"stall samples {tot_s}", f"shared wavefronts {tot_w}", f"ideal {tot_i}",
                  f"instructions executed {sum(int(col(r, 'Instructions Executed') or 0) for r in rs)}"])
    out.writerow(["index", "sass", "executed", "stall_samples", "shared_wavefronts", "shared_wavefronts_ideal", "shared_wavefronts_excessive",
                  "shared_conflict_n_way"])
    top = set(i for i, _ in sorted(enumerate(rs), key=lambda t: -int(col(t[1], "Warp Stall Sampling (All Samples)") or 0))[:25])
    for i, r in enumerate(rs):
        s = r[1].strip()
        shared = any(t in s for t in ("LDS", "STS", "LDGSTS", "ATOMS", "LDSM")) and int(col(r, "Instructions Executed") or 0) > 0
        if shared or i in top:
            out.writerow([i, s[:96], col(r, "Instructions Executed"), col(r, "Warp Stall Sampling (All Samples)"), col(r, "L1 Wavefronts Shared"),
                          col(r, "L1 Wavefronts Shared Ideal"), col(r, "L1 Wavefronts Shared Excessive"), col(r, "L1 Conflicts Shared N-Way")])
