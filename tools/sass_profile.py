"""Summarise an `ncu --page source --csv` dump: executed warp-instructions per warp, per SASS line."""
import csv, sys
path, nwarps = sys.argv[1], float(sys.argv[2])
full = len(sys.argv) > 3
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
ia, isrc, ie = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
ist = hdr.index("Warp Stall Sampling (All Samples)")
tot = 0; out = []
for d in data:
    try: e = float(d[ie])
    except Exception: continue
    tot += e
    out.append((d[isrc], e / nwarps, float(d[ist] or 0)))
print("executed warp-instr per warp:", round(tot / nwarps, 1), " SASS lines:", len(out))
ssum = sum(o[2] for o in out)
from collections import Counter
ops = Counter(); stall = Counter()
for s, f, st in out:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    op = op.split(".")[0]
    ops[op] += f; stall[op] += st
print("by opcode (instr/warp, stall-sample %):")
for op, f in ops.most_common(25):
    print(f"  {op:10s} {f:7.1f}  {100*stall[op]/max(ssum,1):5.1f}%")
if full:
    for s, f, st in out:
        print(f"{f:6.2f} {100*st/max(ssum,1):5.2f}% {s[:120]}")
