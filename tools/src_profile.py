"""Per CUDA-source-line executed instructions from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
path = sys.argv[1]; denom = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
cur_file = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if len(r) > 7 and r[2] == "-":   # a cuda source line summary
        try: out.append((cur_file, int(r[0]), r[1].strip(), float(r[7] or 0), float(r[4] or 0)))
        except ValueError: pass
tot = sum(o[3] for o in out); st = sum(o[4] for o in out)
print("total instr", tot, "per unit", tot / denom)
for f, ln, src, n, s in sorted(out, key=lambda o: -o[3])[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    print(f"{n/denom:8.1f} {100*n/tot:5.1f}% stall {100*s/max(st,1):5.1f}%  {f}:{ln}  {src[:90]}")
