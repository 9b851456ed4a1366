"""profiles/traffic_r2.json + profiles/launches_r2_summary.json from an ncu CSV of ONE forward+backward step
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv ... tools/prof_step.py 64 bilinear fast 1`).
usage: ncu_traffic.py <csv> <out.json>"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
idi = h.index("ID")
launches = {}
for r in rows[start + 1:]:
    if len(r) <= vi or not r[idi].isdigit():
        continue
    d = launches.setdefault(int(r[idi]), {"kernel": r[ki].split("(")[0].replace("void ", "")})
    val = float(r[vi].replace(",", ""))
    unit = r[ui]
    if r[mi].startswith("dram__bytes"):
        val *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    if r[mi].startswith("gpu__time"):
        val *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1)      # -> ms
    d[r[mi]] = val
ours = [d for _, d in sorted(launches.items()) if any(k in d["kernel"] for k in ("sl_", "pole_", "rows_", "plane_reach"))]
# the LAST step in the capture: forward = up to the first pole_rows_fix, backward = the rest
names = [d["kernel"] for d in ours]
last_fwd = max(i for i, n in enumerate(names) if "sl_fwd" in n)
start_idx = max(i for i in range(last_fwd + 1) if "pole_means" in names[i] and i <= last_fwd)
nxt = [i for i in range(last_fwd + 1, len(names)) if "sl_fwd" in names[i]]
step = ours[start_idx:]
fwd_end = last_fwd - start_idx + 2          # pole_means, sl_fwd, pole_rows_fix
phases = {"forward": step[:fwd_end], "backward": step[fwd_end:]}
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum (cold cache, serialised launches)",
       "phases": {}}
for ph, ks in phases.items():
    tot = sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in ks)
    ms = sum(k.get("gpu__time_duration.sum", 0) for k in ks)
    out["phases"][ph] = {"dram_bytes": tot, "ms_serialised": ms,
                         "kernels": [{"kernel": k["kernel"], "ms": round(k.get("gpu__time_duration.sum", 0), 5),
                                      "dram_bytes": k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0),
                                      "share_of_step": None} for k in ks]}
total_ms = sum(p["ms_serialised"] for p in out["phases"].values())
for p in out["phases"].values():
    for k in p["kernels"]:
        k["share_of_step"] = round(k["ms"] / total_ms, 4)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({ph: (p["dram_bytes"], round(p["ms_serialised"], 4)) for ph, p in out["phases"].items()}))
