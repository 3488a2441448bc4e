"""Summarise `ncu -i X.ncu-rep --page source --csv` output: stall samples per region (regions split at SASS
instructions matching a marker regex) and the top instructions of each region with their dominant stall reasons."""
import csv, re, sys

path, marker = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else r"USETMAXREG")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
rows = list(csv.reader(open(path)))
hdr = rows[1]
I = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
regions, cur = [], {"name": "prologue", "rows": []}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    if re.search(marker, r[I["Source"]]):
        regions.append(cur)
        cur = {"name": r[I["Source"]].strip()[:60], "rows": []}
    cur["rows"].append(r)
regions.append(cur)
total = sum(int(r[I["# Samples"]] or 0) for reg in regions for r in reg["rows"])
print(f"total samples {total}")
for reg in regions:
    n = sum(int(r[I["# Samples"]] or 0) for r in reg["rows"])
    ex = sum(int(r[I["Instructions Executed"]] or 0) for r in reg["rows"])
    agg = {c: sum(int(r[I[c]] or 0) for r in reg["rows"]) for c in stall_cols}
    tops = sorted(agg.items(), key=lambda kv: -kv[1])[:5]
    print(f"\n== region [{reg['name']}]: {n} samples ({100.0 * n / max(total, 1):.1f} %), {ex} warp instr; " +
          ", ".join(f"{k[6:]} {v}" for k, v in tops if v))
    for r in sorted(reg["rows"], key=lambda r: -int(r[I["# Samples"]] or 0))[:top]:
        s = int(r[I["# Samples"]] or 0)
        if not s:
            break
        rs = sorted(((c[6:], int(r[I[c]] or 0)) for c in stall_cols), key=lambda kv: -kv[1])[:2]
        print(f"   {s:6d}  {r[I['Source']].strip()[:70]:70s} {rs}")
