"""Per region (split at a marker regex), per mbarrier address: stall samples spent in the try_wait loops, plus the
samples of the code between waits.  Usage: ncu_waits.py src.csv MARKER base_hex name0,name1,...  (names of the 8-byte
barriers starting at base)."""
import csv, re, sys
path, marker, base, names = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), sys.argv[4].split(",")
rows = list(csv.reader(open(path)))
hdr = rows[1]; I = {h: i for i, h in enumerate(hdr)}
region, segs, cur = 0, [], None
def name_of(src):
    m = re.search(r"\+(0x[0-9a-f]+)\]", src)
    if not m: return "?"
    off = int(m.group(1), 16) - base
    k = off // 8
    return names[k] if 0 <= k < len(names) and off % 8 == 0 else m.group(1)
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[I["Source"]].strip()
    if re.search(marker, src): region += 1; cur = None
    s = int(r[I["# Samples"]] or 0)
    if "SYNCS.PHASECHK" in src:
        nm = (region, "WAIT " + name_of(src))
        if cur is None or cur[0] != nm: cur = [nm, 0, 0, 0]; segs.append(cur)
    elif "NANOSLEEP" in src and cur is not None and cur[0][1].startswith("WAIT"):
        pass
    elif cur is None or cur[0][1].startswith("WAIT") and not any(k in src for k in ("BRA", "BREAK", "BSYNC", "BSSY", "NOP", "LDS", "ISETP", "VIADD", "IADD3", "LOP3", "CS2R", "S2UR", "UMOV", "ULEA", "IMAD", "R2UR", "MOV", "PLOP3", "SEL")):
        cur = [(region, "code@" + src[:28]), 0, 0, 0]; segs.append(cur)
    cur[1] += s; cur[2] += 1; cur[3] += "UTCHMMA" in src
tot = {}
for (reg, nm), s, n, mma in segs: tot[reg] = tot.get(reg, 0) + s
for (reg, nm), s, n, mma in segs:
    if s * 200 > tot[reg]:
        print(f"region {reg} {nm:42s} samples {s:8d} ({100.0 * s / tot[reg]:5.1f} % of region) instrs {n:4d} mma {mma}")
