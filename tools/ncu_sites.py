"""Wait sites of a warp-specialised kernel from `ncu --page source --csv` (SASS): each run of unrolled
mbarrier try_wait polls is one site; prints its stall samples, how often it was entered, and the landmark
instructions (MMA, TMEM load / store, bulk copy, shared-memory stores, arrives) between the sites with their samples."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; I = {h: i for i, h in enumerate(hdr)}
data = [(r[I['Source']].strip(), int(r[I['# Samples']] or 0), int(r[I['Instructions Executed']] or 0))
        for r in rows[2:] if len(r) >= len(hdr)]
total = sum(d[1] for d in data)
LAND = re.compile(r'UTCHMMA|UTCBAR|UBLKCP|STTM|LDTM|SYNCS\.ARRIVE|BAR\.SYNC|EXIT|FENCE|STG|LDS|STS')
i, n = 0, len(data)
work, marks = 0, {}
def flush():
    global work, marks
    if work or marks:
        print(f"      work {work:5d} ({100 * work / total:4.1f} %)  " + ", ".join(f"{k} x{v[0]} s={v[1]}" for k, v in marks.items()))
    work, marks = 0, {}
while i < n:
    src, s, ex = data[i]
    if 'TRYWAIT' in src:
        flush()
        j, last, tot, entered = i, i, 0, ex
        while j < n and j - last <= 6:
            if 'TRYWAIT' in data[j][0]: last = j
            j += 1
        j = last + 2                      # the poll and its branch
        tot = sum(d[1] for d in data[i:j])
        polls = sum(d[2] for d in data[i:j] if 'TRYWAIT' in d[0])
        print(f"WAIT site @{i}: samples {tot:5d} ({100 * tot / total:4.1f} %), entered {entered}, polls {polls}")
        i = j
        continue
    work += s
    m = LAND.search(src)
    if m:
        k = m.group(0)
        c = marks.setdefault(k, [0, 0]); c[0] += 1; c[1] += s
    i += 1
flush()
print("total", total)
