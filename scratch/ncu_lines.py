"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by CUDA source line."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None; cur_file = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])
tot = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; S = hdr.index("# Samples"); E = hdr.index("Instructions Executed")
        st = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
        continue
    if hdr is None or len(r) < len(hdr): continue
    try: smp = int(r[S])
    except ValueError: continue
    key = (cur_file, r[0])
    a = agg[key]; a[0] += smp; a[1] += int(r[E] or 0); a[3] = r[1].strip()[:90]
    for i, h in st:
        try: a[2][h] += int(r[i])
        except ValueError: pass
    tot += smp
print("total samples", tot)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0]}:{key[1]:>4s} {100*a[0]/tot:5.1f}% exec={a[1]:>10d} {[(h[6:], v) for h, v in a[2].most_common(2)]} | {a[3]}")
