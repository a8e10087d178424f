"""Summarise an ncu --page source --csv dump: stall reasons overall, by opcode, top instructions."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
S = hdr.index('# Samples'); stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[S]) for r in data)
print("total samples", tot)
agg = collections.Counter()
for r in data:
    for i, h in stall:
        agg[h] += int(r[i])
print("stall reasons:", [(h, round(100 * v / tot, 1)) for h, v in agg.most_common(8)])
byop = collections.Counter(); cnt = collections.Counter(); execd = collections.Counter()
for r in data:
    op = r[1].strip().split()[0] if r[1].strip() else '?'
    if op.startswith('@'): op = r[1].strip().split()[1]
    byop[op] += int(r[S]); cnt[op] += 1; execd[op] += int(r[5])
print("by opcode (samples%, static count, executed):")
for op, v in byop.most_common(18):
    print(f"   {op:28s} {100*v/tot:5.1f}%  n={cnt[op]:4d} exec={execd[op]}")
print("top instructions:")
idx = sorted(range(len(data)), key=lambda i: -int(data[i][S]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in idx:
    r = data[i]
    top = sorted(((int(r[j]), h) for j, h in stall), reverse=True)[:2]
    print(f"   #{i:5d} {int(r[S]):6d} {100*int(r[S])/tot:4.1f}%  {r[1].strip()[:70]:70s} {top}")
