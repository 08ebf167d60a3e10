"""Segments an `ncu --page source --csv` SASS export by barriers and prints where the stall samples
and the executed instructions fall (first launch in the file only)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
body = []
for r in rows[2:]:
    if r and r[0].startswith('0x'):
        body.append(r)
    elif body:
        break
isrc, isamp, iinst = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
S = [int(r[isamp]) for r in body]
I = [int(r[iinst]) for r in body]
ts, ti = sum(S), sum(I)
print('samples', ts, 'warp-inst', ti, 'static', len(body))
opre = re.compile(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)')
seg_start, segs = 0, []
for k, r in enumerate(body):
    if 'BAR.SYNC' in r[isrc]:
        segs.append((seg_start, k + 1))
        seg_start = k + 1
segs.append((seg_start, len(body)))
KEY = {'LDG', 'STG', 'MATCH', 'LDS', 'STS', 'FFMA', 'SHFL', 'ATOMS', 'F2I', 'MUFU', 'LDGSTS'}
for a, b in segs:
    s, i = sum(S[a:b]), sum(I[a:b])
    ops = set()
    for k in range(a, b):
        m = opre.match(body[k][isrc])
        if m:
            ops.add(m.group(2))
    print(f"seg [{a:5d},{b:5d}) samples {s/ts*100:5.1f}%  inst {i/ti*100:5.1f}%  ops: {' '.join(sorted(ops & KEY))}")
top = sorted(range(len(body)), key=lambda k: -S[k])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]
for k in sorted(top):
    print(f"{S[k]/ts*100:5.2f}% #{k:5d} {body[k][isrc][:100]}")
