import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
body=[]
for r in rows[2:]:
    if r and r[0].startswith('0x'): body.append(r)
    elif body: break
isrc, isamp, iinst = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
opre = re.compile(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)')
cnt = collections.Counter(); smp = collections.Counter()
ti = 0; ts = 0
for r in body:
    m = opre.match(r[isrc]); op = m.group(2) if m else '?'
    n = int(r[iinst]); cnt[op] += n; ti += n
    s = int(r[isamp]); smp[op] += s; ts += s
npart = float(sys.argv[2]) if len(sys.argv) > 2 else 8e6
print('total warp-inst', ti, 'per particle (thread-inst/part)', ti * 32 / npart, 'samples', ts)
for op, n in cnt.most_common(40):
    print(f"{op:10s} {n/ti*100:6.2f}%  {n*32/npart:8.1f}/particle   stall {smp[op]/ts*100:5.1f}%")
