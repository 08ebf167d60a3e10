import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
npart=8e6
cur=None; out=[]
hdr=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; ii=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples'); continue
    if r[0] and r[0].isdigit():
        try: out.append((cur, int(r[0]), r[1], int(r[ii]), int(r[isamp])))
        except ValueError: pass
ti=sum(o[3] for o in out); ts=sum(o[4] for o in out)
print('total', ti, ti*32/npart, 'samples', ts)
byfile={}
for f,l,s,i,sm in out: byfile[f]=byfile.get(f,0)+i
for f,i in byfile.items(): print(f.split('/')[-1], i*32/npart)
mode=sys.argv[2] if len(sys.argv)>2 else 'top'
if mode=='top':
    for f,l,s,i,sm in sorted(out,key=lambda o:-o[3])[:int(sys.argv[3]) if len(sys.argv)>3 else 60]:
        print(f"{i*32/npart:7.1f} {sm/ts*100:5.1f}% {f.split('/')[-1][:14]:14s}:{l:5d} {s.strip()[:110]}")
else:
    for f,l,s,i,sm in out:
        if i*32/npart>=float(sys.argv[3]): print(f"{i*32/npart:7.1f} {sm/ts*100:5.1f}% {f.split('/')[-1][:14]:14s}:{l:5d} {s.strip()[:110]}")
