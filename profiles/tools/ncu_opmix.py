import csv,sys,re,collections
rows=list(csv.reader(open('/tmp/src.csv')))
hdr=rows[1]
ix={h:i for i,h in enumerate(hdr)}
tot=collections.Counter(); samples=collections.Counter()
stall=collections.Counter()
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
total_inst=0; total_samp=0
data=[]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    src=r[ix['Source']].strip()
    m=re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_]+)',src)
    op=m.group(2) if m else src
    n=int(r[ix['Instructions Executed']] or 0)
    s=int(r[ix['# Samples']] or 0)
    tot[op]+=n; samples[op]+=s; total_inst+=n; total_samp+=s
    for c in stall_cols: stall[c]+=int(r[ix[c]] or 0)
    data.append((r[ix['Address']],src,n,s))
print('total warp inst',total_inst,'samples',total_samp)
for op,n in tot.most_common(30): print(f'{op:10s} {n:12d} {100*n/total_inst:5.1f}%  samples {100*samples[op]/total_samp:5.1f}%')
print()
for c,n in stall.most_common(12): print(c,n,f'{100*n/total_samp:.1f}%')
import pickle; pickle.dump(data,open('/tmp/src.pkl','wb'))
