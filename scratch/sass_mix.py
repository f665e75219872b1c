import csv, sys, collections, re
f=sys.argv[1]
rows=list(csv.reader(open(f)))
hdr=rows[1]; ia=hdr.index('Instructions Executed'); isrc=hdr.index('Source'); isamp=hdr.index('# Samples')
mix=collections.Counter(); samp=collections.Counter(); tot=0
for r in rows[2:]:
    if len(r)<=ia: continue
    try: n=int(r[ia])
    except: continue
    src=r[isrc].strip()
    m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)',src)
    op=m.group(2) if m else src[:10]
    op='.'.join(op.split('.')[:2]) if op.startswith(('LDS','STS','LDG','STG','SHFL','BAR','MUFU','F2','I2')) else op.split('.')[0]
    mix[op]+=n; tot+=n
    try: samp[op]+=int(r[isamp])
    except: pass
print('total warp instr',tot)
ts=sum(samp.values())
for op,n in mix.most_common(28): print(f"{op:14s} {n:10d} {100*n/tot:5.1f}%   samples {100*samp[op]/max(ts,1):5.1f}%")
