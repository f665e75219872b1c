import csv,collections,re,sys
rows=list(csv.reader(open(sys.argv[1])))
thr=int(sys.argv[2]) if len(sys.argv)>2 else 1000
hdr=rows[1]; ia=hdr.index('Instructions Executed'); isrc=hdr.index('Source'); isamp=hdr.index('# Samples')
cur=None; grp=[]; out=[]
def flush():
    global grp,cur
    if not grp: return
    c=collections.Counter(); s=0
    for src,sm in grp:
        m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)',src.strip()); c[m.group(2) if m else src[:8]]+=1; s+=sm
    out.append((cur,len(grp),s,' '.join(f"{k}x{v}" for k,v in c.most_common(16))))
    grp=[]
for r in rows[2:]:
    try: n=int(r[ia]); sm=int(r[isamp])
    except: continue
    if n!=cur: flush(); cur=n
    grp.append((r[isrc],sm))
flush()
tot=sum(o[0]*o[1] for o in out); ts=sum(o[2] for o in out)
print('total',tot,'samples',ts)
for o in out:
    if o[0]*o[1]>=thr*50 or o[2]>ts*0.01: print(f"exec={o[0]:8d} n={o[1]:4d} ({100*o[0]*o[1]/tot:4.1f}%) samp={100*o[2]/ts:4.1f}% : {o[3]}")
