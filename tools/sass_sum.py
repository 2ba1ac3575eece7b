import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
acc=collections.Counter(); cnt=collections.Counter(); exe=collections.Counter()
tot=0; totexe=0
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
bystall=collections.Counter()
def I(x):
    try: return int(x)
    except: return 0
for r in rows[2:]:
    if len(r)<len(hdr) or r[0]=='Address': continue
    src=r[idx['Source']].strip()
    parts=src.split()
    op=parts[1] if parts[0].startswith('@') else parts[0]
    op=op.split('.')[0]
    n=I(r[idx['# Samples']])
    e=I(r[idx['Instructions Executed']])
    acc[op]+=n; cnt[op]+=1; exe[op]+=e; tot+=n; totexe+=e
    for h in stall_cols:
        v=I(r[idx[h]])
        if v: bystall[(op,h)]+=v
print('total samples',tot,'sass lines',len(rows)-2,'executed warp instr',totexe)
for op,n in acc.most_common(18): print('%-10s samples %6d (%.1f%%)  static %6d  executed %10d (%.1f%%)'%(op,n,100*n/tot,cnt[op],exe[op],100*exe[op]/totexe))
print()
for (op,h),v in bystall.most_common(14): print(op,h,v)
