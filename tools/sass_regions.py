"""Summarise an `ncu --page source --print-source sass --csv` dump into regions of
constant execution count (loop bodies), with samples and dominant stall reasons."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def I(x):
    try: return int(x)
    except: return 0
regs = []; cur = None
for r in rows[2:]:
    if len(r) < len(hdr): continue
    e = I(r[ix['Instructions Executed']]); s = I(r[ix['# Samples']])
    if e == 0: continue
    op = r[ix['Source']].split()
    op = (op[1] if op[0].startswith('@') else op[0]).split('.')[0]
    if cur is None or abs(e - cur['e']) > 0.02 * max(e, cur['e']) or op == 'BAR':
        cur = dict(e=e, n=0, samples=0, st=collections.Counter(), ops=collections.Counter(), first=r[ix['Address']])
        regs.append(cur)
    cur['n'] += 1; cur['samples'] += s; cur['ops'][op] += 1
    for h in stalls:
        v = I(r[ix[h]])
        if v: cur['st'][h[6:]] += v
tot = sum(r['samples'] for r in regs)
print('total samples', tot)
for r in regs:
    if r['samples'] < 0.004 * tot: continue
    print('%s exec/instr %8d  n_instr %5d  samples %6d (%4.1f%%)  %s | %s' % (
        r['first'][-6:], r['e'], r['n'], r['samples'], 100.0 * r['samples'] / tot,
        ' '.join('%s:%d' % kv for kv in r['st'].most_common(4)),
        ' '.join('%s:%d' % kv for kv in r['ops'].most_common(6))))
