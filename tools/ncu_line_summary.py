"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line (samples, instructions)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, '', collections.Counter()])
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or cur is None or not r[0].isdigit() or len(r) < len(hdr): continue
    ix = {}
    for i, h in enumerate(hdr): ix.setdefault(h, i)
    key = (cur, int(r[0]))
    a = agg[key]
    num = lambda v: int(v) if v.lstrip('-').isdigit() else 0
    a[0] += num(r[ix['# Samples']]); a[1] += num(r[ix['Instructions Executed']]); a[2] = r[1].strip()[:80]
    for h in hdr:
        if h.startswith('stall_'): a[3][h] += num(r[ix[h]])
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print('total samples', tot, 'warp-instructions', toti)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ', '.join('%s %d' % (n[6:], c) for n, c in a[3].most_common(2))
    print('%-20s %4d  samp %5d (%4.1f%%)  inst %8d (%4.1f%%)  [%s]  %s' % (k[0], k[1], a[0], 100 * a[0] / max(tot, 1), a[1], 100 * a[1] / max(toti, 1), st, a[2]))
