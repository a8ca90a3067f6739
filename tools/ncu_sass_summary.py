"""Summarise an `ncu --page source --csv` dump: stall reasons, opcode mix, hottest instructions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter(); ops = collections.Counter(); samples_by_op = collections.Counter()
inst_total = 0; samp_total = 0
body = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break      # only the first profiled launch
    body.append(r)
for r in body:
    if len(r) < len(hdr): continue
    src = r[ix['Source']].strip(); op = src.split()[0] if src else '?'
    if op.startswith('@'): op = src.split()[1]
    op = op.split('.')[0]
    n = int(r[ix['Instructions Executed']] or 0); s = int(r[ix['# Samples']] or 0)
    ops[op] += n; samples_by_op[op] += s; inst_total += n; samp_total += s
    for c in stall_cols: tot[c] += int(r[ix[c]] or 0)
print('warp-instructions executed: %d, samples: %d' % (inst_total, samp_total))
print('stall reasons (all samples):')
for k, v in tot.most_common(12): print('  %-24s %6d  %5.1f%%' % (k, v, 100.0 * v / max(1, samp_total)))
print('opcode mix (executed / samples):')
for k, v in ops.most_common(18): print('  %-10s %9d %5.1f%%   samples %6d %5.1f%%' % (k, v, 100.0 * v / inst_total, samples_by_op[k], 100.0 * samples_by_op[k] / max(1, samp_total)))
if len(sys.argv) > 2:
    print('hottest instructions:')
    hot = sorted(body, key=lambda r: -int(r[ix['# Samples']] or 0))[:int(sys.argv[2])]
    for r in hot:
        st = {c: int(r[ix[c]] or 0) for c in stall_cols}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print('  %s  %-50s samples %5s  %s' % (r[0][-5:], r[ix['Source']].strip()[:50], r[ix['# Samples']], top))
