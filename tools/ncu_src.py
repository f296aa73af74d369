"""Summarise an `ncu --page source --csv` dump: opcode mix and hottest SASS lines."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1]))]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) == len(hdr):
        data.append(r)
F = lambda r, k: float(r[idx[k]] or 0)
tot_inst = sum(F(r, 'Instructions Executed') for r in data); tot_samp = sum(F(r, '# Samples') for r in data)
print("total warp-inst %.4g  samples %d  sass lines %d" % (tot_inst, tot_samp, len(data)))
op = defaultdict(lambda: [0, 0])
for r in data:
    s = r[idx['Source']].strip().split()
    if not s: continue
    o = s[0] if not s[0].startswith('@') else (s[1] if len(s) > 1 else s[0])
    o = o.split('.')[0]
    op[o][0] += F(r, 'Instructions Executed'); op[o][1] += F(r, '# Samples')
for k, v in sorted(op.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%-12s inst %5.1f%%  samples %5.1f%%" % (k, 100 * v[0] / tot_inst, 100 * v[1] / tot_samp))
print("--- hottest lines by stall samples")
for r in sorted(data, key=lambda r: -F(r, '# Samples'))[:22]:
    print("%6s %5.1f%% inst %9.4g  %s" % (r[idx['Address']][-5:], 100 * F(r, '# Samples') / tot_samp, F(r, 'Instructions Executed'), r[idx['Source']][:100]))
