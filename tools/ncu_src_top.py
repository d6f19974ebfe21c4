"""Top SASS lines of an ncu report by executed instructions and by stall samples (source page, CSV):
    ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_src_top.py src.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot_inst = sum(int(r[ix['Instructions Executed']]) for r in body)
tot_samp = sum(int(r[ix['# Samples']]) for r in body)
print('total warp instructions %d, samples %d' % (tot_inst, tot_samp))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('stall totals:', sorted(((sum(int(r[ix[s]]) for r in body), s) for s in stalls), reverse=True)[:8])
print('--- by instructions executed')
for li, r in sorted(enumerate(body), key=lambda t: -int(t[1][ix['Instructions Executed']]))[:n]:
    print('%5d %9s %5.1f%%  samples %6s  %s' % (li, r[ix['Instructions Executed']], 100.0 * int(r[ix['Instructions Executed']]) / tot_inst,
                                                r[ix['# Samples']], r[ix['Source']].strip()[:90]))
print('--- by stall samples')
for li, r in sorted(enumerate(body), key=lambda t: -int(t[1][ix['# Samples']]))[:n]:
    top = sorted(((int(r[ix[s]]), s) for s in stalls), reverse=True)[:2]
    print('%5d %7s %5.1f%%  %s  %s' % (li, r[ix['# Samples']], 100.0 * int(r[ix['# Samples']]) / max(tot_samp, 1), r[ix['Source']].strip()[:70], top))
