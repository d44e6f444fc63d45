"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv; python scripts/ncu_lines.py x.csv [top]"""
import csv, sys, collections
# ncu does not escape quotes inside the source column (inline asm): split on the field separator instead of csv.reader
rows = [l.rstrip('\n')[1:-1].split('","') if l.startswith('"') else next(csv.reader([l])) for l in open(sys.argv[1]) if l.strip()]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = collections.OrderedDict()
src = {}
cur_file = None
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"):
        cur_file = r[1]; continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not cur_file or 'cmpc' not in cur_file:
        continue
    d = dict(zip(hdr, r))
    try:
        ln = int(r[0])
    except ValueError:
        continue
    a = agg.setdefault(ln, dict(samples=0, inst=0, sass=0, text=r[1], stalls=collections.Counter()))
    num = lambda v: int(v) if v not in ('', '-') else 0
    a['samples'] += num(d['# Samples'])
    a['inst'] += num(d['Instructions Executed'])
    for k in hdr:
        if k.startswith('stall_') and 'Not Issued' not in k:
            a['stalls'][k] += num(d[k])
tot_s = sum(a['samples'] for a in agg.values()); tot_i = sum(a['inst'] for a in agg.values())
print(f"total samples {tot_s}, total warp instructions {tot_i}")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top]:
    st = ' '.join(f"{k[6:]}:{v}" for k, v in a['stalls'].most_common(3))
    print(f"{ln:4d} smp {100*a['samples']/max(tot_s,1):5.1f}% inst {100*a['inst']/max(tot_i,1):5.1f}% sass {a['sass']:4d} | {a['text'].strip()[:70]:70s} | {st}")
