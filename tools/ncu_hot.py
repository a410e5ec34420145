"""Stall hot spots of the first kernel in an `ncu --page source --csv --print-source sass` export (gzipped).
usage: python tools/ncu_hot.py file.csv.gz [topn]"""
import csv, gzip, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 24
rows = list(csv.reader(gzip.open(path, 'rt')))
hdr = rows[1]
body = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break
    if len(r) == len(hdr):
        body.append(r)
ci = {h: k for k, h in enumerate(hdr)}
E = ci['Instructions Executed']; S = ci['# Samples']; src = ci.get('Source', 1)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tots = sum(int(r[S]) for r in body)
agg = {h[6:]: sum(int(r[ci[h]]) for r in body) for h in stall_cols}
print(rows[0][1][:70], len(body), 'instructions; samples', tots, 'warp-inst', sum(int(r[E]) for r in body))
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tots * 0.01})
idx = sorted(range(len(body)), key=lambda k: -int(body[k][S]))[:topn]
for k in sorted(idx):
    r = body[k]
    st = {h[6:]: int(r[ci[h]]) for h in stall_cols if int(r[ci[h]]) > 30}
    print(f'{k:5d} {int(r[S]):6d} {100 * int(r[S]) / tots:5.1f}% exec {r[E]:>7s} {r[src].strip()[:56]:56s} {st}')
for b in range(0, len(body), 400):
    seg = body[b:b + 400]
    print(f'   [{b:5d}] samples {sum(int(r[S]) for r in seg):6d} no_inst {sum(int(r[ci["stall_no_inst"]]) for r in seg) if "stall_no_inst" in ci else -1:5d} '
          f'long_sb {sum(int(r[ci["stall_long_sb"]]) for r in seg) if "stall_long_sb" in ci else -1:5d} exec {sum(int(r[E]) for r in seg)}')
