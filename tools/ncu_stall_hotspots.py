import csv, gzip, sys, io
# sections: "Kernel Name" row, header row, instruction rows
path = sys.argv[1]; which = int(sys.argv[2]); topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(gzip.open(path, 'rt')))
secs = []
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name = rows[i][1]; hdr = rows[i+1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
            if len(rows[j]) == len(hdr): body.append(rows[j])
            j += 1
        secs.append((name, hdr, body)); i = j
    else: i += 1
name, hdr, body = secs[which]
print(name[:80], len(body), 'instructions')
ci = {h: k for k, h in enumerate(hdr)}
S = ci['# Samples']; E = ci['Instructions Executed']
tot = sum(int(r[S]) for r in body)
print('total samples', tot, 'total warp-inst', sum(int(r[E]) for r in body))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ci[h]]) for r in body) for h in stall_cols}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
# cumulative samples by position (to find phases)
acc = 0
print('--- top instructions by samples')
idx = sorted(range(len(body)), key=lambda k: -int(body[k][S]))[:topn]
for k in sorted(idx):
    r = body[k]
    st = {h[6:]: int(r[ci[h]]) for h in stall_cols if int(r[ci[h]])}
    print(f'{k:5d} {int(r[S]):6d} {100*int(r[S])/tot:5.1f}%  exec {r[E]:>6s}  {r[1].strip()[:70]:70s} {st}')
# phase summary: samples per 100-instruction bucket
print('--- samples per 100-instruction bucket')
for b in range(0, len(body), 100):
    s = sum(int(r[S]) for r in body[b:b+100]); e = sum(int(r[E]) for r in body[b:b+100])
    print(f'{b:5d} samples {s:6d} ({100*s/tot:4.1f}%) exec {e}')
