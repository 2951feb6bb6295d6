"""Prints the hot SASS lines of one kernel from an ncu source-page CSV.
usage: ncu -i rep --page source --csv --kernel-name regex:NAME > f.csv; python scripts/ncu_hot.py f.csv [frac]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
want = sys.argv[3] if len(sys.argv) > 3 else ''
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
for a, b in zip(starts[:-1], starts[1:]):
    if want in rows[a][1]:
        rows = rows[a:b]
        break
print(rows[0][1][:80])
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
si, ci, ei = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
body = [r for r in rows[hi + 1:] if len(r) > ci]
tot = sum(float(r[ci] or 0) for r in body)
print('total samples', tot, 'instructions', len(body))
agg = {h[i]: sum(float(r[i] or 0) for r in body) for i in stall}
print('stall totals:', {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for n, r in enumerate(body):
    v = float(r[ci] or 0)
    if v > tot * frac:
        top = sorted([(float(r[i] or 0), h[i][6:]) for i in stall], reverse=True)[:2]
        print(f"{n:5d} {v:8.0f} {100*v/tot:5.1f}% ex={r[ei]:>9} {r[si][:70]:70s} {top}")
