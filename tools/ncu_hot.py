"""Per-instruction stall hot spots from `ncu -i X.ncu-rep --page source --csv --kernel-id :::N` (stdin)."""
import csv, sys
thr = float(sys.argv[1]) if len(sys.argv) > 1 else 0.006
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
print(rows[0][1][:110] if rows[0] else "")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[0] == 'Address':
        break
    data.append(r)
tot = sum(int(r[idx['# Samples']]) for r in data)
cols = ['stall_long_sb', 'stall_wait', 'stall_math', 'stall_dispatch', 'stall_not_selected', 'stall_selected', 'stall_short_sb',
        'stall_no_inst', 'stall_mio', 'stall_lg', 'stall_branch_resolving', 'stall_barrier']
print('total samples', tot, 'instructions', len(data))
print('totals', {c[6:]: sum(int(r[idx[c]]) for r in data) for c in cols})
for n, r in enumerate(data):
    s = int(r[idx['# Samples']])
    if s > tot * thr:
        print(f"{n:4d} {r[idx['Source']].strip()[:64]:64s} {s:6d} x{r[idx['Instructions Executed']]:>9s} " +
              ' '.join(f'{c[6:]}={r[idx[c]]}' for c in cols if int(r[idx[c]]) > s * 0.15))
