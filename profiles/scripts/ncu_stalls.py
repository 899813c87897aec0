"""Aggregates the warp-stall samples of an `ncu --page source --csv` export: totals per stall reason and the hottest
SASS lines.  usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv; python ncu_stalls.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = {s: 0 for s in stalls}; lines = []
def num(x):
    try: return int(float(x))
    except: return 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address": continue
    n = num(r[idx['# Samples']])
    for s in stalls: tot[s] += num(r[idx[s]])
    lines.append((n, r[idx['Source']].strip(), {s: num(r[idx[s]]) for s in stalls if num(r[idx[s]]) > 0}, num(r[idx['Instructions Executed']])))
T = sum(tot.values())
print(rows[0][1][:100] if rows[0] else '', '| total samples', T)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]: print(f"  {s:24s} {v:8d} {100 * v / max(T, 1):5.1f} %")
for n, src, st, ex in sorted(lines, key=lambda t: -t[0])[:top]:
    print(f"{n:7d} x{ex:<9d} {src[:84]:84s} {dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])}")
