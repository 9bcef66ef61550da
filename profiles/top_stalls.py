"""Summarise `ncu -i X.ncu-rep --page source --csv` output: per kernel section, the top lines by stall samples."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
only = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else None
sections = []; cur = None
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == 'Kernel Name':
        cur = {'name': row[1], 'hdr': None, 'rows': []}; sections.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr'] = row; continue
    cur['rows'].append(row)
for si, s in enumerate(sections):
    if only and si not in only: continue
    h = {n: i for i, n in enumerate(s['hdr'])}
    def f(r, n):
        try: return float(r[h[n]])
        except Exception: return 0.0
    tot = sum(f(r, '# Samples') for r in s['rows'])
    inst = sum(f(r, 'Instructions Executed') for r in s['rows'])
    print(f"== section {si}: {s['name']}  rows={len(s['rows'])} samples={tot:.0f} warp-instr={inst:.0f}")
    stalls = [n for n in s['hdr'] if n.startswith('stall_') and 'Not Issued' not in n]
    agg = {n: sum(f(r, n) for r in s['rows']) for n in stalls}
    print('   stall totals:', ', '.join(f"{k[6:]}={v:.0f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(s['rows'], key=lambda r: -f(r, '# Samples'))[:topn]:
        top = sorted(((f(r, n), n[6:]) for n in stalls), reverse=True)[:2]
        print(f"  {f(r,'# Samples')/max(tot,1)*100:5.1f}%  exec={f(r,'Instructions Executed'):9.0f} shw={f(r,'L1 Wavefronts Shared'):9.0f}  {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f} | {r[h['Source']][:110]}")
