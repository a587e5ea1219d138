"""Top stall lines of each kernel in a source-page CSV (ncu -i rep --page source --csv [--print-source sass|cuda,sass])."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdrs = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdrs.append(len(rows))
for hi in range(len(hdrs) - 1):
    H = rows[hdrs[hi]]
    body = [r for r in rows[hdrs[hi] + 1:hdrs[hi + 1]] if len(r) == len(H)]
    si = H.index('# Samples'); so = H.index('Source'); ie = H.index('Instructions Executed')
    tot = sum(int(r[si] or 0) for r in body)
    print('=' * 110); print('kernel block', hi, 'total samples', tot, 'instructions', len(body))
    stall_cols = [i for i, h in enumerate(H) if h.startswith('stall_')]
    for r in sorted(body, key=lambda r: -int(r[si] or 0))[:top]:
        st = sorted([(int(r[i] or 0), H[i][6:]) for i in stall_cols], reverse=True)[:2]
        print(f"{100*int(r[si] or 0)/max(tot,1):5.1f}%  exec={r[ie]:>10s}  {r[so][:70]:70s} {st}")
