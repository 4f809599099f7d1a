"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: python scripts/ncu_lines.py dump.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; continue
    if r[0] == 'Function Name' or hdr is None:
        continue
    if r[0] != '' and r[2] == '-':     # a CUDA source line with aggregated metrics
        d = dict(zip(hdr, r))
        def f(k):
            try: return float(d.get(k, '0') or 0)
            except ValueError: return 0.0
        out.append((cur_file, r[0], r[1], f('Warp Stall Sampling (All Samples)'), f('Instructions Executed'), d))
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out)
print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    d = o[5]
    stalls = {k[6:]: float(v) for k, v in d.items() if k.startswith('stall_') and '(Not' not in k and v not in ('', '-', '0')}
    s3 = ", ".join(f"{k}:{v:.0f}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
    print(f"{o[0]:>14s}:{o[1]:>4s} samp {100*o[3]/ts:5.1f}% inst {100*o[4]/ti:5.1f}% [{s3}] {o[2].strip()[:110]}")
