#!/usr/bin/env python
"""Per-source-line stall samples of one kernel from an .ncu-rep (captured with --import-source on).
usage: ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; lines = []; hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] not in ('', '-') and hdr:
        d = dict(zip(hdr[4:], r[4:]))
        lines.append((cur, r[0], r[1], d))
def f(x):
    try: return float(x)
    except: return 0.0
tot = sum(f(l[3].get('# Samples')) for l in lines); toti = sum(f(l[3].get('Instructions Executed')) for l in lines)
print('total samples %d, warp instructions %.3e' % (tot, toti))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for fn, ln, src, d in lines:
    s = f(d.get('# Samples'))
    if s >= minp / 100 * tot:
        top = sorted(((f(d.get(k)), k[6:]) for k in stalls), reverse=True)[:3]
        print('%-22s %4s %5.1f%% smp %5.1f%% inst  %-40s | %s' % (fn[:22], ln, 100 * s / tot, 100 * f(d.get('Instructions Executed')) / toti,
              ' '.join('%s=%.0f%%' % (k, 100 * v / max(s, 1)) for v, k in top), src.strip()[:90]))
