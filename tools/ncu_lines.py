"""Per-source-line instruction / stall-sample shares from `ncu --page source --print-source cuda,sass --csv` of a report.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg, cur, fname = collections.OrderedDict(), None, ''
ie = isamp = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        ie, isamp = r.index('Instructions Executed'), r.index('# Samples')
        continue
    if ie is None or len(r) <= ie:
        continue
    ln = r[0].strip()
    if ln.isdigit():
        cur = (fname, int(ln), r[1].strip())
        agg.setdefault(cur, [0, 0])
    if cur is None or not r[2].strip():
        continue
    try:
        n = int(r[ie].replace(',', '')) if r[ie] else 0
        s = int(r[isamp].replace(',', '')) if r[isamp] else 0
    except ValueError:
        continue
    agg[cur][0] += n
    agg[cur][1] += s
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print('total warp instructions', tot, 'stall samples', tots)
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{f}:{ln:<4d} {100 * n / tot:5.1f}% inst {100 * s / tots:5.1f}% samples  {src[:100]}')
