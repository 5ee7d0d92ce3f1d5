"""Stall samples / instructions of an ncu report aggregated by source function (for kernels built from several headers).
usage: python tools/ncu_funcs.py report.ncu-rep src1 [src2 ...]   (sources as they are NOW: capture and sources must match)"""
import collections
import csv
import re
import subprocess
import sys

rep, srcs = sys.argv[1], sys.argv[2:]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg, cur, fname, ie, isamp = collections.OrderedDict(), None, '', None, None
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
        cur = (fname, int(ln))
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
pat = re.compile(r'\s*(?:template\s*<[^>]*>\s*)?(?:RIG_HD_NI|RIG_HD|RIG_FN|__device__ __forceinline__|__global__|__device__|__host__ __device__ inline)\s+[^;(]*?(\w+)\s*\(')
R = {}
for p in srcs:
    out = []
    for i, l in enumerate(open(p), 1):
        m = pat.match(l)
        if m and not l.strip().startswith('//'):
            out.append((i, m.group(1)))
    R[p.split('/')[-1]] = out
tot, toti = collections.Counter(), collections.Counter()
for (f, ln), (n, s) in agg.items():
    name = f
    if f in R:
        fn = '?'
        for st, nm in R[f]:
            if st <= ln:
                fn = nm
        name = f + ':' + fn
    tot[name] += s
    toti[name] += n
S, I = sum(tot.values()) or 1, sum(toti.values()) or 1
print('total warp instructions', I, 'stall samples', S)
for k, v in tot.most_common(40):
    print(f'{k:58s} {100 * v / S:5.1f}% samples {100 * toti[k] / I:5.1f}% inst')
