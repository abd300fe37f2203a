#!/usr/bin/env python
"""Aggregate an ncu source page by line ranges of mdsf_splat.cuh / mdsf_fft.cuh (phases)."""
import collections, csv, subprocess, sys, re
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
src = open('md-structure-factor_b200/csrc/mdsf_splat.cuh').read().split('\n')
marks = []
for i, l in enumerate(src, 1):
    for tag, pat in [('A0', '// ---------------- A0'), ('ballot', '// ballot transpose'), ('A1', '// ---------------- A1'),
                     ('B', '// ---------------- B'), ('tail', '    __syncthreads();\n'), ('dump', 'if (dens_dump != nullptr)'),
                     ('store', 'const int lty'), ('setup', 'extern __shared__')]:
        if pat.strip() in l and (tag != 'tail'):
            marks.append((i, tag))
marks.sort()
def phase(fn, ln):
    if fn != 'mdsf_splat.cuh': return fn
    cur = 'pre'
    for i, t in marks:
        if ln >= i: cur = t
    return 'splat:' + cur
agg = collections.Counter(); samp = collections.Counter(); cur = None; hdr = None
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and cur:
        try:
            ln = int(r[0]); ie = float(r[hdr.index('Instructions Executed')] or 0); ns = float(r[hdr.index('# Samples')] or 0)
        except Exception: continue
        agg[phase(cur, ln)] += ie; samp[phase(cur, ln)] += ns
tot = sum(agg.values()); ts = sum(samp.values())
print('total warp inst %.3g samples %d' % (tot, ts)); print(marks)
for k, v in agg.most_common(20):
    print('%6.2f%% inst %6.2f%% samples  %s' % (100 * v / tot, 100 * samp[k] / ts, k))
