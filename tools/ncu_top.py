#!/usr/bin/env python
"""Top source lines by executed warp instructions / stall samples from an ncu report (source page).
usage: ncu_top.py report.ncu-rep kernel_regex [n]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
agg = collections.Counter(); samp = collections.Counter(); cur = None; hdr = None; seen_kernel = 0
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and cur:
        try:
            ln = int(r[0]); ie = float(r[hdr.index('Instructions Executed')] or 0); ns = float(r[hdr.index('# Samples')] or 0)
        except Exception: continue
        agg[(cur, ln, r[1][:100])] += ie; samp[(cur, ln, r[1][:100])] += ns
tot = sum(agg.values()); ts = sum(samp.values())
print('total warp inst %.3g, samples %d' % (tot, ts))
for k, v in agg.most_common(n):
    print('%6.2f%% inst  %5.1f%% samp  %s:%d  %s' % (100 * v / tot, 100 * samp[k] / max(ts, 1), k[0], k[1], k[2]))
