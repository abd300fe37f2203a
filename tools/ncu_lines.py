#!/usr/bin/env python
"""Per source line: stall samples, instructions, shared-memory wavefronts (total / excessive). usage: ncu_lines.py rep kernel [n]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
exc = collections.Counter(); wf = collections.Counter(); ie = collections.Counter(); smp = collections.Counter(); txt = {}
cur = None; hdr = None
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and cur:
        try: ln = int(r[0])
        except ValueError: continue
        def g(nm):
            try: return float(r[hdr.index(nm)] or 0)
            except ValueError: return 0.0
        k = (cur, ln)
        exc[k] += g('L1 Wavefronts Shared Excessive'); wf[k] += g('L1 Wavefronts Shared'); ie[k] += g('Instructions Executed'); smp[k] += g('# Samples'); txt[k] = r[1][:90]
tot = sum(wf.values()); te = sum(exc.values()); ts = sum(smp.values()); ti = sum(ie.values())
print('warp inst %.3g  samples %d  smem wavefronts %.3g (excessive %.3g)' % (ti, ts, tot, te))
for k, v in sorted(smp.items(), key=lambda kv: -kv[1])[:n]:
    print('%5.1f%% samp %5.1f%% inst  wf %5.1f%% exc %5.1f%%  %s:%d %s' % (100 * v / ts, 100 * ie[k] / ti, 100 * wf[k] / max(tot, 1), 100 * exc[k] / max(te, 1), k[0], k[1], txt[k]))
