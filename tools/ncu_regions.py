#!/usr/bin/env python
"""Per-region instruction / stall-sample shares of the splat kernel from an ncu source-page CSV
(`ncu -i rep --page source --csv --print-source cuda,sass`).  usage: ncu_regions.py csv [n_top_lines]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur = hdr = mode = None
agg, samp, wf, txt = collections.Counter(), collections.Counter(), collections.Counter(), {}
def f(x):
    try: return float(x)
    except ValueError: return 0.0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == 'Line No': hdr = r; mode = 'cuda'; continue
    if r[0] in ('Address', '#'): hdr = r; mode = 'sass'; continue
    if hdr and mode == 'cuda' and cur:
        try: ln = int(r[0])
        except ValueError: continue
        k = (cur, ln)
        agg[k] += f(r[hdr.index('Instructions Executed')]); samp[k] += f(r[hdr.index('# Samples')])
        if 'L1 Wavefronts Shared' in hdr: wf[k] += f(r[hdr.index('L1 Wavefronts Shared')])
        txt[k] = r[1][:80]
tot, ts, tw = sum(agg.values()), sum(samp.values()), max(sum(wf.values()), 1)
print('total warp inst %.3g  samples %d  smem wavefronts %.3g' % (tot, ts, tw))
S = 'mdsf_splat.cuh'
src = open('md-structure-factor_b200/csrc/mdsf_splat.cuh').read().split('\n')
def ln(marker, start=0):
    for i, l in enumerate(src):
        if i >= start and marker in l: return i + 1
    return -1
marks = [('zstage + helpers', 1), ('setup', ln('splat_zfft_kernel(const uint4')), ('item / list head', ln('for (int item = warp; item < nitems;)')),
         ('batch head, produce()', ln('for (int b = 0; b < nmax; b += RB)')), ('record loop', ln('for (int r_i = 0; r_i < mmax; ++r_i)')),
         ('convert', ln('// fixed point -> fp64 (overflow')), ('twiddle load / dump', ln('if (ovf) atomicExch(err_flag, 2);')),
         ('fft dispatch', ln('if (FUSE) {', ln('if (ovf) atomicExch(err_flag, 2);'))), ('store', ln('// store: every tile row x')), ('end', 10 ** 9)]
for (name, a), (_, b) in zip(marks, marks[1:]):
    p = lambda k: k[0] == S and a <= k[1] < b
    print('%-24s lines %4d-%-4d inst %5.1f%%  samples %5.1f%%  smem wf %5.1f%%' % (name, a, min(b, len(src)), 100 * sum(v for k, v in agg.items() if p(k)) / tot,
          100 * sum(v for k, v in samp.items() if p(k)) / ts, 100 * sum(v for k, v in wf.items() if p(k)) / tw))
for fn in sorted(set(k[0] for k in agg) - {S}):
    p = lambda k: k[0] == fn
    print('%-24s %16s inst %5.1f%%  samples %5.1f%%  smem wf %5.1f%%' % (fn[:24], '', 100 * sum(v for k, v in agg.items() if p(k)) / tot,
          100 * sum(v for k, v in samp.items() if p(k)) / ts, 100 * sum(v for k, v in wf.items() if p(k)) / tw))
print()
for k, v in agg.most_common(ntop):
    print('%5.2f%% inst %5.2f%% samp %5.2f%% wf  %s:%d %s' % (100 * v / tot, 100 * samp[k] / ts, 100 * wf[k] / tw, k[0], k[1], txt[k]))
