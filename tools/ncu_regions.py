#!/usr/bin/env python
"""Per-region instruction / stall-sample shares of the splat kernel (z-lane variant) from an ncu source-page CSV
(`ncu -i rep --page source --csv --print-source cuda,sass`, see tools/gpu_src.sh), plus the top stall reasons by source
line.  usage: ncu_regions.py source.csv [raw.csv] > profiles/r02_splat_regions_c3.md"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
def f(x):
    try: return float(x)
    except ValueError: return 0.0
cur = hdr = mode = None
agg, samp, txt = collections.Counter(), collections.Counter(), {}
stall, tot = collections.defaultdict(collections.Counter), collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == 'Line No': hdr = r; mode = 'cuda'; continue
    if r[0] in ('Address', '#'): hdr = r; mode = 'sass'; continue
    if hdr and mode == 'cuda' and cur:
        try: ln = int(r[0])
        except ValueError: continue
        k = (cur, ln); txt[k] = r[1][:90]
        agg[k] += f(r[hdr.index('Instructions Executed')]); samp[k] += f(r[hdr.index('# Samples')])
        for i, hn in enumerate(hdr):
            if hn.startswith('stall_') and 'Not Issued' not in hn:
                stall[hn][k] += f(r[i]); tot[hn] += f(r[i])
T, S = sum(agg.values()), sum(samp.values())
src = open('md-structure-factor_b200/csrc/mdsf_splat.cuh').read().split('\n')
def ln(m, start=0):
    for i, l in enumerate(src):
        if i >= start and m in l: return i + 1
    return -1
end = ln('if (ovf) atomicExch(err_flag, 2);')
marks = [('helpers, z stages (zstage2)', 1), ('kernel set-up', ln('splat_zfft_kernel(const uint4')), ('z-lane: list bounds, item order', ln('if constexpr (ZL &&')),
         ('z-lane: batch head + products', ln('const int m = min(RBZ, n - b);')), ('z-lane: record loop', ln('auto accumulate = [&]')),
         ('z-lane: convert', ln('// fixed point -> fp64 (overflow: a cell held > 2048 peak amplitudes)')), ('half-warp / general / density loop', ln('// work items: (group of SUB consecutive slabs')),
         ('twiddle late load, density tap', end), ('z FFT dispatch', ln('if (FUSE) {', end)), ('tile store', ln('// store: every tile row x')), ('end', 10 ** 9)]
print('# splat_zfft_kernel (c3, 8 frames), ncu source view: %.3g warp instructions, %d stall samples\n' % (T, S))
print('| region of mdsf_splat.cuh | lines | instructions | stall samples |\n|---|---|---|---|')
for (name, a), (_, b) in zip(marks, marks[1:]):
    sel = [k for k in agg if k[0] == 'mdsf_splat.cuh' and a <= k[1] < b]
    print('| %s | %d-%d | %.1f %% | %.1f %% |' % (name, a, min(b, len(src)), 100 * sum(agg[k] for k in sel) / T, 100 * sum(samp[k] for k in sel) / S))
for fn in sorted(set(k[0] for k in agg) - {'mdsf_splat.cuh'}):
    sel = [k for k in agg if k[0] == fn]
    print('| %s | | %.1f %% | %.1f %% |' % (fn, 100 * sum(agg[k] for k in sel) / T, 100 * sum(samp[k] for k in sel) / S))
TT = sum(tot.values())
print('\n## stall reasons (share of all samples) and their top source lines\n')
for hn, v in tot.most_common(6):
    print('* **%s %.1f %%**' % (hn.replace('stall_', ''), 100 * v / TT))
    for k, x in stall[hn].most_common(3): print('  * %.0f %% at `%s:%d` `%s`' % (100 * x / v, k[0], k[1], txt[k].strip()[:70]))
if len(sys.argv) > 2:
    raw = list(csv.reader(open(sys.argv[2])))
    h, v = raw[0], raw[2]
    print('\n## kernel metrics\n')
    for w in ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.per_cycle_active', 'launch__registers_per_thread',
              'launch__shared_mem_per_block_dynamic', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
              'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):
        if w in h: print('- %s = %s %s' % (w, v[h.index(w)], raw[1][h.index(w)]))
