#!/usr/bin/env python
"""Turn gpurun_out/{prof_r01_c2.ncu-rep, launches.csv, bench_default.json} into the committed summaries
under profiles/ (run here, no GPU needed)."""
import collections, csv, json, shutil, subprocess
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] / 1e3      # TB/s, driver-written for this pod
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_r01_c2.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
def num(s):
    try: return float(s.replace(',', ''))
    except ValueError: return 0.0
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_shared_atom.sum', 'lts__t_sector_hit_rate.pct']
stalls = [n for n in h if 'issue_stalled' in n and n.endswith('.ratio')]
bench = json.loads(open('gpurun_out/bench_default.json').readline())
F = bench['config']['frames_per_step']
out, summ = [], {}
for r in rows[2:]:
    k = r[h.index('Kernel Name')].split('(')[0]
    out.append('## ' + k)
    d = {}
    for w in want:
        if w in h:
            out.append('- %s = %s %s' % (w, r[h.index(w)], units[h.index(w)])); d[w] = num(r[h.index(w)]); d[w + '_unit'] = units[h.index(w)]
    st = sorted([(n, num(r[h.index(n)])) for n in stalls], key=lambda x: -x[1])
    out.append('- top stalls (warps per issue): ' + ', '.join('%s %.2f' % (n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for n, v in st[:5]))
    summ[k] = d
def tobytes(v, u): return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
tr, lines = {}, []
for k, d in summ.items():
    b = tobytes(d['dram__bytes_read.sum'], d['dram__bytes_read.sum_unit']) + tobytes(d['dram__bytes_write.sum'], d['dram__bytes_write.sum_unit'])
    ms = d['gpu__time_duration.sum'] * {'ms': 1, 'us': 1e-3, 's': 1e3}.get(d['gpu__time_duration.sum_unit'], 1)
    tr[k] = {'dram_bytes_per_launch': b, 'frames_per_launch': F}
    lines.append('| %s | %.2f | %.3f | %.2f | %.0f %% |' % (k, b / 1e9, ms, b / 1e9 / ms, 100 * b / 1e9 / ms / PEAK))
head = ('# Round 1 - ncu --set full, c2 (256^3, 105456 atoms), %d frames per launch, %s splat mode\n\n'
        'Command: `ncu --set full --clock-control none --import-source on -k regex:"splat_zfft|fft_y|fft_x" -s 3 -c 3 python bench.py --steps 2 --warmup 1 --no-cpu`\n'
        '(the .ncu-rep itself is not committed: 40 MB).\n\n| kernel | DRAM GB / launch | ms (under ncu) | TB/s | of measured %.2f TB/s |\n|---|---|---|---|---|\n' % (F, bench['config']['splat'], PEAK)
        + '\n'.join(lines) + '\n\nThe y and x passes are HBM-bound; the fused splat + z pass writes the pair volumes once and is bound by instruction\n'
        'issue / dependent-load latency of its per-tile phases, not by HBM.\n\n')
open('profiles/r01_ncu_c2_kernels.md', 'w').write(head + '\n'.join(out) + '\n')
json.dump({'workload': 'c2', 'source': 'profiles/r01_ncu_c2_kernels.md', 'kernels': tr}, open('profiles/r01_traffic_c2.json', 'w'), indent=1)
rows = list(csv.reader(open('gpurun_out/launches.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size'); bi = h.index('Block Size')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    d = agg.setdefault(r[ki].split('(')[0][:70], [0, 0.0, r[gi], r[bi]]); d[0] += 1; d[1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in agg.values())
st = bench['stage_ms_per_step']
with open('profiles/r01_launches_c2.md', 'w') as f:
    f.write('# Round 1 - ncu launch list, c2 (256^3, 105456 atoms), %d frames per step, %s splat mode\n\n' % (F, bench['config']['splat']))
    f.write('Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv python bench.py --steps 3 --warmup 2 --no-cpu`.\n')
    main3 = sum(st[k] for k in ('splat_zfft', 'fft_y', 'fft_x_accum'))
    ncu3 = {k: sum(v[1] for n, v in agg.items() if k in n) for k in ('splat_zfft', 'fft_y', 'fft_x_accum')}
    f.write('Times are cold-cache and serialised under the profiler: compare SHARES with the CUDA-event stage times of the un-profiled\n'
            'bench (profiles/r01_bench_c2.json). Among the three compute-stream kernels: splat_zfft %.0f %% (ncu %.0f %%), fft_y %.0f %% (ncu %.0f %%), '
            'fft_x_accum %.0f %% (ncu %.0f %%).\nprep+bin (prep_atoms incl. list-length counting, scan, bin_pairs) runs on its own stream underneath the '
            'previous batch\'s y/x passes; its event span (%.2f ms) includes that waiting, its serialised ncu time is %.2f ms per step.\n\n'
            % (100 * st['splat_zfft'] / main3, 100 * ncu3['splat_zfft'] / sum(ncu3.values()), 100 * st['fft_y'] / main3, 100 * ncu3['fft_y'] / sum(ncu3.values()),
               100 * st['fft_x_accum'] / main3, 100 * ncu3['fft_x_accum'] / sum(ncu3.values()), st['prep_bin'],
               (tot - sum(ncu3.values())) / 1e6 / max(1, sum(v[0] for n, v in agg.items() if 'splat_zfft' in n))))
    f.write('| kernel | launches | total us | share | grid | block |\n|---|---|---|---|---|---|\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('| %s | %d | %.1f | %.1f%% | %s | %s |\n' % (k, v[0], v[1] / 1e3, 100 * v[1] / tot, v[2], v[3]))
shutil.copy('gpurun_out/launches.csv', 'profiles/r01_launches_c2.csv')
shutil.copy('gpurun_out/bench_default.json', 'profiles/r01_bench_c2.json')
print(head)
print(open('profiles/r01_launches_c2.md').read()[:1400])
