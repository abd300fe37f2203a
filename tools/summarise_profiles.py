#!/usr/bin/env python
"""Turn gpurun_out/{prof_r02_<wl>.ncu-rep, launches_<wl>.csv} into the committed summaries under profiles/
(run here, no GPU needed):  profiles/r02_ncu_<wl>_kernels.md, r02_traffic_<wl>.json, r02_launches_<wl>.csv / .md.
usage: summarise_profiles.py <workload> <frames_per_launch>"""
import collections, csv, json, shutil, subprocess, sys
wl, F = sys.argv[1], int(sys.argv[2])
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] / 1e3      # TB/s, driver-written for this pod
rows = list(csv.reader(open("gpurun_out/prof_r02_%s_raw.csv" % wl)))     # `ncu -i prof.ncu-rep --page raw --csv`, exported on the GPU box
h, units = rows[0], rows[1]
def num(s):
    try: return float(s.replace(',', ''))
    except ValueError: return 0.0
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
stalls = [n for n in h if 'issue_stalled' in n and n.endswith('.ratio')]
short = {'prep_atoms': 'prep_bin', 'bin_place': 'prep_bin', 'bin_count': 'prep_bin', 'order_atoms': 'prep_bin', 'tma_pass_kernel<0>': 'fft_y', 'tma_pass_kernel<1>': 'fft_x_accum', 'DeviceScan': 'prep_bin', 'splat_zfft': 'splat_zfft', 'fft3_pass_kernel<8, 8, 8, 3, 256, 2, 0>': 'fft_y',
         'fft3_pass_kernel<8, 8, 8, 3, 256, 2, 1>': 'fft_x_accum', 'fft_y': 'fft_y', 'fft_x': 'fft_x_accum', 'yx_pass': 'fft_yx'}
out, traffic = ['# ncu --set full, %s, one launch per kernel (%d frames per launch); peak = %.2f TB/s measured' % (wl, F, PEAK), ''], collections.OrderedDict()
seen = set()
for r in rows[2:]:
    k = r[h.index('Kernel Name')]
    base = k.split('(')[0]
    if base in seen: continue
    seen.add(base)
    out.append('## ' + base)
    d = {}
    for w in want:
        if w in h:
            out.append('- %s = %s %s' % (w, r[h.index(w)], units[h.index(w)])); d[w] = num(r[h.index(w)])
    st = sorted([(n, num(r[h.index(n)])) for n in stalls], key=lambda x: -x[1])
    out.append('- top stalls (warps per issue): ' + ', '.join('%s %.2f' % (n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for n, v in st[:6]))
    def tobytes(name):
        v, u = d.get(name, 0.0), units[h.index(name)] if name in h else ''
        return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}.get(u, 1.0)
    by = tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum')
    ms = d.get('gpu__time_duration.sum', 0.0) * {'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}.get(units[h.index('gpu__time_duration.sum')], 1.0)
    out.append('- DRAM traffic %.3f GB in %.3f ms = %.2f TB/s = %.1f%% of the measured peak' % (by / 1e9, ms, by / 1e9 / max(ms, 1e-9), 100 * by / 1e9 / max(ms, 1e-9) / PEAK))
    out.append('')
    stage = next((v for kk, v in short.items() if kk in k), base)
    t = traffic.setdefault(stage, {'dram_bytes_per_launch': 0.0, 'frames_per_launch': F, 'kernels': []})
    t['dram_bytes_per_launch'] += by
    t['kernels'].append(base)
open('profiles/r02_ncu_%s_kernels.md' % wl, 'w').write('\n'.join(out) + '\n')
json.dump({'workload': wl, 'source': 'ncu --set full --clock-control none, gpurun_out/prof_r02_%s_raw.csv (ncu --page raw --csv of the capture)' % wl, 'kernels': traffic}, open('profiles/r02_traffic_%s.json' % wl, 'w'), indent=1)
# launch list
src = 'gpurun_out/launches_%s.csv' % wl
lr = list(csv.reader(open(src)))
for i, r in enumerate(lr):
    if r and r[0] == 'ID': hdr = r; start = i + 1; break
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
tot = collections.OrderedDict()
with open('profiles/r02_launches_%s.csv' % wl, 'w') as fh:
    fh.write('id,kernel,ns\n')
    for r in lr[start:]:
        if len(r) > vi:
            name = r[ki].split('(')[0]
            fh.write('%s,"%s",%s\n' % (r[0], name, r[vi]))
            tot[name] = tot.get(name, 0.0) + num(r[vi])
s = sum(tot.values())
with open('profiles/r02_launches_%s.md' % wl, 'w') as fh:
    fh.write('# ncu launch list (gpu__time_duration.sum, cold-cache, serialised), %s: share of the step per kernel\n\n' % wl)
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        fh.write('- %5.1f%%  %8.3f ms  %s\n' % (100 * v / s, v / 1e6, k))
print(open('profiles/r02_launches_%s.md' % wl).read())
