"""profiles/<tag>_launch_summary.md and <tag>_ncu_full_summary.md from the outputs of tools/final_profile.sh <tag>
(gpurun_out/<tag>_launches.csv, <tag>_raw.csv, <tag>_bench.json).  usage: python tools/profile_summaries.py <tag>"""
import collections, csv, json, os, sys
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
bench = json.loads(open(os.path.join(G, tag + '_bench.json')).read().strip().splitlines()[-1])

# ---- launch list
with open(os.path.join(G, tag + '_launches.csv')) as f:
    lines = [l for l in f if l.startswith('"')]
r = list(csv.reader(lines)); hdr = r[0]; data = r[1:]
ci = {h: i for i, h in enumerate(hdr)}
names = [d[ci['Kernel Name']].split('(')[0].replace('pkb::', '') for d in data]
vals = [float(d[ci['Metric Value']]) / 1000.0 for d in data]
idx = [i for i, n in enumerate(names) if n == 'k_bvn_setup']
seg = list(zip(names[idx[0]:idx[1]], vals[idx[0]:idx[1]]))
tot = sum(v for _, v in seg)
agg = collections.OrderedDict()
for n, v in seg:
    c = agg.setdefault(n, [0, 0.0]); c[0] += 1; c[1] += v
fam = bench['roofline'].get('family_ms_per_solve', {})
km = bench['roofline_chain'].get('kernel_ms', {})


def live(n):
    if n in fam:
        return '%.3f' % fam[n]
    if n == 'k_kernel_rows':
        return '%.3f' % km.get('k_kernel_rows_win', 0)
    return '%.3f' % km[n] if n in km else ''


out = ['# %s: ncu launch list of one C4 solve\n' % tag,
       '`ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 900 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras`\n'
       '(`tools/final_profile.sh %s`; raw list: `profiles/%s_launches.csv`).  The table is the solve between two `k_bvn_setup` launches:\n'
       '%d launches, %.2f ms of kernel time (serialised, cold cache -- ncu replays every launch alone; compare SHARES with the\n'
       'live per-kernel CUDA-event figures of the same bench command, right column, which overlap the side-stream kernels with the chain).\n'
       % (tag, tag, len(seg), tot / 1000),
       '| kernel | launches | us total (ncu) | share | ms per solve, CUDA events in `bench.py` (`profiles/%s_bench.json`) |\n|---|---|---|---|---|' % tag]
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append('| `%s` | %d | %.1f | %.1f %% | %s |' % (n, c, v, 100 * v / tot, live(n)))
fft = sum(v for n, (c, v) in agg.items() if n in ('k_rows_fwd', 'k_cols', 'k_rows_inv'))
out.append('\nThe three FFT kernels take %.0f %% of the kernel time under ncu; `k_rows_fwd/k_cols/k_rows_inv` launches on support windows (54 of the 59 '
           'steps) and on the whole torus are the same kernels (profiled separately as `k_*_win` by the live events).  The bench line of the same build: '
           '%.1f days/s device-resident (%.2f ms), %.1f days/s end to end (%.2f ms), `roofline_chain.frac` %.3f.'
           % (100 * fft / tot, bench['value'], bench['ms_per_step'], bench['e2e']['value'], bench['e2e']['ms_per_step'], bench['roofline_chain']['frac']))
open(os.path.join(P, tag + '_launch_summary.md'), 'w').write('\n'.join(out) + '\n')

# ---- full capture
rows = list(csv.reader(open(os.path.join(G, tag + '_raw.csv'))))
hdr = rows[0]; data = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}


def col(k, fmt='%.1f', scale=1.0):
    o = []
    for rr in data:
        try:
            o.append(fmt % (float(rr[ci[k]]) * scale))
        except ValueError:
            o.append(rr[ci[k]])
    return o


names = [rr[ci['Kernel Name']].split('(')[0] for rr in data]
tab = [('duration [us]', 'gpu__time_duration.sum', '%.1f', 1), ('DRAM read [MB]', 'dram__bytes_read.sum', '%.1f', 1), ('DRAM write [MB]', 'dram__bytes_write.sum', '%.1f', 1),
       ('fp64 pipe active [%]', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', '%.1f', 1),
       ('issue slots active [%]', 'smsp__issue_active.avg.pct_of_peak_sustained_active', '%.1f', 1),
       ('shared-memory data pipe [% of peak]', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', '%.1f', 1),
       ('warps active [% of peak]', 'sm__warps_active.avg.pct_of_peak_sustained_active', '%.1f', 1), ('registers / thread', 'launch__registers_per_thread', '%.0f', 1),
       ('threads / CTA', 'launch__block_size', '%.0f', 1), ('grid', 'launch__grid_size', '%.0f', 1),
       ('dynamic shared memory [KB / CTA]', 'launch__shared_mem_per_block_dynamic', '%.1f', 1),
       ('occupancy limit smem [CTAs/SM]', 'launch__occupancy_limit_shared_mem', '%.0f', 1), ('occupancy limit regs [CTAs/SM]', 'launch__occupancy_limit_registers', '%.0f', 1),
       ('warp instructions [M]', 'smsp__inst_executed.sum', '%.1f', 1e-6), ('shared wavefronts [M]', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', '%.1f', 1e-6),
       ('shared bank conflicts [M]', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', '%.1f', 1e-6), ('L2 hit rate [%]', 'lts__t_sector_hit_rate.pct', '%.1f', 1)]
out = ['# %s: `ncu --set full` of the three FFT kernels of one chain step of the C4 solve\n' % tag,
       'Source: gpurun_out/%s_prof.ncu-rep (`ncu --set full --clock-control none --import-source on -k regex:^k_cols|^k_rows_fwd|^k_rows_inv -s 270 -c 3\n'
       'python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras`, `tools/final_profile.sh %s`, run right after the same command exited 0\n'
       'without ncu), read with `ncu -i ... --page raw --csv`.  The captured step is a SUPPORT-WINDOW step in the middle of the chain (the kind of step 54 of\n'
       'the 59 days of C4 run as); same capture point as the earlier `*_ncu_full_summary.md` files, so the tables compare build against build.\n' % (tag, tag),
       '| metric | ' + ' | '.join(names) + ' |', '|---|' + '---|' * len(names)]
for label, k, fmt, sc in tab:
    out.append('| %s | %s |' % (label, ' | '.join(col(k, fmt, sc))))
stalls = [h for h in hdr if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
for j, n in enumerate(names):
    v = sorted(((float(data[j][ci[h]]), h.split('stalled_')[1]) for h in stalls), reverse=True)[:6]
    out.append('\n`%s` stall reasons (pc sampling, top 6): %s' % (n, ', '.join('%s %d' % (b, a) for a, b in v)))
out.append('\nReading: see `r2r_ncu_full_summary.md` (same kernels; this capture is of the last build of the round).')
open(os.path.join(P, tag + '_ncu_full_summary.md'), 'w').write('\n'.join(out) + '\n')
print('wrote', tag)
