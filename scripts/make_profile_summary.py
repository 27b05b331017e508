"""Builds profiles/<tag>_summary.md and profiles/traffic.json from the ncu artefacts brought back in gpurun_out/ (run here, no
GPU needed): the launch list (launches_<tag>.csv), the raw metric pages of the `ncu --set full` captures
(prof_<tag>_<name>_raw.csv, extracted on the GPU box by scripts/ncu_step.sh) and the per-launch CUDA-event table.

The tensor-pipe figure quoted is sm__ops_path_tensor_op_utchmma_*.pct_of_peak_sustained_elapsed: it counts the FLOPs that went
through the tcgen05 (UTCHMMA) path against 8192 FLOP/cycle/SM.  (sm__pipe_tensor_cycles_active_realtime, the counter round 1
quoted, tracks the legacy HMMA sub-pipe and under-reports UTCHMMA work several-fold; the two are printed side by side.)"""
import collections, csv, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
commit = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
out = ['# ncu summary {} (B200, CIFAR-10 32x32, B=256, fp16x3 fwd / bf16x3 grad; commands in scripts/ncu_step.sh; tree at {})\n'.format(tag, commit)]
lp = 'gpurun_out/launches_{}.csv'.format(tag)
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]; kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    data = [(r[kn], float(r[mv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > mv]
    # a step of scripts/profile_step.py begins with the two input conversions (codes, images): the last step is everything from
    # the second-to-last nchw_to_nhwc launch on
    marks = [i for i, (k, _) in enumerate(data) if 'nchw_to_nhwc' in k]
    start = marks[-2] if len(marks) >= 2 else len(data) - len(data) // 4
    last = data[start:]
    per = len(last)
    agg = collections.OrderedDict()
    for k, v in last:
        a = agg.setdefault(k.split('(')[0].replace('void ', '')[:64], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out.append('## Launch list of one step (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n')
    out.append('{} launches, {:.2f} ms summed\n'.format(per, tot / 1e6))
    out.append('| kernel | launches | sum us | share |\n|---|---|---|---|')
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append('| `{}` | {} | {:.1f} | {:.1f}% |'.format(k, c, v / 1e3, 100 * v / tot))
    out.append('')
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'sm__cycles_elapsed.avg.per_second',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed']
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
traffic = None
for name in ['conv_fwd', 'conv_dgrad', 'conv_n128', 'conv_dgrad_n64', 'wgrad', 'wgrad_pair', 'refresh', 'mmd']:
    path = 'gpurun_out/prof_{}_{}_raw.csv'.format(tag, name)
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    di = hdr.index('gpu__time_duration.sum')
    r = max(rows[2:], key=lambda x: float(x[di].replace(',', '')) if len(x) > di and x[di] else 0.0)      # the longest of the captured launches
    d = {h: (u, v) for h, u, v in zip(hdr, units, r)}
    out.append('## `ncu --set full` : {}\n\n```\nKERNEL {}  grid {}'.format(name, d['Kernel Name'][1][:90], d.get('Grid Size', ('', ''))[1]))
    for k in WANT:
        if k in d:
            out.append('   {:<100s} {:<10s} {}'.format(k, d[k][0], d[k][1]))
    for k in sorted(d):
        if k.startswith('sm__ops_path_tensor_op_utchmma') and d[k][1] not in ('0', '', 'n/a') and (k.endswith('.sum') or k.endswith('.avg.pct_of_peak_sustained_elapsed')):
            out.append('   {:<100s} {:<10s} {}'.format(k, d[k][0], d[k][1]))
    out.append('```\n')
    if name == 'conv_fwd':
        f = lambda key: float(d[key][1].replace(',', '')) * SCALE.get(d[key][0], 1.0)
        traffic = {'workload': 'cifar', 'batch': 256, 'commit': commit, 'capture': os.path.basename(path), 'kernel': d['Kernel Name'][1],
                   'launch': 'D forward conv 256->256 3x3 @8x8 on 512 images (M=32768, N=256, K=2304): algorithmic operand bytes 36.2 MB',
                   'dram_bytes_read': f('dram__bytes_read.sum'), 'dram_bytes_write': f('dram__bytes_write.sum'),
                   'duration_us_under_ncu': float(d['gpu__time_duration.sum'][1]),
                   'l2_to_sm_bytes': float(d['lts__t_sectors_srcunit_tex_op_read.sum'][1]) * 32.0,
                   'lts_hit_rate_pct': float(d['lts__t_sector_hit_rate.pct'][1])}
        traffic['dram_bytes'] = traffic['dram_bytes_read'] + traffic['dram_bytes_write']
        for k in d:
            if k.startswith('sm__ops_path_tensor_op_utchmma') and k.endswith('fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed') and float(d[k][1] or 0) > 0:
                traffic['tensor_pipe_pct_utchmma'] = float(d[k][1])
for extra in ['gpurun_out/events_{}.txt'.format(tag)]:
    if os.path.exists(extra):
        out.append('## {}\n\n```\n{}```\n'.format(os.path.basename(extra), open(extra).read()))
os.makedirs('profiles', exist_ok=True)
open('profiles/{}_summary.md'.format(tag), 'w').write('\n'.join(out))
if traffic:
    json.dump(traffic, open('profiles/traffic.json', 'w'), indent=1)
    print(json.dumps(traffic, indent=1))
print('wrote profiles/{}_summary.md'.format(tag), len(out))
