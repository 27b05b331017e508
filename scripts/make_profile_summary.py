"""Builds profiles/<tag>_summary.md from the ncu artefacts brought back in gpurun_out/ (run here, no GPU needed)."""
import collections, csv, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
out = ['# ncu summary {} (B200, CIFAR-10 32x32, B=256, fp16x3 fwd / bf16x3 grad; commands in scripts/ncu_step.sh)\n'.format(tag)]
lp = 'gpurun_out/launches_{}.csv'.format(tag)
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]; kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    data = [(r[kn], float(r[mv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > mv]
    per = len(data) // 4
    last = data[-per:]
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
for name in ['conv_fwd', 'conv_dgrad', 'conv_dgrad_n64', 'wgrad', 'mmd']:
    rep = 'gpurun_out/prof_{}_{}.ncu-rep'.format(tag, name)
    if os.path.exists(rep):
        txt = subprocess.run([sys.executable, 'scripts/ncu_top.py', rep, '12'], capture_output=True, text=True).stdout
        out.append('## `ncu --set full` : {}\n\n```\n{}```\n'.format(name, txt))
for extra in ['gpurun_out/events_{}.txt'.format(tag), 'gpurun_out/debug_{}.txt'.format(tag)]:
    if os.path.exists(extra):
        out.append('## {}\n\n```\n{}```\n'.format(os.path.basename(extra), open(extra).read()))
os.makedirs('profiles', exist_ok=True)
open('profiles/{}_summary.md'.format(tag), 'w').write('\n'.join(out))
print('wrote profiles/{}_summary.md'.format(tag), len(out))
