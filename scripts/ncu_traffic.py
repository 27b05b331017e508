"""profiles/traffic.json from an `ncu --set full` report (run here, no GPU): the DRAM traffic of ONE launch of the top kernel,
which bench.py puts into `roofline.traffic` next to the commit the capture was taken at.

  python scripts/ncu_traffic.py gpurun_out/prof_<tag>_conv_fwd.ncu-rep cifar 256 [commit]
"""
import csv, json, os, subprocess, sys
rep, workload, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
commit = sys.argv[4] if len(sys.argv) > 4 else subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}


def get(name, default=None):
    for h, u, v in zip(hdr, units, r):
        if h == name:
            return float(v.replace(',', '')) * SCALE.get(u, 1.0)
    return default


out = {'workload': workload, 'batch': batch, 'commit': commit, 'report': os.path.basename(rep),
       'kernel': r[hdr.index('Kernel Name')], 'grid': r[hdr.index('Grid Size')] if 'Grid Size' in hdr else None,
       'dram_bytes_read': get('dram__bytes_read.sum'), 'dram_bytes_write': get('dram__bytes_write.sum'),
       'duration_us_under_ncu': get('gpu__time_duration.sum'),
       'lts_sectors_tex_read': get('lts__t_sectors_srcunit_tex_op_read.sum'),
       'lts_hit_rate_pct': get('lts__t_sector_hit_rate.pct'),
       'tensor_pipe_pct': get('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed')}
out['dram_bytes'] = (out['dram_bytes_read'] or 0.0) + (out['dram_bytes_write'] or 0.0)
if out['lts_sectors_tex_read']:
    out['l2_to_sm_bytes'] = out['lts_sectors_tex_read'] * 32.0
os.makedirs('profiles', exist_ok=True)
json.dump(out, open('profiles/traffic.json', 'w'), indent=1)
print(json.dumps(out, indent=1))
