"""Summarise an ncu report here (no GPU): key raw metrics + the SASS lines with the most stall samples."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__cycles_elapsed.max', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print('KERNEL', r[hdr.index('Kernel Name')][:70], 'grid', r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
    for h, u, v in zip(hdr, units, r):
        if any(h.endswith(w) for w in want):
            print('   %-95s %-10s %s' % (h, u, v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
his = [i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r]
for k, hi in enumerate(his[:1]):
    hdr = rows[hi]
    si, sa = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
    end = his[k + 1] if k + 1 < len(his) else len(rows)
    data = []
    for r in rows[hi + 1:end]:
        try:
            data.append((float(r[sa]), r[si][:120]))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1
    print('--- top stall lines (of %d samples)' % tot)
    for v, s in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
        print('%6.1f%%  %s' % (100 * v / tot, s))
