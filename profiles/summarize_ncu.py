#!/usr/bin/env python
"""Turns an `ncu --set full` report into the short text summary committed next to it.

    python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep > profiles/r01_x.txt

Columns: duration, DRAM read / write bytes (=> `roofline.traffic`), achieved DRAM GB/s, L2 / L1 / SM throughput in % of
peak, tensor-pipe activity, registers, occupancy limits, top warp-stall reasons."""
import csv
import subprocess
import sys

KEYS = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2%'), ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM%'),
        ('sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('sm__pipe_tensor_subpipe_tmem_cycles_active.avg.pct_of_peak_sustained_active', 'tcgen05%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('launch__registers_per_thread', 'regs'),
        ('launch__occupancy_limit_registers', 'lim_reg'), ('launch__occupancy_limit_shared_mem', 'lim_smem'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'block')]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    unit = dict(zip(h, units))
    print('# %s' % path)
    for row in rows[2:]:
        d = dict(zip(h, row))
        name = d['Kernel Name'].replace('void tha4::<unnamed>::', '').split('(CUtensorMap')[0].split('(const')[0]
        print('\n== %s' % name[:110])
        vals = []
        for k, label in KEYS:
            v = d.get(k)
            if v in (None, ''):
                continue
            vals.append('%s=%s%s' % (label, v.replace('.000000', ''), (' ' + unit.get(k, '')) if label in ('us', 'dram_rd', 'dram_wr') else ''))
        print('   ' + '  '.join(vals))
        try:
            t_us = float(d['gpu__time_duration.sum']) * {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 's': 1e6}.get(unit['gpu__time_duration.sum'], 1.0)
            def to_bytes(k):
                v, u = float(d[k]), unit[k]
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            tot = to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')
            print('   dram traffic %.2f MB  =>  %.0f GB/s over the launch (%.1f us)' % (tot / 1e6, tot / t_us / 1e3, t_us))
        except Exception:
            pass
        tensor = [(k, d[k]) for k in h if ('tensor' in k or 'tmem' in k) and 'pct' in k and d.get(k) not in (None, '', 'n/a')]
        for k, v in tensor[:6]:
            print('   %s = %s' % (k, v))
        st = {}
        for k in h:
            if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio'):
                try:
                    st[k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]] = float(d[k])
                except ValueError:
                    pass
        print('   top stalls: ' + ', '.join('%s %.1f' % kv for kv in sorted(st.items(), key=lambda x: -x[1])[:5]))


if __name__ == '__main__':
    for pth in sys.argv[1:]:
        main(pth)
