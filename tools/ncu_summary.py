"""CPU-only: key metrics of every kernel in an ncu report, as a markdown table.  usage: python tools/ncu_summary.py rep [rep ...]"""
import csv
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'us', 1e-3), ('dram__bytes_read.sum', 'rd MB', 1e-6), ('dram__bytes_write.sum', 'wr MB', 1e-6),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram %', 1), ('lts__t_sectors_srcunit_tex.sum', 'L2<->SM MB', 32e-6),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %', 1), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts %', 1),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor %', 1),
        ('sm__inst_executed_pipe_tensor.sum', 'tensor inst', 1), ('sm__cycles_active.avg', 'SM cycles', 1),
        ('launch__registers_per_thread', 'regs', 1), ('launch__grid_size', 'grid', 1),
        ('smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'long_sb', 1),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm %', 1),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem conflicts', 1)]


def rows_of(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    res = []
    for r in rows[2:]:
        d = {'name': r[hdr.index('Kernel Name')][:60]}
        for m, label, k in WANT:
            if m in hdr:
                try:
                    d[label] = float(r[hdr.index(m)].replace(',', '')) * k
                except ValueError:
                    d[label] = None
        res.append(d)
    return res


if __name__ == '__main__':
    for path in sys.argv[1:]:
        rows = rows_of(path)
        labels = [l for _, l, _ in WANT if any(l in r for r in rows)]
        print(f'### {path}\n')
        print('| # | kernel | ' + ' | '.join(labels) + ' |')
        print('|---|---|' + '---|' * len(labels))
        for i, r in enumerate(rows):
            print(f'| {i} | {r["name"]} | ' + ' | '.join('' if r.get(l) is None else (f'{r[l]:.1f}' if r[l] < 1e6 else f'{r[l]:.3g}') for l in labels) + ' |')
        print()
