"""Metric table for the kernels of BASELINE.json's other configurations from `ncu --page raw --csv` exports
(one file per configuration): one row per distinct kernel instantiation, first launch captured.

    python tools/ncu_configs_summary.py out.md cfg3a=raw1.csv cfg4=raw2.csv ...
"""
import csv
import re
import sys

COLS = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'R'), ('dram__bytes_write.sum', 'W'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
        ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma %'),
        ('launch__registers_per_thread', 'regs'), ('launch__block_size', 'threads'), ('launch__grid_size', 'grid'),
        ('launch__shared_mem_per_block_dynamic', 'smem')]


def main():
    out_path = sys.argv[1]
    lines = ['# Round 2: `ncu --set full` counters of the kernels behind the other configurations\n',
             'Captured under gpurun (one B200) with `tools/gpu/g15.sh`: `ncu --set full --clock-control none` over',
             '`python tools/bench_configs.py <cfg>`, a few launches of each kernel after warm-up; one row per distinct',
             'kernel instantiation (first captured launch).  R / W: DRAM bytes read / written by the launch.\n']
    for arg in sys.argv[2:]:
        cfg, path = arg.split('=', 1)
        try:
            rows = list(csv.reader(open(path, newline='')))
        except Exception as e:
            lines.append('## %s\n\nno capture (%s)\n' % (cfg, e))
            continue
        hdr_i = next((i for i, r in enumerate(rows) if 'Kernel Name' in r), None)
        if hdr_i is None:
            lines.append('## %s\n\nno kernels captured\n' % cfg)
            continue
        hdr, units = rows[hdr_i], rows[hdr_i + 1]
        lines.append('## %s\n' % cfg)
        lines.append('| kernel | ' + ' | '.join(c[1] for c in COLS) + ' |')
        lines.append('|---|' + '---|' * len(COLS))
        seen = set()
        for r in rows[hdr_i + 2:]:
            if len(r) != len(hdr):
                continue
            name = r[hdr.index('Kernel Name')]
            short = re.sub(r'\(.*', '', name).replace('void spcsc::', '').replace('void ', '')
            if short in seen:
                continue
            seen.add(short)
            cells = []
            for m, _ in COLS:
                hit = [i for i, h in enumerate(hdr) if h.endswith(m)]
                if not hit:
                    cells.append('')
                    continue
                v, u = r[hit[0]], units[hit[0]]
                try:
                    fv = float(v.replace(',', ''))
                    v = ('%.3g' % fv) if fv < 1000 else ('%.0f' % fv)
                except ValueError:
                    pass
                cells.append('%s %s' % (v, u) if u and u not in ('%',) else v)
            lines.append('| `%s` | ' % short + ' | '.join(cells) + ' |')
        lines.append('')
    open(out_path, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
