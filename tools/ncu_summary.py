"""Turn an `ncu --page raw --csv` export into profiles/<tag>_ncu_summary.md (a metric table per kernel) and merge
the DRAM traffic per launch into profiles/r02_traffic.json (read by bench.py for roofline.traffic).

    python tools/ncu_summary.py <raw.csv> <tag> "<command that was profiled>" [source.csv]
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    ('gpu__time_duration.sum', 'us'),
    ('dram__bytes_read.sum', ''), ('dram__bytes_write.sum', ''),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', '%'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', '%'),
    ('launch__registers_per_thread', ''), ('launch__block_size', ''), ('launch__grid_size', ''),
    ('launch__cluster_size', ''), ('launch__shared_mem_per_block_dynamic', ''),
    ('launch__occupancy_limit_registers', ''), ('launch__occupancy_limit_shared_mem', ''),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', '%'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', '%'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', '%'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_active', '%'),
    ('l1tex__t_sector_hit_rate.pct', '%'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', ''),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', ''),
    ('smsp__inst_executed.sum', ''),
    ('lts__t_sector_hit_rate.pct', '%'),
]


def to_bytes(val, unit):
    v = float(val.replace(',', ''))
    u = unit.lower()
    return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'tbyte': 1e12}.get(u, 1)


def main():
    raw, tag, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    src = sys.argv[4] if len(sys.argv) > 4 else None
    rows = list(csv.reader(open(raw, newline='')))
    hdr, units = rows[0], rows[1]
    out = ['# Round 2 (%s): `ncu --set full` of the hot kernels\n' % tag,
           'Command (under gpurun, one B200): `%s`  ' % cmd,
           '(the report is exported on the box with `ncu -i ... --page raw --csv` and `--page source --csv`; the `.ncu-rep` with',
           'imported sources is too large to travel).  One launch per kernel, iterations 10+ of the bench solve.\n']
    traffic_path = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    try:
        traffic = json.load(open(traffic_path))
    except Exception:
        traffic = {'kernels': {}}
    traffic['source'] = 'ncu --set full, %s' % tag
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        short = re.sub(r'\(.*', '', name).replace('void spcsc::', '').replace('void ', '')
        key = short.split('<')[0]
        if key in seen:
            continue
        seen.add(key)
        out.append('## %s\n' % short)
        out.append('| metric | value | unit |')
        out.append('|---|---|---|')
        vals = {}
        for i, h in enumerate(hdr):
            base = h.split('.', 2)[-1] if h.count('.') > 2 and h.split('.')[0].isupper() else h
            vals[h] = (r[i], units[i])
            vals[base] = (r[i], units[i])
        for m, _ in WANT:
            hit = [k for k in vals if k.endswith(m)]
            if hit:
                v, u = vals[hit[0]]
                out.append('| %s | %s | %s |' % (m, v, u))
        rd = [k for k in vals if k.endswith('dram__bytes_read.sum')]
        wr = [k for k in vals if k.endswith('dram__bytes_write.sum')]
        if rd and wr:
            tb = to_bytes(*vals[rd[0]]) + to_bytes(*vals[wr[0]])
            traffic['kernels'][key] = {'dram_bytes': tb, 'full_name': short}
            out.append('| dram read + write per launch | %.4f | Gbyte |' % (tb / 1e9))
        out.append('')
    if src:
        out.append('## Warp-stall samples by SASS instruction (`--page source`)\n')
        out.append('```')
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_source_summary.py'), src, '22'],
                           stdout=subprocess.PIPE, text=True)
        out.append(r.stdout.rstrip())
        out.append('```')
    open(os.path.join(ROOT, 'profiles', '%s_ncu_summary.md' % tag), 'w').write('\n'.join(out) + '\n')
    json.dump(traffic, open(traffic_path, 'w'), indent=1)
    print('\n'.join(out[:60]))


if __name__ == '__main__':
    main()
