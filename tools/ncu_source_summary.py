"""Summarise `ncu --page source --csv` output (SASS view): per kernel the stall-sample totals by reason, by
opcode class, and the hottest instructions.   python tools/ncu_source_summary.py file.csv [top]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    kernels = []
    cur = None
    with open(path, newline='') as f:
        for row in csv.reader(f):
            if not row:
                continue
            if row[0] == 'Kernel Name':
                cur = {'name': row[1], 'hdr': None, 'rows': []}
                kernels.append(cur)
            elif row[0] == 'Address':
                cur['hdr'] = row
            elif cur is not None and cur['hdr'] is not None:
                cur['rows'].append(row)
    for k in kernels:
        h = k['hdr']
        ix = {n: i for i, n in enumerate(h)}
        reasons = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
        tot = collections.Counter()
        byop = collections.Counter()
        byop_reason = collections.defaultdict(collections.Counter)
        inst = []
        nsamp = 0
        executed = 0
        for r in k['rows']:
            if len(r) < len(h):
                r = r + ['0'] * (len(h) - len(r))
            s = int(r[ix['# Samples']] or 0)
            nsamp += s
            executed += int(r[ix['Instructions Executed']] or 0)
            src = r[ix['Source']].strip()
            tok = src.split()
            op = tok[1] if tok and tok[0].startswith('@') and len(tok) > 1 else (tok[0] if tok else '?')
            opc = op.split('.')[0]
            byop[opc] += s
            for n in reasons:
                v = int(r[ix[n]] or 0)
                tot[n] += v
                byop_reason[opc][n] += v
            inst.append((s, r[ix['Address']][-5:], src, {n: int(r[ix[n]] or 0) for n in reasons}))
        print('=' * 100)
        print(k['name'][:140])
        print('samples %d, warp instructions executed %d' % (nsamp, executed))
        print('by reason: ' + ', '.join('%s %.1f%%' % (n[6:], 100.0 * v / max(nsamp, 1)) for n, v in tot.most_common(10)))
        print('by opcode: ' + ', '.join('%s %.1f%%' % (o, 100.0 * v / max(nsamp, 1)) for o, v in byop.most_common(14)))
        for o, v in byop.most_common(8):
            print('   %-8s %5.1f%%: %s' % (o, 100.0 * v / max(nsamp, 1), ', '.join(
                '%s %d' % (n[6:], c) for n, c in byop_reason[o].most_common(4))))
        print('hottest instructions:')
        for s, addr, src, rs in sorted(inst, key=lambda t: -t[0])[:top]:
            main_r = sorted(rs.items(), key=lambda kv: -kv[1])[:2]
            print('  %5d (%4.1f%%) %s  %-60s %s' % (s, 100.0 * s / max(nsamp, 1), addr, src[:60],
                                                   ', '.join('%s %d' % (n[6:], c) for n, c in main_r)))


if __name__ == '__main__':
    main()
