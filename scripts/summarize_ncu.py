"""Summarise an ncu launch list (csv from --metrics gpu__time_duration.sum) into per-kernel totals and shares."""
import collections
import csv
import sys


def main(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    kn, mv = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 2:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(',', ''))
        except ValueError:
            continue
        name = r[kn].split('(')[0][:70]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write(f'# source: {path}  (gpu__time_duration.sum, cold-cache serialised launches: compare SHARES)\n')
        f.write(f'# total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches\n')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{v[1] / 1e6:10.3f} ms {v[0]:6d} launches {100 * v[1] / tot:6.2f}%  {k}\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
