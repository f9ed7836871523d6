"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) into a markdown table of per-kernel shares.

    python tools/launch_summary.py gpurun_out/launches.csv "title" > profiles/rNN_launches.md
"""
import collections
import csv
import sys


def main(path, title):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        name = r[ik].split('(')[0].replace('void ', '')[:70]
        agg.setdefault(name, []).append(float(r[iv].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    print('# Launch list: %s\n' % title)
    print('`ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times are cold-cache and serialised: '
          'compare SHARES, not absolutes).\n')
    print('| kernel | launches | mean ns | share |')
    print('|---|---|---|---|')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print('| `%s` | %d | %.0f | %.1f %% |' % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
