"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (shares of the step; per-launch times under ncu
are cold-cache and serialised -- compare SHARES with bench.py's CUDA-event breakdown, not absolutes)."""
import csv
import sys
from collections import defaultdict


def main(path, top=45):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    tot = defaultdict(float); cnt = defaultdict(int)
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(',', ''))
        u = r[iu]
        us = v / 1000 if u in ('nsecond', 'ns') else (v * 1000 if u in ('msecond', 'ms') else v)
        name = r[ik].split('(')[0][-70:]
        tot[name] += us; cnt[name] += 1
    total = sum(tot.values())
    print('%12s %7s %7s  kernel   (total %.1f ms over %d launches)' % ('total_us', 'share', 'count', total / 1000, sum(cnt.values())))
    for name, us in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print('%12.1f %6.2f%% %7d  %s' % (us, 100 * us / total, cnt[name], name))


if __name__ == '__main__':
    main(sys.argv[1])
