"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total time, count and share."""
import collections
import csv
import sys


def main(path, top=30):
  lines = [l for l in open(path) if not l.startswith('==')]
  tot = collections.defaultdict(float)
  cnt = collections.Counter()
  for x in csv.DictReader(lines):
    name = x['Kernel Name']
    v = float(x['Metric Value'].replace(',', ''))
    unit = x.get('Metric Unit', 'ns')
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1e-3)
    tot[name] += v
    cnt[name] += 1
  total = sum(tot.values())
  print('total %.1f us over %d launches' % (total, sum(cnt.values())))
  for name, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
    print('%9.1f us %5.1f%% %5d x  %s' % (v, 100 * v / total, cnt[name], name[:110]))


if __name__ == '__main__':
  main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
