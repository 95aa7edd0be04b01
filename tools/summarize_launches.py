"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel shares.
  python tools/summarize_launches.py gpurun_out/X_launches.csv profiles/X_launch_shares.csv "comment"
"""
import collections
import csv
import sys


def main(src, dst, comment=''):
  with open(src) as f:
    lines = [l for l in f if not l.startswith('==')]
  agg = collections.OrderedDict()
  for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
      continue
    k = row['Kernel Name'].split('(')[0].replace('void ', '')
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
  tot = sum(a[1] for a in agg.values())
  with open(dst, 'w') as f:
    f.write('# {}\n'.format(comment))
    f.write('# source: {} ; per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n'.format(src))
    f.write('kernel,launches,total_us,share\n')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      f.write('{},{},{:.1f},{:.4f}\n'.format(k, a[0], a[1], a[1] / tot))


if __name__ == '__main__':
  main(*sys.argv[1:])
