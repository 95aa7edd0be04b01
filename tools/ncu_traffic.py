"""ncu raw-page CSVs (ncu -i X.ncu-rep --page raw --csv) -> a compact per-launch summary CSV and
profiles/ncu_traffic.json (measured DRAM bytes per launch per kernel; bench.py reads it for `roofline.traffic`).

  python tools/ncu_traffic.py profiles/SUMMARY.csv gpurun_out/A_raw.csv [gpurun_out/B_raw.csv ...]
"""
import csv
import json
import os
import sys

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
]


def to_bytes(v, unit):
  v = float(v.replace(',', ''))
  u = unit.lower()
  return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)


def to_us(v, unit):
  v = float(v.replace(',', ''))
  return v * {'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6}.get(unit, 1)


def main(dst, *srcs):
  rows_out, traffic = [], {}
  for src in srcs:
    lines = [l for l in open(src) if not l.startswith('==')]
    rd = list(csv.reader(lines))
    hdr, units = rd[0], rd[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rd[2:]:
      if len(r) < len(hdr):
        continue
      name = r[col['Kernel Name']].split('(')[0].replace('void ', '').split('::')[-1]
      rec = {'kernel': name, 'grid': r[col['Grid Size']], 'block': r[col['Block Size']], 'source': os.path.basename(src)}
      for k in KEEP:
        if k in col:
          rec[k] = r[col[k]]
          rec[k + ' [unit]'] = units[col[k]]
      rows_out.append(rec)
      if 'dram__bytes_read.sum' in col:
        rdb = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
        wrb = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
        us = to_us(r[col['gpu__time_duration.sum']], units[col['gpu__time_duration.sum']])
        t = traffic.setdefault(name, {'launches': 0, 'dram_bytes': 0.0, 'time_us': 0.0})
        t['launches'] += 1
        t['dram_bytes'] += rdb + wrb
        t['time_us'] += us
  keys = ['kernel', 'grid', 'block', 'source'] + [k for k in KEEP if any(k in r for r in rows_out)]
  with open(dst, 'w') as f:
    f.write('# ncu --set full --clock-control none, raw page; one row per captured launch (cold-cache, serialised)\n')
    w = csv.writer(f)
    w.writerow(keys + ['units: ' + '; '.join('{}={}'.format(k, rows_out[0].get(k + ' [unit]', '')) for k in KEEP if k in rows_out[0])])
    for r in rows_out:
      w.writerow([r.get(k, '') for k in keys])
  out = {k: {'dram_bytes_per_launch': v['dram_bytes'] / v['launches'], 'time_us_per_launch': v['time_us'] / v['launches'],
             'launches': v['launches']} for k, v in traffic.items()}
  jpath = os.path.join(os.path.dirname(dst), 'ncu_traffic.json')
  old = json.load(open(jpath)) if os.path.exists(jpath) else {}
  old.update(out)
  json.dump(old, open(jpath, 'w'), indent=1, sort_keys=True)
  for k, v in sorted(out.items()):
    print('{:32s} {:9.1f} us  {:9.1f} MB dram / launch  -> {:7.1f} GB/s'.format(
        k, v['time_us_per_launch'], v['dram_bytes_per_launch'] / 1e6, v['dram_bytes_per_launch'] / v['time_us_per_launch'] / 1e3))


if __name__ == '__main__':
  main(*sys.argv[1:])
