"""Kernel timeline of ONE CUDA-graph replay of FullModel.train_step (torch.profiler / CUPTI): kernel time vs span, the gaps
between consecutive kernels and where the large ones sit.  python tools/trace_train_graph.py [--eval]"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from rec_attend_b200 import config, synthetic  # noqa: E402
from rec_attend_b200.full_model import FullModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--eval', action='store_true')
ap.add_argument('--config', type=int, default=2)
ap.add_argument('--batch', type=int, default=0)
ap.add_argument('--by-grid', default='', help='substring of kernel names to break down by launch grid')
ap.add_argument('--dump-step', type=int, default=-1, help='print the kernel sequence of this decode step')
args = ap.parse_args()
cfg = config.BASELINE_CONFIGS[args.config]
opt = dict(config.baseline_opt(args.config), use_knob=not args.eval)
B = args.batch or cfg['B']
batch = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_batch(opt, B).items()}
draws = None if args.eval else synthetic.make_knob_draws(opt, B, global_step=0, seed=7, device='cuda')
model = FullModel(opt).load_weights(synthetic.make_weights(opt))
step = (lambda: model.forward(batch)) if args.eval else (lambda: model.train_step(batch, draws=draws))
for _ in range(3):
  step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
  step()
  torch.cuda.synchronize()
path = '/tmp/trace.json'
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
ev.sort(key=lambda e: e['ts'])


def short(n):
  n = n.replace('void ', '').replace('(anonymous namespace)::', '')
  return n.split('(')[0][:70]


span = ev[-1]['ts'] + ev[-1]['dur'] - ev[0]['ts']
busy = 0.0
end = ev[0]['ts']
gaps = []
for i, e in enumerate(ev):
  s, d = e['ts'], e['dur']
  if s > end:
    gaps.append((s - end, ev[i - 1]['name'], e['name']))
    busy += d
  else:
    busy += max(0.0, s + d - end)
  end = max(end, s + d)
print('events %d  span %.2f ms  busy (union) %.2f ms  sum of durations %.2f ms  idle %.2f ms' %
      (len(ev), span / 1e3, busy / 1e3, sum(e['dur'] for e in ev) / 1e3, (span - busy) / 1e3))
by = collections.defaultdict(lambda: [0, 0.0])
for g, a, b in gaps:
  k = short(a) + '  ->  ' + short(b)
  by[k][0] += 1
  by[k][1] += g
print('largest idle totals by (previous kernel -> next kernel):')
for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:25]:
  print('  %8.1f us  x%-5d avg %6.2f us   %s' % (t, n, t / n, k))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
  agg[short(e['name'])][0] += 1
  agg[short(e['name'])][1] += e['dur']
print('kernel time by name:')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
  print('  %9.1f us  x%-5d %s' % (t, n, k))

if args.dump_step >= 0:
  starts = [i for i, e in enumerate(ev) if 'controller_cluster_kernel' in e['name']]
  if args.dump_step + 1 < len(starts):
    a, b = starts[args.dump_step], starts[args.dump_step + 1]
    t0 = ev[a]['ts']
    print('decode step %d: start (us, relative), duration, end, kernel' % args.dump_step)
    for e in ev[a - 8:b + 1]:
      print('  %9.2f  %8.2f  %9.2f  %s' % (e['ts'] - t0, e['dur'], e['ts'] + e['dur'] - t0, short(e['name'])))

if args.by_grid:
  byg = collections.defaultdict(lambda: [0, 0.0])
  for e in ev:
    if args.by_grid in e['name']:
      k = (short(e['name']), tuple(e.get('args', {}).get('grid', [])))
      byg[k][0] += 1
      byg[k][1] += e['dur']
  print('kernels matching %r by grid:' % args.by_grid)
  for k, (n, t) in sorted(byg.items(), key=lambda kv: -kv[1][1])[:40]:
    print('  %9.1f us  x%-4d avg %7.2f us  %-34s grid %s' % (t, n, t / n, k[0], k[1]))
