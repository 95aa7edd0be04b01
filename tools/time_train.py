"""Timing of FullModel.train_step at a BASELINE config: ms per step (CUDA graph), eager per-entry-point breakdown.
  python tools/time_train.py [--config 2] [--batch B] [--knob 0|1]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from rec_attend_b200 import _lib, config, synthetic  # noqa: E402
from rec_attend_b200.full_model import FullModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--config', type=int, default=2)
ap.add_argument('--batch', type=int, default=0)
ap.add_argument('--knob', type=int, default=1)
ap.add_argument('--steps', type=int, default=3)
args = ap.parse_args()
cfg = config.BASELINE_CONFIGS[args.config]
opt = dict(config.baseline_opt(args.config), use_knob=bool(args.knob))
B = args.batch or cfg['B']
batch = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_batch(opt, B).items()}
draws = synthetic.make_knob_draws(opt, B, global_step=0, seed=7, device='cuda') if args.knob else None
model = FullModel(opt).load_weights(synthetic.make_weights(opt))
for _ in range(2):
  model.train_step(batch, draws=draws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
  r = model.train_step(batch, draws=draws)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
res = {'workload': cfg['name'], 'B': B, 'knob': args.knob, 'train_step_ms': ms, 'masks_per_s': B * cfg['T'] / ms * 1e3,
       'loss': float(r['loss']), 'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9}
# the same step without the CUDA graph (host-launched): separates graph replay cost from kernel time
for _ in range(1):
  model.train_step(batch, draws=draws, use_graph=False)
torch.cuda.synchronize()
e0.record()
for _ in range(args.steps):
  model.train_step(batch, draws=draws, use_graph=False)
e1.record()
torch.cuda.synchronize()
res['train_step_eager_ms'] = e0.elapsed_time(e1) / args.steps
lib = _lib.lib()
n0 = lib.ra_launch_count()
with bench.OpTimer(torch, _lib) as ot:
  torch.cuda._sleep(int(0.2 * 1.9e9))
  model.forward(batch, phase_train=True, draws=draws, use_graph=False, _tape=True)
agg = ot.summary()
res['launches_fwd_bwd'] = int(lib.ra_launch_count() - n0)
groups = {}
for key, d in agg.items():
  g = d['tag'] + ':' + d['entry']
  gg = groups.setdefault(g, {'ms': 0.0, 'n': 0})
  gg['ms'] += d['ms']
  gg['n'] += d['n']
res['eager_kernel_ms'] = sum(d['ms'] for d in groups.values())
res['per_shape'] = {k: round(d['ms'], 3) for k, d in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])
                    if k.startswith(('wgrad', 'bn_bwd', 'bn_fwd'))}
res['groups'] = {k: {'ms': round(v['ms'], 3), 'n': v['n']} for k, v in sorted(groups.items(), key=lambda kv: -kv[1]['ms'])[:40]}
print(json.dumps(res, indent=1))
