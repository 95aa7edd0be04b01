"""Per-entry-point device time of ONE eager training-mode forward (KITTI 256x512, T=20, B=32, use_knob)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rec_attend_b200 import _lib, config, synthetic
from rec_attend_b200.full_model import FullModel

opt = dict(config.baseline_opt(2), use_knob=True)
B = 32
batch = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_batch(opt, B, seed=1234).items()}
model = FullModel(opt).load_weights(synthetic.make_weights(opt))
draws = synthetic.make_knob_draws(opt, B, global_step=0, seed=7)
model.forward(batch, phase_train=True, draws=draws, use_graph=False)
torch.cuda.synchronize()
with bench.OpTimer(torch, _lib) as ot:
  torch.cuda._sleep(int(0.3 * 1.9e9))
  model.forward(batch, phase_train=True, draws=draws, use_graph=False)
agg = ot.summary()
tot = {}
for k, d in agg.items():
  e = d['entry']
  t = tot.setdefault(e, [0.0, 0])
  t[0] += d['ms']; t[1] += d['n']
for e, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
  print('{:32s} {:9.3f} ms  {:5d} launches'.format(e, ms, n))
print('sum', sum(v[0] for v in tot.values()))
