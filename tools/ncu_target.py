"""One eager (un-graphed) eval forward - or with --train one taped training forward + backward - of a BASELINE config,
for ncu: every kernel is a separate launch.  python tools/ncu_target.py [--config 2] [--batch B] [--train]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rec_attend_b200 import config, synthetic
from rec_attend_b200.full_model import FullModel

ap = argparse.ArgumentParser()
ap.add_argument('--config', type=int, default=2)
ap.add_argument('--batch', type=int, default=0)
ap.add_argument('--train', action='store_true')
ap.add_argument('--reps', type=int, default=2)
a = ap.parse_args()
cfg = config.BASELINE_CONFIGS[a.config]
opt = dict(config.baseline_opt(a.config), use_knob=bool(a.train))
B = a.batch or cfg['B']
batch = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_batch(opt, B).items()}
model = FullModel(opt).load_weights(synthetic.make_weights(opt))
draws = synthetic.make_knob_draws(opt, B, global_step=0, seed=7) if a.train else None
if a.train:
  from rec_attend_b200 import train as TR
  model._trainer = TR.Trainer(model)
for _ in range(a.reps):
  if a.train:
    model.forward(batch, phase_train=True, draws=draws, use_graph=False, _tape=True)
  else:
    model.forward(batch, use_graph=False)
torch.cuda.synchronize()
print('done')
