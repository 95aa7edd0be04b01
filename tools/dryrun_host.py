"""CPU dry run of the host control flow: every C-ABI call becomes a no-op, tensors live on the CPU."""
import sys, types
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rec_attend_b200 as ra
from rec_attend_b200 import _lib, ops, full_model, box_model, train as TR

calls = []
real_call = _lib.call
def fake_call(name, *a):
  calls.append(name)
  if 'plan' in name or 'layout' in name:
    return real_call(name, *a)
_lib.call = fake_call
class FakeStream:
  cuda_stream = 0
  def wait_stream(self, s): pass
  def wait_event(self, e): pass
class FakeEvent:
  def record(self, s=None): pass
torch.cuda.is_available = lambda: True
torch.cuda.current_device = lambda: 0
torch.cuda.current_stream = lambda *a: FakeStream()
torch.cuda.is_current_stream_capturing = lambda: False
torch.cuda.Event = FakeEvent
import contextlib
torch.cuda.stream = lambda s: contextlib.nullcontext()
ops._chk = lambda *a: None
_lib.lib().ra_pairwise_iou_workspace = lambda *a: 16

_orig_device = torch.device
def init(self, opt, device=None):
  pass
orig_init = full_model._ModelBase.__init__
def patched_init(self, opt, device=None):
  orig_init(self, opt, device='cpu')
full_model._ModelBase.__init__ = patched_init
# torch.tensor(..., device=cuda) paths: chan_map uses self.device which is set after; patch torch.device('cuda', idx)
real_td = torch.device
class DevShim:
  pass
import builtins
def run(kind, arch, knob):
  if kind == 'full':
    opt = ra.config.full_model_opt(arch, 64, 128, 2, use_knob=knob)
    B = 3
    batch = ra.synthetic.make_batch(opt, B, seed=1)
    w = ra.synthetic.make_weights(opt, seed=1)
    draws = ra.synthetic.make_knob_draws(opt, B, global_step=9000, seed=3) if knob else None
    m = full_model.FullModel(opt).load_weights(w)
    m.forward(batch, use_graph=False)
    r = m.train_step(batch, draws=draws, use_graph=False)
    tb = m._trainer._scatter[B]
    print(kind, arch, knob, 'covered', tb['covered'], 'of', m._trainer.optim.params.numel(), 'sync segs', m._sync_table['nseg'], m._sync_table['total'])
    m.forward(batch, use_graph=False)
  else:
    opt = ra.config.box_model_opt(64, 128, 2, use_iou_box=knob)
    B = 3
    batch = ra.synthetic.make_batch(opt, B, seed=1)
    w = ra.synthetic.make_weights(opt, seed=1, model='box')
    m = box_model.BoxModel(opt).load_weights(w)
    m.forward(batch, use_graph=False)
    r = m.train_step(batch, use_graph=False)
    tb = m._trainer._scatter[B]
    print(kind, knob, 'covered', tb['covered'], 'of', m._trainer.optim.params.numel())
for a in [('full','cvppp',False),('full','kitti',False),('full','kitti',True),('full','cityscapes',True),('box','',False),('box','',True)]:
  run(*a)
print(len(calls), 'calls; distinct', len(set(calls)))
