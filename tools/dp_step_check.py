"""N-rank check of the data-parallel optimiser tail on real GPUs (run under torchrun):
every rank holds the same weights and its own gradient; after AdamOptimizer.step (NCCL SUM all-reduce of the flat
bucket, x 1/world, weight decay, clip, Adam) all ranks must hold identical parameters, equal to the CPU oracle fed
with the rank-averaged gradient.   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_step_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import rec_attend_b200 as ra
from rec_attend_b200 import dist_util, optim
from oracle import optim as OO

rank, local_rank, world = dist_util.env_world()
torch.cuda.set_device(local_rank)
dist_util.init('nccl', torch.device('cuda', local_rank))
opt = ra.config.full_model_opt('kitti', 64, 128, 3)
w = ra.synthetic.make_weights(opt)
o = optim.AdamOptimizer(opt, w)
fp = o.flat
var = {k: np.asarray(w[k], np.float32) for k in fp.keys}
m = {k: np.zeros_like(var[k]) for k in fp.keys}
v = {k: np.zeros_like(var[k]) for k in fp.keys}
wd = {k: (np.float32(opt['weight_decay']) if optim.has_weight_decay(k) else 0.0) for k in fp.keys}
worst = 0.0
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for step in range(4):
  grads = [np.random.default_rng(100 * step + r).standard_normal(fp.numel).astype(np.float32) * 0.5 for r in range(world)]
  g = torch.from_numpy(grads[rank]).cuda()
  ev0.record()
  lr = o.step(g)
  ev1.record()
  mean = fp.unflatten(np.sum(np.stack(grads).astype(np.float32), axis=0, dtype=np.float32) / np.float32(world))
  var, m, v = OO.adam_step(var, mean, m, v, wd, lr, step + 1)
  got = o.params.cpu().numpy()
  ref = fp.flatten(var)
  worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
torch.cuda.synchronize()
# identical on every rank
chk = torch.tensor([float(o.params.double().sum())], dtype=torch.float64, device='cuda')
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(float(a) == float(allc[0]) for a in allc)
if rank == 0:
  print('dp_step_check: world={} params={} worst_rel_err_vs_oracle={:.2e} identical_across_ranks={} last_step_ms={:.3f}'.format(
      world, fp.numel, worst, same, ev0.elapsed_time(ev1)))
  assert worst < 2e-6 and same
dist_util.finalize()
