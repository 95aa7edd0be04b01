"""CPU tests of the N>1 plumbing with the gloo backend (world_size 2): rendezvous on 127.0.0.1, barrier,
max-over-ranks timing reduction, per-rank shards/seeds and the aggregate-throughput arithmetic that
bench.py uses under torchrun."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, q):
  os.environ.update({'RANK': str(rank), 'LOCAL_RANK': str(rank), 'WORLD_SIZE': str(world),
                     'MASTER_ADDR': '127.0.0.1', 'MASTER_PORT': str(port)})
  import sys
  sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  import rec_attend_b200 as ra
  from rec_attend_b200 import dist_util
  r, lr, w = dist_util.init('gloo')
  assert (r, w) == (rank, world)
  dist_util.barrier()
  # every rank "measures" its own step time; the job reports the slowest
  ms = dist_util.max_over_ranks(10.0 + 5.0 * rank)
  # shards of a global batch are disjoint and cover it; per-rank synthetic batches differ
  lo, hi = dist_util.shard(33, rank, world)
  opt = ra.config.full_model_opt('kitti', 32, 64, 3)
  b = ra.synthetic.make_batch(opt, 2, seed=dist_util.rank_seed(1234, 2, rank))
  chk = torch.tensor([float(b['x'].sum())], dtype=torch.float64)
  gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
  torch.distributed.all_gather(gathered, chk)
  # EMA shadows of a data-parallel run: the mean over ranks, in place, several tensors in one collective
  ema = [torch.full((3,), float(rank + 1)), torch.full((2, 2), 10.0 * (rank + 1))]
  dist_util.average_(ema)
  assert torch.equal(ema[0], torch.full((3,), (1 + world) / 2.0)) and torch.equal(ema[1], torch.full((2, 2), 5.0 * (1 + world)))
  q.put((rank, ms, (lo, hi), [float(g) for g in gathered]))
  dist_util.finalize()


def test_two_rank_gloo_plumbing():
  world = 2
  port = _free_port()
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = sorted(q.get(timeout=120) for _ in range(world))
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert [r[1] for r in res] == [15.0, 15.0], 'max over ranks'
  assert res[0][2] == (0, 17) and res[1][2] == (17, 33)
  assert res[0][3] == res[1][3] and res[0][3][0] != res[0][3][1], 'ranks draw different shards'


def test_single_process_helpers():
  from rec_attend_b200 import dist_util
  assert dist_util.max_over_ranks(3.5) == 3.5
  assert dist_util.aggregate_masks_per_sec(8, 32, 20, 25.0) == pytest.approx(8 * 32 * 20 / 0.025)
  cover = [dist_util.shard(32, r, 8) for r in range(8)]
  assert cover[0] == (0, 4) and cover[-1] == (28, 32)
  assert dist_util.env_world()[2] >= 1
  with pytest.raises(ValueError):
    dist_util.shard(33, 0, 8, require_equal=True)  # training refuses uneven shards (mis-weighted gradient mean)
  assert dist_util.shard(32, 3, 8, require_equal=True) == (12, 16)
  t = [torch.ones(2)]
  dist_util.average_(t)  # single process: no-op
  assert torch.equal(t[0], torch.ones(2))


def _dp_worker(rank, world, port, q):
  """One rank of a data-parallel TRAINING step on the CPU: per-shard gradients (oracle autograd), flat bucket,
  SUM all-reduce over gloo, x 1/world, clip, Adam - the contract optim.AdamOptimizer.step implements on the GPU."""
  os.environ.update({'RANK': str(rank), 'LOCAL_RANK': str(rank), 'WORLD_SIZE': str(world),
                     'MASTER_ADDR': '127.0.0.1', 'MASTER_PORT': str(port)})
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  sys.path.insert(0, root)
  import numpy as np
  import rec_attend_b200 as ra
  from rec_attend_b200 import dist_util, optim
  from oracle import grads as OG
  from oracle import optim as OO
  torch.set_num_threads(2)
  dist_util.init('gloo')
  opt = ra.config.full_model_opt('cvppp', 64, 64, 2, use_knob=False)
  weights = ra.synthetic.make_weights(opt, seed=4321)
  full = ra.synthetic.make_batch(opt, 4, seed=21)
  lo, hi = dist_util.shard(4, rank, world)
  shard = {k: v[lo:hi] for k, v in full.items()}
  g, _ = OG.full_model_grads(opt, weights, shard, include_weight_decay=False)
  flat = optim.FlatParams(weights)
  bucket = torch.from_numpy(flat.flatten(g))
  n = optim.all_reduce_sum_(bucket)
  mean = flat.unflatten((bucket / n).numpy())
  keys = flat.keys
  var = {k: np.asarray(weights[k], np.float32) for k in keys}
  zeros = {k: np.zeros_like(var[k]) for k in keys}
  wd = {k: (np.float32(opt['weight_decay']) if optim.has_weight_decay(k) else 0.0) for k in keys}
  new, _, _ = OO.adam_step(var, mean, zeros, zeros, wd, optim.learn_rate(opt, 0), 1)
  q.put((rank, n, {k: new[k] for k in ('ctrl_lstm_w_xi', 'attn_dcnn_w_3', 'ctrl_cnn_0_1_gamma')},
         {k: g[k] for k in ('ctrl_lstm_w_xi',)}))
  dist_util.finalize()


def test_two_rank_data_parallel_training_step_contract():
  """Both ranks end with identical parameters, equal to the single-process oracle step fed with the other rank's
  gradient (mean over ranks, clip AFTER averaging; per-rank BN statistics, §9.12)."""
  import numpy as np
  import rec_attend_b200 as ra
  from oracle import grads as OG
  world, port = 2, _free_port()
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda r: r[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert res[0][1] == res[1][1] == 2
  for k in res[0][2]:
    assert np.array_equal(res[0][2][k], res[1][2][k]), k  # replicas stay bit-identical
  # single-process reference: rank 0's shard + rank 1's gradient handed in as world_grads
  opt = ra.config.full_model_opt('cvppp', 64, 64, 2, use_knob=False)
  weights = ra.synthetic.make_weights(opt, seed=4321)
  full = ra.synthetic.make_batch(opt, 4, seed=21)
  keys = OG.trainable_keys(weights)
  g1, _ = OG.full_model_grads(opt, weights, {k: v[2:4] for k, v in full.items()}, include_weight_decay=False)
  zeros = {k: np.zeros_like(weights[k]) for k in keys}
  ref, _, _, _ = OG.train_step(opt, weights, {k: v[0:2] for k, v in full.items()}, zeros, zeros, 0, world_grads=[g1])
  for k in res[0][2]:
    assert np.allclose(res[0][2][k], ref[k], rtol=0, atol=2e-6), k  # lr = 1e-3: agreement to 0.2 % of one step
  assert not np.allclose(res[0][3]['ctrl_lstm_w_xi'], res[1][3]['ctrl_lstm_w_xi'])  # the shards' gradients differ
