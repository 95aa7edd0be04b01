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
