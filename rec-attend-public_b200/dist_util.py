"""Multi-GPU plumbing of the hot path: one process per GPU (torchrun), the batch shards over ranks with no
data-path collective in eval mode; torch.distributed is only used for the barrier and for the
max-over-ranks reduction of the device time.  Backend 'nccl' on GPUs, 'gloo' in the CPU tests."""
import os

import torch
import torch.distributed as dist


def env_world():
  """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
  return (int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')),
          int(os.environ.get('WORLD_SIZE', '1')))


def init(backend, device=None):
  """Initialise the default process group from MASTER_ADDR / MASTER_PORT (torchrun). No-op for 1 rank."""
  rank, local_rank, world = env_world()
  if world > 1 and not dist.is_initialized():
    kw = {}
    if backend == 'nccl' and device is not None:
      kw['device_id'] = device
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)
  return rank, local_rank, world


def barrier():
  if dist.is_available() and dist.is_initialized():
    dist.barrier()


def max_over_ranks(value, device='cpu'):
  """MAX all-reduce of a python float (the step time every rank measured on its own device)."""
  if not (dist.is_available() and dist.is_initialized()):
    return float(value)
  t = torch.tensor([float(value)], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t.item())


def rank_seed(base, config_idx, rank):
  """Every rank draws its own shard of the synthetic global batch."""
  return base + config_idx + 1000 * rank


def average_(tensors):
  """In-place mean over the ranks of a list of same-device tensors (one flattened all-reduce).  No-op for 1 rank."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or not tensors:
    return
  flat = torch.cat([t.reshape(-1) for t in tensors])
  dist.all_reduce(flat, op=dist.ReduceOp.SUM)
  flat /= dist.get_world_size()
  off = 0
  for t in tensors:
    n = t.numel()
    t.copy_(flat[off:off + n].view_as(t))
    off += n


def shard(global_batch, rank, world, require_equal=False):
  """Contiguous shard [lo, hi) of a global batch over `world` ranks (remainder to the low ranks).
  require_equal (TRAINING): the gradient all-reduce averages per-rank means with weight 1/world, which is the
  global-batch mean only for equal shards - uneven splits are refused instead of silently mis-weighted."""
  q, r = divmod(global_batch, world)
  if require_equal and r != 0:
    raise ValueError('global batch {} does not split evenly over {} ranks: pad the batch or pass grad_scale = '
                     'shard_size / global_batch to train_step on every rank'.format(global_batch, world))
  lo = rank * q + min(rank, r)
  return lo, lo + q + (1 if rank < r else 0)


def aggregate_masks_per_sec(world, batch_per_rank, timespan, ms_per_step):
  """Whole-job throughput: masks of all ranks / the slowest rank's time."""
  return world * batch_per_rank * timespan / (ms_per_step / 1e3)


def finalize():
  if dist.is_available() and dist.is_initialized():
    dist.barrier()
    dist.destroy_process_group()
