#!/usr/bin/env python
"""Benchmark of the recurrent-attention decoding hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config IDX]

A "step" is one full eval-mode forward of the model over one synthetic batch: the T-step
decode loop (controller CNN + glimpse LSTM + Gaussian glimpse + patch CNN + deconv mask head
+ paste-back/canvas) plus the matching loss block (pairwise IoU, Hungarian, losses).
Metric: instance-masks/sec = (masks produced by all ranks) / (max-over-ranks device time).
Default workload = BASELINE.json configs[2] (KITTI-arch 256x512, T=20, B=32 per GPU), the
configuration the metric is quoted on; it fits one GPU.  Multi-GPU: one process per GPU
(torchrun), the batch shards with no data-path collective (eval forward) -> weak scaling.

The same line carries, under "train_step", the TRAINING step of the same workload (taped
training-mode forward + backward + NCCL all-reduce of the flat gradient bucket + clip/Adam +
device-side weight re-pack, all inside the timed region) - weak (B per GPU fixed) and strong
(BASELINE configs[2] as worded: B=32 TOTAL, sharded over the ranks) - and under "strong" the
eval forward at B=32 total.

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for the roofline accounting.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def load_peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    d = json.load(open(p))
    return {'hbm_gbs': float(d['hbm_gbs']), 'bf16_tflops': float(d['bf16_tflops']), 'source': 'measured'}
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'source': 'fallback'}


# C-ABI entry point -> the CUDA kernels it launches (names as ncu prints them), for `traffic`
ENTRY_KERNELS = {
    'ra_conv3x3_umma_chain_run': ['conv3x3_umma_chain_kernel'],
    'ra_gaussian_extract_f32': ['extract_rows_kernel<4>', 'extract_cols_kernel'],
    'ra_paste_back_f32': ['paste_back_kernel'],
    'ra_pairwise_iou_f32': ['pairwise_iou_kernel'],
    'ra_gt_box_f32': ['gt_box_kernel'],
    'ra_canvas_conv_f32': ['canvas_conv_bulk_kernel<2>'],
    'ra_conv3x3_umma_f32': ['conv3x3_umma_kernel'],
}


def load_traffic():
  """Measured DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum from the committed ncu --set full
  captures, profiles/ncu_traffic.json written by tools/ncu_traffic.py).  {} if the file is absent."""
  p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
  try:
    return json.load(open(p))
  except (OSError, ValueError):
    return {}


def entry_traffic(traffic, entry):
  ks = ENTRY_KERNELS.get(entry.split(':')[0])
  if not ks or any(k not in traffic for k in ks):
    return None
  return float(sum(traffic[k]['dram_bytes_per_launch'] for k in ks))


class ClockSampler(object):
  """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
  Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index = index
    self.lines = []
    self.proc = None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True)
      self.thread = threading.Thread(target=self._read)
      self.thread.daemon = True
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for l in self.lines:
      f = [t.strip() for t in l.split(',')]
      if len(f) < 7:
        continue
      try:
        sm.append(float(f[0]))
        mx.append(float(f[1]))
      except ValueError:
        continue
      for n, v in zip(names, f[3:7]):
        if v.lower().startswith('active'):
          reasons.add(n)
    sm.sort()
    return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------- roofline model
def kernel_algorithmic_work(opt, B, Bc=None):
  """Algorithmic bytes / flops PER LAUNCH of the HBM-/tensor-bound kernels.
  Keys are the C-ABI entry points.  Decode-loop kernels process one step of Bc images per launch (Bc = B / chains,
  FullModel._chains); the loss-side kernels see the whole batch B."""
  Bc = B if Bc is None else Bc
  H, W, T, F = opt['inp_height'], opt['inp_width'], opt['timespan'], opt['filter_height']
  from rec_attend_b200.config import input_depths
  D = input_depths(opt)[0]
  c0 = opt['ctrl_cnn_depth'][0]
  p0 = opt['ctrl_cnn_pool'][0]
  work = {
      # (H*W*D + F^2*D)*4 per image-step
      'ra_gaussian_extract_f32': {'bytes': Bc * (H * W * D + F * F * D) * 4.0},
      # (F^2 + 3*H*W)*4 (read patch + read canvas + write y_out + write canvas) + H*W*4 (attn_box write)
      'ra_paste_back_f32': {'bytes': Bc * (F * F + 4 * H * W) * 4.0},
      # pairwise IoU: 2*B*T*H*W*4 per call
      'ra_pairwise_iou_f32': {'bytes': 2.0 * B * T * H * W * 4, 'flops': 2.0 * B * T * T * H * W},
      'ra_gt_box_f32': {'bytes': 1.0 * B * T * H * W * 4},
  }
  # controller CNN: flops per decode step (all 8 layers, as the reference computes them)
  ch = [D] + list(opt['ctrl_cnn_depth'])
  h, w, fl, fl_umma = H, W, 0.0, 0.0
  for i, pl in enumerate(opt['ctrl_cnn_pool']):
    f_i = 2.0 * h * w * ch[i + 1] * 9 * ch[i]
    fl += f_i
    if i > 0:  # layer 0 runs as `prepare` (static half, once per forward) + ra_canvas_conv_f32 (per step)
      fl_umma += f_i
    h, w = h // pl, w // pl
  work['ctrl_cnn_step'] = {'flops': B * fl}  # all chains together: one decode step of the whole batch
  work['ctrl_cnn_umma_step'] = {'flops': B * fl_umma}  # the layers the tcgen05 group `ctrl_cnn` actually runs
  # first controller layer, per-step half: read static_pre + canvas at full resolution, write the pooled output
  work['ra_canvas_conv_f32'] = {'bytes': Bc * (H * W * (c0 + 1) + (H // p0) * (W // p0) * c0) * 4.0}
  return work


class OpTimer(object):
  """CUDA-event timing of every C-ABI call (on the stream the kernels are launched on)."""

  def __init__(self, torch, lib_mod):
    self.torch = torch
    self.lib_mod = lib_mod
    self.records = []
    self.orig = None

  def __enter__(self):
    self.orig = self.lib_mod.call
    torch = self.torch

    def timed_call(name, *args):
      e0 = torch.cuda.Event(enable_timing=True)
      e1 = torch.cuda.Event(enable_timing=True)
      e0.record()
      self.orig(name, *args)
      e1.record()
      self.records.append((name, args, e0, e1, self.lib_mod.TAG))

    self.lib_mod.call = timed_call
    return self

  def __exit__(self, *a):
    self.lib_mod.call = self.orig

  def summary(self):
    self.torch.cuda.synchronize()
    agg = {}
    for name, args, e0, e1, tag in self.records:
      key = name
      if name == 'ra_conv3x3_f32':
        # args: x1,C1,x2,C2,w,scale,shift,add_to,B,Hin,Win,Cout,up,pool,relu,y,stream
        key = '{}:conv3x3[{}x{} {}+{}->{} up{} pool{}]'.format(tag, args[9], args[10], args[1], args[3], args[11],
                                                             args[12], args[13])
      if name == 'ra_conv3x3_umma_f32':
        # args: x1,C1,x2,C2,wpack,scale,shift,B,Hin,Win,Cout,up,pool,relu,y,stream
        key = '{}:conv3x3_umma[{}x{} {}+{}->{} up{} pool{}]'.format(tag, args[8], args[9], args[1], args[3], args[10],
                                                                  args[11], args[12])
      if name == 'ra_conv3x3_bwd_weight_ex_f32':
        # args: x1,C1,x1_bmod,x2,C2,d_out,B,Hin,Win,Cout,up,ws,dw,db,stream
        key = 'wgrad[N{} {}x{} {}+{}->{} up{}]'.format(args[6], args[7], args[8], args[1], args[4], args[9], args[10])
      if name == 'ra_bn_train_block_bwd_grouped_f32':
        # args: raw,dy,gamma,beta,mean,var,G,B,H,W,C,pool,...
        key = 'bn_bwd[G{} B{} {}x{}x{} pool{}]'.format(args[6], args[7], args[8], args[9], args[10], args[11])
      if name == 'ra_bn_train_block_f32':
        # args: x,B,H,W,C,gamma,beta,eps,decay,pool,relu,...
        key = 'bn_fwd[B{} {}x{}x{} pool{}]'.format(args[1], args[2], args[3], args[4], args[9])
      d = agg.setdefault(key, {'entry': name, 'tag': tag, 'ms': 0.0, 'n': 0})
      d['ms'] += e0.elapsed_time(e1)
      d['n'] += 1
    return agg


# --------------------------------------------------------------------------- reference arm
def cpu_reference_time(opt, B_sample, steps, warmup, seed=1234, model='full'):
  """The reference's CPU implementation of the path = the structure-faithful oracle
  (PyTorch-CPU restatement + C restatement of hungarian.cc), all host threads."""
  import torch
  import rec_attend_b200.config  # noqa: F401  (pure-python config/synthetic only; no CUDA)
  from rec_attend_b200 import synthetic
  from oracle import model as OM
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  batch = synthetic.make_batch(opt, B_sample, seed=seed)
  weights = synthetic.make_weights(opt, model=model)
  fwd = OM.box_model_forward if model == 'box' else OM.full_model_forward  # BASELINE configs[4] is the box model
  times = []
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    with torch.no_grad():
      fwd(opt, weights, batch)
    dt = time.perf_counter() - t0
    if i >= warmup:
      times.append(dt)
  return sum(times) / len(times), cores


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return 0
  from rec_attend_b200 import config
  cfg = config.BASELINE_CONFIGS[args.config]
  opt = config.baseline_opt(args.config)
  B_sample = args.ref_batch
  is_box = cfg.get('model', 'full') == 'box'
  sec, cores = cpu_reference_time(opt, B_sample, max(1, args.steps), max(0, args.warmup),
                                  model='box' if is_box else 'full')
  masks = B_sample * cfg['T']
  val = masks / sec
  metric, unit = ('box-steps/sec', 'box-steps/s') if is_box else ('instance-masks/sec', 'masks/s')
  line = {
      'impl': 'reference', 'metric': metric, 'value': val, 'unit': unit, 'n_gpus': args.gpus,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': cfg['name'], 'arch': cfg.get('arch'), 'batch_per_gpu': cfg['B'], 'timespan': cfg['T'],
                 'height': cfg['H'], 'width': cfg['W'], 'parallelism': 'cpu',
                 'sample': 'throughput measured on B={} examples of the workload batch (full T={} decode + loss block '
                           'each): masks/s is per-example work, so it is the same metric on a bounded sample'.format(
                               B_sample, cfg['T'])},
      'cpu_baseline': {'value': val, 'unit': unit, 'cores': cores, 'kind': 'port',
                       'sample': 'oracle (PyTorch-CPU restatement + C hungarian), B={} x T={} at {}x{}'.format(
                           B_sample, cfg['T'], cfg['H'], cfg['W'])},
      'e2e': {'value': val, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  emit(line)
  return 0


# --------------------------------------------------------------------------- our arm
def run_ours(args):
  import torch
  from rec_attend_b200 import _lib, config, synthetic
  from rec_attend_b200.full_model import FullModel

  from rec_attend_b200 import dist_util
  rank, local_rank, world = dist_util.env_world()
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
  torch.cuda.set_device(local_rank)
  dist_util.init('nccl', torch.device('cuda', local_rank))

  cfg = config.BASELINE_CONFIGS[args.config]
  if cfg['model'] != 'full':
    return run_box(args, cfg, rank, local_rank, world)
  opt = config.baseline_opt(args.config)
  B = args.batch or cfg['B']
  T = cfg['T']
  batch_np = synthetic.make_batch(opt, B, seed=dist_util.rank_seed(1234, args.config, rank))
  weights = synthetic.make_weights(opt)
  model = FullModel(opt).load_weights(weights)
  lib = _lib.lib()

  # device-resident inputs for `value`; pinned host copies for `e2e`.  The {0,1} ground-truth masks travel as
  # uint8 (the datasets hold PNG masks, data_api/ins_seg_dataset.py:169-172) and are expanded on the device.
  from rec_attend_b200 import postprocess as PP
  dev_batch = {k: torch.from_numpy(v).cuda() for k, v in batch_np.items()}
  host_np = dict(batch_np, y_gt=batch_np['y_gt'].astype('uint8'))
  pinned = {k: torch.from_numpy(v).pin_memory() for k, v in host_np.items()}
  fetch = ['loss', 'segm_loss', 'box_loss', 'conf_loss', 'iou_soft', 's_out', 'match']
  fetch_e2e = fetch + ['y_out']
  host_out = {}

  def step_device():
    return model.forward(dev_batch, outputs=fetch)

  # two pinned host copies of the step's inputs, used alternately (a real loader hands over a new batch each step)
  pinned2 = {k: v.clone().pin_memory() for k, v in pinned.items()}
  host_batches = [pinned, pinned2]
  e2e_state = {'i': 0}

  def step_e2e():
    # public API call with HOST buffers: H2D of every input, forward, D2H of the step's results.  The H2D of
    # step i+1 is issued (model.prefetch) before step i computes, i.e. the copies are double buffered; every
    # step still copies all of its inputs from pinned host memory inside the timed region.
    i = e2e_state['i']
    cur, nxt = host_batches[i % 2], host_batches[(i + 1) % 2]
    e2e_state['i'] = i + 1
    if i == 0:
      model.prefetch(cur)
    out = model.forward(cur, outputs=fetch_e2e)
    model.prefetch(nxt)
    # what full_model_eval.py:100-125 does with y_out / s_out: confidence-weighted one-label map, on the device
    pp = PP.postprocess(out['y_out'], out['s_out'], thresh=0.3)
    res = {k: out[k] for k in fetch}
    res['label'], res['conf'] = pp['label'], pp['conf']
    for k, v in res.items():
      if k not in host_out:
        host_out[k] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
    # The small results are static outputs of the CUDA graph (the next forward overwrites them): their D2H stays on the
    # compute stream.  The label map and the confidences (16.8 MB, fresh tensors of the post-processing) leave on a
    # copy stream, so that the next step's forward overlaps their D2H like its H2D overlaps this step's compute; the
    # timed region ends with a device-wide synchronize, i.e. it contains every copy of every step.
    for k in fetch:
      host_out[k].copy_(res[k], non_blocking=True)
    cur_stream = torch.cuda.current_stream()
    if 'd2h_stream' not in e2e_state:
      e2e_state['d2h_stream'] = torch.cuda.Stream()
    d2h = e2e_state['d2h_stream']
    ready = torch.cuda.Event()
    ready.record(cur_stream)
    with torch.cuda.stream(d2h):
      d2h.wait_event(ready)
      for k in ('label', 'conf'):
        # (the pinned destination of step i is rewritten by step i+1 on the same copy stream: ordered)
        host_out[k].copy_(res[k], non_blocking=True)
        res[k].record_stream(d2h)
    return res

  def barrier():
    dist_util.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps):
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    n0 = lib.ra_launch_count()
    e0.record()
    for _ in range(steps):
      fn()
    if 'd2h_stream' in e2e_state:  # the last step's label-map D2H must end inside the timed region
      torch.cuda.current_stream().wait_stream(e2e_state['d2h_stream'])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.ra_launch_count() - n0
    ms = dist_util.max_over_ranks(ms, device='cuda')
    return ms, launches

  for _ in range(max(3, args.warmup)):
    step_device()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  ms, launches = timed(step_device, args.steps)
  for _ in range(max(4, args.warmup)):  # untimed: first touches of the pinned buffers, copy-stream creation, graph capture
    step_e2e()
  ms_e2e, _ = timed(step_e2e, args.steps)
  clocks = sampler.stop() if rank == 0 else None  # sampled over both timed regions (device-resident and end-to-end)

  # ---- the TRAINING step of the same workload: taped training-mode forward (batch-statistics BN, EMA update,
  # scheduled sampling with explicit draws) + backward + NCCL all-reduce(SUM) of the flat gradient bucket + clip / Adam
  # + device-side re-pack of every weight image, all inside the timed region (FullModel.train_step).
  import torch.distributed as dist

  def bench_train(Bt, seed_off):
    opt_t = dict(opt, use_knob=True)
    bt_np = synthetic.make_batch(opt_t, Bt, seed=dist_util.rank_seed(1234, args.config, rank) + seed_off)
    bt = {k: torch.from_numpy(v).cuda() for k, v in bt_np.items()}
    model_t = FullModel(opt_t).load_weights(weights)
    # the scheduled-sampling draws are drawn ON the device, fresh every step and inside the timed region, like the
    # reference's tf.random_uniform nodes (full_model.py:573-579,608-610,836-837): B*T*H*W floats of canvas noise
    dev_name = 'cuda:%d' % torch.cuda.current_device()
    draw_no = [0]

    def new_draws():
      draw_no[0] += 1
      return synthetic.make_knob_draws(opt_t, Bt, global_step=0, seed=7 + rank + 1000 * draw_no[0], device=dev_name)

    draws = new_draws()

    def step_train():
      return model_t.train_step(bt, draws=new_draws())

    for _ in range(2):
      step_train()
    n_t = max(2, min(args.steps, 5))
    ms_t, launches_t = timed(step_train, n_t)
    tr = model_t._trainer
    # the pieces, each timed alone (CUDA events, max over ranks): forward+backward graph, all-reduce, optimiser tail
    ms_g, _ = timed(lambda: model_t.forward(bt, phase_train=True, draws=draws, _tape=True), n_t)
    ms_ar = 0.0
    if world > 1:
      ms_ar, _ = timed(lambda: dist.all_reduce(tr.grad_flat, op=dist.ReduceOp.SUM), 20)
    ms_tail, _ = timed(lambda: tr.apply(), n_t)
    n0 = lib.ra_launch_count()  # kernels of one step, counted in an eager (un-graphed) step: the graph replays the same
    model_t.forward(bt, phase_train=True, draws=draws, use_graph=False, _tape=True)
    tr.apply()
    torch.cuda.synchronize()
    launches_eager = int(lib.ra_launch_count() - n0)
    in_sync = True
    if world > 1:
      ref = tr.optim.params.clone()
      dist.broadcast(ref, src=0)
      diff = (ref - tr.optim.params).abs().max()
      dist.all_reduce(diff, op=dist.ReduceOp.MAX)
      in_sync = bool(float(diff) == 0.0)
    res = {'batch_per_gpu': Bt, 'global_batch': Bt * world, 'ms_per_step': ms_t / n_t, 'steps': n_t,
           'value': world * Bt * T / (ms_t / n_t / 1e3), 'unit': 'masks/s',
           'gpu_launches_per_step': launches_eager,
           'fwd_bwd_graph_ms': ms_g / n_t, 'allreduce_us': ms_ar / 20 * 1e3,
           'optimiser_tail_ms': ms_tail / n_t, 'bucket_bytes': int(tr.grad_flat.numel() * 4),
           'params_in_sync_across_ranks': in_sync, 'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9}
    del model_t, bt
    torch.cuda.empty_cache()
    return res

  train_step = None
  strong = None
  if not args.no_train_step:
    train_step = {'weak': bench_train(B, 0),
                  'step': 'train_step = taped training-mode forward + backward + all-reduce(SUM) of the flat gradient '
                          'bucket (NCCL, inside the timed region) + clip + Adam + device-side weight re-pack',
                  'scheduled_sampling': 'use_knob=True, knob probabilities of global_step 0; fresh draws every step, '
                                        'generated on the device inside the timed region'}
    Bs = cfg['B'] // world
    if world > 1 and Bs >= 1 and cfg['B'] % world == 0:
      train_step['strong'] = bench_train(Bs, 500)
      train_step['strong']['scaling'] = 'strong'
    else:
      train_step['strong'] = dict(train_step['weak'], scaling='strong (N=1: identical to weak)')
    train_step['weak']['scaling'] = 'weak'
  # eval forward, strong scaling: BASELINE configs[2] as worded - B = 32 TOTAL sharded over the ranks
  Bs = cfg['B'] // world
  if world > 1 and Bs >= 1 and cfg['B'] % world == 0:
    bs_np = synthetic.make_batch(opt, Bs, seed=dist_util.rank_seed(1234, args.config, rank) + 900)
    bs_dev = {k: torch.from_numpy(v).cuda() for k, v in bs_np.items()}
    for _ in range(3):
      model.forward(bs_dev, outputs=fetch)
    ms_s, _ = timed(lambda: model.forward(bs_dev, outputs=fetch), args.steps)
    strong = {'scaling': 'strong', 'global_batch': cfg['B'], 'batch_per_gpu': Bs, 'ms_per_step': ms_s / args.steps,
              'value': cfg['B'] * T / (ms_s / args.steps / 1e3), 'unit': 'masks/s'}

  value = dist_util.aggregate_masks_per_sec(world, B, T, ms / args.steps)
  e2e_value = dist_util.aggregate_masks_per_sec(world, B, T, ms_e2e / args.steps)
  h2d = sum(v.numel() * v.element_size() for v in pinned.values())
  d2h = sum(v.numel() * v.element_size() for v in host_out.values())

  line = None
  if rank == 0:
    # ---- per-kernel device times (CUDA events around every C-ABI call, one extra instrumented step)
    peaks = load_peaks()
    n0 = lib.ra_launch_count()
    with OpTimer(torch, _lib) as ot:
      # The eager step is enqueued behind a ~60 ms spin kernel: the host runs ahead, so the GPU executes the
      # event / kernel / event triples back to back and the intervals hold no host launch latency.
      torch.cuda._sleep(int(0.06 * torch.cuda.get_device_properties(local_rank).clock_rate * 1e3))
      model.forward(dev_batch, outputs=fetch, use_graph=False)  # eager: one C-ABI call per kernel group
    agg = ot.summary()
    launches_per_step = int(lib.ra_launch_count() - n0)
    chains = model._chains(B)
    work = kernel_algorithmic_work(opt, B, chains[0][1] - chains[0][0])
    total_ms = sum(d['ms'] for d in agg.values())
    groups = {}
    for key, d in agg.items():
      g = d['entry']
      if g in ('ra_conv3x3_f32', 'ra_conv3x3_umma_f32', 'ra_conv3x3_umma_chain_run'):
        g = g + ':' + d['tag']
      gg = groups.setdefault(g, {'ms': 0.0, 'n': 0})
      gg['ms'] += d['ms']
      gg['n'] += d['n']
    kernels = {}
    # the committed ncu capture is of the default workload: no `traffic` claim for any other configuration
    traffic = load_traffic() if (args.config == 2 and B == cfg['B']) else {}
    for g, d in sorted(groups.items(), key=lambda kv: -kv[1]['ms']):
      ent = {'ms_per_step': round(d['ms'], 4), 'launches': d['n'], 'share': round(d['ms'] / total_ms, 4)}
      tr = entry_traffic(traffic, g)
      if tr is not None:
        ent['traffic'] = tr
      if g in work and 'bytes' in work[g]:
        gbs = work[g]['bytes'] / (d['ms'] / d['n'] / 1e3) / 1e9
        ent['achieved_gbs'] = round(gbs, 1)
        ent['hbm_frac'] = round(gbs / peaks['hbm_gbs'], 4)
      kernels[g] = ent
    # controller-CNN as a group: flops of the 8 conv layers of one decode step / their time
    ccnn_ms = sum(d['ms'] for k, d in agg.items()
                  if d['entry'] in ('ra_conv3x3_f32', 'ra_conv3x3_umma_f32', 'ra_conv3x3_umma_chain_run') and
                  d['tag'] == 'ctrl_cnn')
    conv_tflops = work['ctrl_cnn_umma_step']['flops'] * T / (ccnn_ms / 1e3) / 1e12 if ccnn_ms > 0 else 0.0
    dom = max(groups.items(), key=lambda kv: kv[1]['ms'])[0]
    if dom.startswith('ra_conv3x3'):
      roofline = {
          'kernel': 'conv3x3_umma (controller CNN layers 1-7, tcgen05 kind::f16 hi/lo operand split - 3 fp16 MACs per '
                    'fp32 MAC, persistent grid, one launch per layer chained by programmatic dependent launch; group = ' +
                    dom + ')',
          'bound': 'tensor',
          'achieved': round(conv_tflops, 3), 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
          'frac': round(conv_tflops / peaks['bf16_tflops'], 5), 'traffic': entry_traffic(traffic, dom),
          'peak_source': peaks['source'] + ' (cuBLAS bf16 burst)',
          'note': 'flops = controller-CNN layers 1-7 (the layers this kernel group runs; layer 0 is the linear split '
                  'prepare + ra_canvas_conv_f32) per decode step x T / their summed CUDA-event time; '
                  'the kernel issues 3 fp16 MACs (hi x hi, hi x lo, lo x hi) per algorithmic MAC for fp32 parity '
                  '(DESIGN 4.1), so frac = 1/3 would be a saturated tensor pipe for this formulation; the layers have '
                  '16-64 output channels and the kernel is bound by shared-memory bandwidth (every filter tap re-reads the '
                  'A operand; operand conversion and pooling go through shared memory too), not by the tensor pipe '
                  '(DESIGN 4.1, profiles/r04j_*, r04s_*)'
      }
    else:
      k = kernels[dom]
      roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': k.get('achieved_gbs'), 'peak': peaks['hbm_gbs'],
                  'unit': 'GB/s', 'frac': k.get('hbm_frac'), 'traffic': k.get('traffic'), 'peak_source': peaks['source']}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
      sec, cores = cpu_reference_time(opt, args.ref_batch, 2, 1)
      cpu_baseline = {
          'value': args.ref_batch * T / sec, 'unit': 'masks/s', 'cores': cores, 'kind': 'port',
          'sample': 'oracle (PyTorch-CPU restatement + C hungarian), B={} x T={} at {}x{}, mean of 2 runs after 1 warm-up'.format(
              args.ref_batch, T, cfg['H'], cfg['W'])
      }
    layers = {k: {'ms': round(d['ms'], 4), 'n': d['n']} for k, d in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])
              if 'conv3x3' in k}
    line = {
        'metric': 'instance-masks/sec', 'value': value, 'unit': 'masks/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['name'], 'arch': cfg['arch'], 'batch_per_gpu': B, 'timespan': T,
                   'height': cfg['H'], 'width': cfg['W'], 'parallelism': 'dp{}'.format(world),
                   'l2': 'inputs {:.0f} MB + per-step working set exceed the 126 MB L2'.format(h2d / 1e6),
                   'step': 'eval forward (T-step decode) + matching loss block',
                   'graph': 'one CUDA graph per forward ({} sub-batch chain{}); gt-box, hard-IoU and score-head '
                            'kernels on parallel branches'.format(len(chains), '' if len(chains) == 1 else 's')},
        'e2e': {'value': e2e_value, 'unit': 'masks/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps, 'fetch': fetch + ['label', 'conf'],
                'inputs': 'x, d_in, y_in fp32; y_gt uint8 {0,1} expanded on the device (ra_u8_to_f32)',
                'outputs': 'loss scalars, s_out, match + the int32 instance label map [B,H,W] and conf of '
                           'ra_postprocess_f32 (full_model_eval.py:100-125), D2H inside the timed region',
                'pipeline': 'H2D of step i+1 overlaps the compute of step i (double-buffered static inputs); the D2H of the '
                         'label maps of step i overlaps the forward of step i+1 (copy stream)'},
        'gpu_launches': launches_per_step * args.steps,
        'gpu_launches_note': '{} kernels of librecattend_b200.so per step (counted in one eager step); the timed '
                             'steps replay exactly these launches from one CUDA graph per step'.format(launches_per_step),
        'clocks': clocks,
        'roofline': roofline,
        'kernels': kernels,
        'conv_layers': layers,
        'cpu_baseline': cpu_baseline,
        'train_step': train_step,
        'strong': strong,
    }
    emit(line)
  dist_util.finalize()
  return 0


def run_box(args, cfg, rank, local_rank, world):
  """BASELINE configs[4]: the controller-only model (box_model.py) - a microbench of the controller CNN + glimpse
  LSTM + attention-box path without the mask head.  Metric: box-steps/sec = B*T / forward time."""
  import torch
  from rec_attend_b200 import config, dist_util, synthetic
  from rec_attend_b200.box_model import BoxModel
  from rec_attend_b200 import _lib
  opt = config.baseline_opt(args.config)
  B, T = args.batch or cfg['B'], cfg['T']
  batch_np = synthetic.make_batch(opt, B, seed=dist_util.rank_seed(1234, args.config, rank))
  model = BoxModel(opt).load_weights(synthetic.make_weights(opt, model='box'))
  dev_batch = {k: torch.from_numpy(v).cuda() for k, v in batch_np.items()}
  fetch = ['loss', 'box_loss', 'conf_loss', 's_out', 'match_box']
  lib = _lib.lib()
  for _ in range(max(3, args.warmup)):
    model.forward(dev_batch, outputs=fetch)
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  dist_util.barrier()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.steps):
    model.forward(dev_batch, outputs=fetch)
  e1.record()
  dist_util.barrier()
  torch.cuda.synchronize()
  ms = dist_util.max_over_ranks(e0.elapsed_time(e1), device='cuda')
  clocks = sampler.stop() if rank == 0 else None
  n0 = lib.ra_launch_count()
  model.forward(dev_batch, outputs=fetch, use_graph=False)
  torch.cuda.synchronize()
  launches = int(lib.ra_launch_count() - n0)
  if rank == 0:
    emit({'metric': 'box-steps/sec', 'value': world * B * T / (ms / args.steps / 1e3), 'unit': 'box-steps/s',
          'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms / args.steps,
          'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
          'config': {'workload': cfg['name'], 'arch': cfg['arch'], 'batch_per_gpu': B, 'timespan': T,
                     'height': cfg['H'], 'width': cfg['W'], 'parallelism': 'dp{}'.format(world),
                     'step': 'eval forward of the controller-only model (T-step loop with GT-driven canvas) + box / '
                             'confidence loss'},
          'gpu_launches': launches * args.steps, 'clocks': clocks})
  dist_util.finalize()
  return 0


_REAL_STDOUT = None


def _quiet_stdout():
  """Everything that libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the single JSON
  line is written to the real stdout by emit()."""
  global _REAL_STDOUT
  if _REAL_STDOUT is None:
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line):
  out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
  out.write(json.dumps(line) + '\n')
  out.flush()


def main():
  _quiet_stdout()
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--config', type=int, default=2, help='index into BASELINE.json configs (default 2: KITTI 256x512 T=20 B=32)')
  ap.add_argument('--batch', type=int, default=0, help='override the per-GPU batch size')
  ap.add_argument('--ref-batch', type=int, default=4, help='batch size of the bounded CPU-baseline sample')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-train-step', action='store_true', help='skip the training-step timing (weak + strong)')
  args = ap.parse_args()
  if args.impl == 'reference':
    return run_reference(args)
  return run_ours(args)


if __name__ == '__main__':
  sys.exit(main())
