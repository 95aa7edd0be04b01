"""Secondary bench line for the foreground / orientation FCN (fg_model.py) — NOT the headline bench.

  python tools/bench_fg.py [--arch kitti|cityscapes|default] [--height 256 --width 512] [--batch 8] [--steps 10]

One step = FgModel.forward on a device-resident batch (29-33 fused conv launches + the head / loss pass), timed with
CUDA events after >= 3 warm-up steps.  Prints one JSON line: images/s, the conv stack's algorithmic TFLOP/s against the
measured bf16 tensor peak (MEASURED_PEAKS.json; the arithmetic is 3xTF32, so 1/6 of that peak is the ceiling of this
formulation), which layers ran on the tcgen05 kernel, and the CPU oracle timed on one image beside it.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def conv_flops_per_image(opt):
  """2 * 9 * Cin * Cout * output pixels of every CNN / DCNN layer (the reference's dense formulation)."""
  from rec_attend_b200.config import fg_skip_wiring
  H, W = opt['inp_height'], opt['inp_width']
  ch = [opt['inp_depth']] + list(opt['cnn_depth'])
  tot, h, w = 0.0, H, W
  for i, p in enumerate(opt['cnn_pool']):
    tot += 2.0 * 9 * ch[i] * ch[i + 1] * h * w
    h, w = h // p, w // p
  _, sch = fg_skip_wiring(opt)
  dch = [ch[-1]] + list(opt['dcnn_depth'])
  for i, p in enumerate(opt['dcnn_pool']):
    h, w = h * p, w * p
    # transposed conv with stride p: each INPUT pixel meets 9 taps -> 9 * Cin * Cout * (output pixels / p^2) MACs
    tot += 2.0 * 9 * (dch[i] + sch[i]) * dch[i + 1] * h * w / float(p * p)
  return tot


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--arch', default='kitti')
  ap.add_argument('--height', type=int, default=256)
  ap.add_argument('--width', type=int, default=512)
  ap.add_argument('--batch', type=int, default=8)  # run_kitti.sh:23 --batch_size 8
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--no-cpu', action='store_true')
  args = ap.parse_args()
  import torch
  import rec_attend_b200 as ra
  from rec_attend_b200 import _lib
  from rec_attend_b200.fg_model import FgModel
  opt = ra.config.fg_model_opt(args.arch, args.height, args.width)
  weights = ra.synthetic.make_fg_weights(opt)
  batch = ra.synthetic.make_fg_batch(opt, args.batch)
  model = FgModel(opt).load_weights(weights)
  dev = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
  for _ in range(max(3, args.warmup)):
    model.forward(dev)
  torch.cuda.synchronize()
  lib = _lib.lib()
  n0 = lib.ra_launch_count()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.steps):
    out = model.forward(dev)
  e1.record()
  torch.cuda.synchronize()
  launches = int(lib.ra_launch_count() - n0)
  ms = e0.elapsed_time(e1) / args.steps
  flops = conv_flops_per_image(opt) * args.batch
  peak = None
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      pk = json.load(f)
    peak = float(pk.get('bf16_tflops_sustained') or pk.get('bf16_tflops') or 0) or None
  except Exception:
    peak = None
  kinds = model.conv_kernels(args.batch)
  line = {
      'metric': 'fg-images/sec', 'value': args.batch / (ms / 1e3), 'unit': 'images/s', 'ms_per_step': ms,
      'steps': args.steps, 'warmup': max(3, args.warmup), 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': 'fg_{}_{}x{}_B{}'.format(args.arch, args.height, args.width, args.batch),
                 'step': 'eval forward of the foreground / orientation FCN + head / loss pass, device-resident inputs'},
      'conv_gflop_per_image': conv_flops_per_image(opt) / 1e9,
      'roofline': {'bound': 'tensor', 'achieved': flops / (ms / 1e3) / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                   'frac': (flops / (ms / 1e3) / 1e12 / peak) if peak else None,
                   'note': 'whole step (convs + head) in the denominator; arithmetic is 3xTF32'},
      'layers_on_tcgen05': kinds.count('umma'), 'layers_on_fp32_kernel': kinds.count('fp32'),
      'gpu_launches': launches, 'loss': float(out['loss']),
  }
  if not args.no_cpu:
    from oracle import model as OM
    one = {k: v[:1] for k, v in batch.items()}
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
      OM.fg_model_forward(opt, weights, one)
      t0 = time.perf_counter()
      OM.fg_model_forward(opt, weights, one)
      dt = time.perf_counter() - t0
    line['cpu_baseline'] = {'value': 1.0 / dt, 'unit': 'images/s', 'cores': os.cpu_count(), 'kind': 'port',
                            'sample': 'oracle (PyTorch-CPU restatement), 1 image'}
  print(json.dumps(line))


if __name__ == '__main__':
  main()
