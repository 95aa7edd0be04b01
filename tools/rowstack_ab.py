"""A/B of the row-stacked-taps mode of the tcgen05 conv (RA_UMMA_ROWSTACK=0|1|2, read once per process): per KITTI
layer the plan, the time (L2 flushed, best of 5) and the max error against the CUDA-core fp32 convolution.
  for m in 0 1 2; do RA_UMMA_ROWSTACK=$m python tools/rowstack_ab.py; done"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rec_attend_b200 import ops

LAYERS = {  # name: (B, H, W, C1, C2, Cout, up, pool)
    'ctrl_L1': (32, 128, 256, 16, 0, 16, 1, 2), 'ctrl_L2': (32, 64, 128, 16, 0, 32, 1, 1),
    'ctrl_L3': (32, 64, 128, 32, 0, 32, 1, 2), 'ctrl_L4': (32, 32, 64, 32, 0, 64, 1, 1),
    'attn_L0': (32, 48, 48, 16, 0, 16, 1, 1), 'attn_L1': (32, 48, 48, 16, 0, 32, 1, 2),
    'attn_L2': (32, 24, 24, 32, 0, 32, 1, 1), 'dcnn_L2': (32, 12, 12, 64, 64, 32, 2, 1),
    'dcnn_L3': (32, 24, 24, 32, 32, 32, 1, 1), 'dcnn_L4': (32, 24, 24, 32, 32, 16, 2, 1),
    'dcnn_L5': (32, 48, 48, 16, 16, 16, 1, 1), 'dcnn_L6': (32, 48, 48, 16, 16, 1, 1, 1),
    'odd': (3, 10, 14, 12, 0, 20, 1, 1), 'small_pool': (2, 8, 12, 8, 0, 8, 1, 2)}
mode = os.environ.get('RA_UMMA_ROWSTACK', '1')
tot = 0.0
for name, (B, H, W, C1, C2, Cout, up, pool) in LAYERS.items():
  info = ops.umma_plan_info(C1 + C2, Cout, H * up, W * up, pool, B)
  rng = np.random.default_rng(1)
  x1 = torch.from_numpy(rng.standard_normal((B, H, W, C1)).astype(np.float32)).cuda()
  x2 = torch.from_numpy(rng.standard_normal((B, H, W, C2)).astype(np.float32)).cuda() if C2 else None
  w = (rng.standard_normal((3, 3, C1 + C2, Cout)) / np.sqrt(9 * (C1 + C2))).astype(np.float32)
  wp = ops.umma_filter_image(w, info['KC'], info['NPc'], info['n_split'], info['rowstack'], 'cuda')
  sc = torch.from_numpy(rng.uniform(0.5, 1.5, Cout).astype(np.float32)).cuda()
  sh = torch.from_numpy(rng.standard_normal(Cout).astype(np.float32)).cuda()
  out = ops.conv3x3_block_umma(x1, wp, Cout, sc, sh, pool=pool, x2=x2, upsample=up)
  ref = ops.conv3x3_block(x1, torch.from_numpy(w).cuda(), sc, sh, pool=pool, x2=x2, upsample=up)
  err = float((out - ref).abs().max() / ref.abs().max())
  flush = torch.empty(64 << 20, device='cuda')
  ts = []
  for _ in range(5):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv3x3_block_umma(x1, wp, Cout, sc, sh, pool=pool, x2=x2, upsample=up, out=out)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
  tot += min(ts)
  print('mode %s %-10s %6.1f us  err %.1e  %s' % (mode, name, min(ts), err, {k: info[k] for k in (
      'KC', 'NPc', 'TH', 'TW', 'n_mt', 'rowstack', 'ksplit', 'nbuf', 'stages', 'w_resident', 'grid')}), flush=True)
print('mode', mode, 'sum', round(tot, 1), 'us')
