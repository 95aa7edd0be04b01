"""Calibration of the tcgen05 conv tile planner: times single layers under forced plans
(RA_UMMA_FORCE="KC,TH,TW,n_split,resident").  Run on the GPU box: python tools/bench_conv_layers.py"""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rec_attend_b200 import ops, _lib

LAYERS = {  # name: (B, H, W, C1, C2, Cout, up, pool)
    'ctrl_L1': (32, 128, 256, 16, 0, 16, 1, 2),
    'ctrl_L3': (32, 64, 128, 32, 0, 32, 1, 2),
    'ctrl_L5': (32, 32, 64, 64, 0, 64, 1, 2),
    'ctrl_L7': (32, 16, 32, 64, 0, 64, 1, 2),
    'attn_L1': (32, 48, 48, 16, 0, 32, 1, 2),
    'attn_L5': (32, 12, 12, 64, 0, 96, 1, 2),
    'dcnn_L2': (32, 12, 12, 64, 64, 32, 2, 1),
    'dcnn_L4': (32, 24, 24, 32, 32, 16, 2, 1),
}


def time_layer(shape, force):
  B, H, W, C1, C2, Cout, up, pool = shape
  if force is None:
    os.environ.pop('RA_UMMA_FORCE', None)
  else:
    os.environ['RA_UMMA_FORCE'] = force
  try:
    info = ops.umma_plan_info(C1 + C2, Cout, H * up, W * up, pool, B)
  except _lib.RecAttendError:
    return None, None
  rng = np.random.default_rng(0)
  x1 = torch.from_numpy(rng.standard_normal((B, H, W, C1)).astype(np.float32)).cuda()
  x2 = torch.from_numpy(rng.standard_normal((B, H, W, C2)).astype(np.float32)).cuda() if C2 else None
  w = rng.standard_normal((3, 3, C1 + C2, Cout)).astype(np.float32)
  wp = torch.from_numpy(ops.pack_umma_weights(w, info['KC'], info['NPc'], info['n_split'], info['rowstack'])).cuda()
  sc = torch.ones(Cout, device='cuda'); sh = torch.zeros(Cout, device='cuda')
  out = ops.conv3x3_block_umma(x1, wp, Cout, sc, sh, pool=pool, x2=x2, upsample=up)
  flush = torch.empty(64 << 20, device='cuda')
  ts = []
  for _ in range(5):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv3x3_block_umma(x1, wp, Cout, sc, sh, pool=pool, x2=x2, upsample=up, out=out)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
  return min(ts), info


for name, shape in LAYERS.items():
  B, H, W, C1, C2, Cout, up, pool = shape
  Ho, Wo = H * up, W * up
  t, info = time_layer(shape, None)
  print('%s auto: %.1f us  %s' % (name, t, {k: info[k] for k in ('KC', 'TH', 'TW', 'n_split', 'w_resident', 'n_mt', 'stages', 'grid')}))
  res = []
  tws = sorted({Wo, Wo // 2, Wo // 4, 32, 16} & {w for w in (Wo, Wo // 2, Wo // 4, 32, 16, 64) if 2 <= w <= Wo and w % 2 == 0})
  for KC, TW, nsp, resd in itertools.product((8, 16, 32), tws, (1, 2, 4), (0, 1)):
    for TH in (2, 4, 6, 8, 12, 16, 24, 32):
      if TH > Ho:
        continue
      t, info = time_layer(shape, '%d,%d,%d,%d,%d' % (KC, TH, TW, nsp, resd))
      if t is not None:
        res.append((t, KC, TH, TW, nsp, resd, info['n_mt'], info['stages'], info['grid']))
  res.sort()
  for r in res[:6]:
    print('   %.1f us KC=%d TH=%d TW=%d nsplit=%d res=%d n_mt=%d stages=%d grid=%d' % r)
  print('   ... worst %.1f us, %d configs' % (res[-1][0], len(res)))
