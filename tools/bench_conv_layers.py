"""Calibration of the tcgen05 conv tile planner: times single layers under forced plans
(RA_UMMA_FORCE="KC,TH,TW,n_split,resident").  Run on the GPU box: python tools/bench_conv_layers.py"""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rec_attend_b200 import ops, _lib

LAYERS = {  # name: (B, H, W, C1, C2, Cout, up, pool)
    'ctrl_L1': (32, 128, 256, 16, 0, 16, 1, 2),
    'ctrl_L2': (32, 64, 128, 16, 0, 32, 1, 1),
    'ctrl_L3': (32, 64, 128, 32, 0, 32, 1, 2),
    'ctrl_L4': (32, 32, 64, 32, 0, 64, 1, 1),
    'ctrl_L5': (32, 32, 64, 64, 0, 64, 1, 2),
    'ctrl_L6': (32, 16, 32, 64, 0, 64, 1, 1),
    'ctrl_L7': (32, 16, 32, 64, 0, 64, 1, 2),
    'attn_L0': (32, 48, 48, 16, 0, 16, 1, 1), 'attn_L1': (32, 48, 48, 16, 0, 32, 1, 2),
    'attn_L2': (32, 24, 24, 32, 0, 32, 1, 1), 'attn_L3': (32, 24, 24, 32, 0, 64, 1, 2),
    'attn_L4': (32, 12, 12, 64, 0, 64, 1, 1), 'attn_L5': (32, 12, 12, 64, 0, 96, 1, 2),
    'dcnn_L0': (32, 6, 6, 96, 0, 64, 2, 1), 'dcnn_L1': (32, 12, 12, 64, 64, 64, 1, 1),
    'dcnn_L2': (32, 12, 12, 64, 64, 32, 2, 1), 'dcnn_L3': (32, 24, 24, 32, 32, 32, 1, 1),
    'dcnn_L4': (32, 24, 24, 32, 32, 16, 2, 1), 'dcnn_L5': (32, 48, 48, 16, 16, 16, 1, 1),
    'dcnn_L6': (32, 48, 48, 16, 16, 1, 1, 1),
}


_inputs = {}


def time_layer(shape, force):
  B, H, W, C1, C2, Cout, up, pool = shape
  if force is None:
    os.environ.pop('RA_UMMA_FORCE', None)
  else:
    os.environ['RA_UMMA_FORCE'] = force
  try:
    info = ops.umma_plan_info(C1 + C2, Cout, H * up, W * up, pool, B)  # (C2 > 0 layers here have C1 % 16 == 0: same plan)
  except _lib.RecAttendError:
    return None, None
  if shape not in _inputs:  # inputs made once per layer (on the device: the sweep is hundreds of plans)
    g = torch.Generator(device='cuda').manual_seed(0)
    _inputs.clear()
    _inputs[shape] = (torch.randn((B, H, W, C1), device='cuda', generator=g),
                      torch.randn((B, H, W, C2), device='cuda', generator=g) if C2 else None,
                      np.random.default_rng(0).standard_normal((3, 3, C1 + C2, Cout)).astype(np.float32),
                      torch.empty(64 << 20, device='cuda'))
  x1, x2, w, flush = _inputs[shape]
  wp = ops.umma_filter_image(w, info['KC'], info['NPc'], info['n_split'], info['rowstack'], 'cuda')
  sc = torch.ones(Cout, device='cuda'); sh = torch.zeros(Cout, device='cuda')
  try:
    out = ops.conv3x3_block_umma(x1, wp, Cout, sc, sh, pool=pool, x2=x2, upsample=up)
  except _lib.RecAttendError:
    return None, None  # e.g. an fp16 plan whose tile the TMA feed cannot serve
  ts = []
  for _ in range(5):
    flush.zero_()
    torch.cuda._sleep(150000)  # the launch is queued behind a busy GPU: the events see the kernel, not the host call
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv3x3_block_umma(x1, wp, Cout, sc, sh, pool=pool, x2=x2, upsample=up, out=out)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
  return min(ts), info


# BCL_LAYERS=ctrl_L1,ctrl_L3 restricts the sweep; BCL_F16=0/1/2 sets the operand format (default: the library's)
if os.environ.get('BCL_F16'):
  ops.umma_set_f16(int(os.environ['BCL_F16']))
KCS = tuple(int(k) for k in os.environ.get('BCL_KC', '8,16,32').split(','))
only = [n for n in os.environ.get('BCL_LAYERS', '').split(',') if n]
for name, shape in LAYERS.items():
  if only and name not in only:
    continue
  B, H, W, C1, C2, Cout, up, pool = shape
  Ho, Wo = H * up, W * up
  t, info = time_layer(shape, None)
  print('%s auto: %.1f us  %s' % (name, t, {k: info[k] for k in ('KC', 'TH', 'TW', 'n_split', 'w_resident', 'n_mt', 'stages', 'grid', 'rowstack')}), flush=True)
  res = []
  tws = sorted({Wo, Wo // 2, Wo // 4, 32, 16} & {w for w in (Wo, Wo // 2, Wo // 4, 32, 16, 64) if 2 <= w <= Wo and w % 2 == 0})
  for KC, TW, nsp, resd in itertools.product(KCS, tws, (1, 2, 4), (0, 1)):
    for TH in (2, 4, 6, 8, 12, 16, 24, 32):
      if TH > Ho:
        continue
      t, info = time_layer(shape, '%d,%d,%d,%d,%d' % (KC, TH, TW, nsp, resd))
      if t is not None:
        res.append((t, KC, TH, TW, nsp, resd, info['n_mt'], info['stages'], info['grid']))
  res.sort()
  for r in res[:8]:
    print('   %.1f us KC=%d TH=%d TW=%d nsplit=%d res=%d n_mt=%d stages=%d grid=%d' % r)
  print('   ... worst %.1f us, %d configs' % (res[-1][0], len(res)))
