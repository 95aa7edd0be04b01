"""Where does the ~28 us fixed cost of a tcgen05 conv launch come from?  Times (a) isolated launches after an
L2 flush kernel, (b) back-to-back launches of the same layer, (c) alternating two layers with different smem."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rec_attend_b200 import ops

def make(shape):
  B, H, W, C1, Cout, pool = shape
  info = ops.umma_plan_info(C1, Cout, H, W, pool, B)
  rng = np.random.default_rng(0)
  x = torch.from_numpy(rng.standard_normal((B, H, W, C1)).astype(np.float32)).cuda()
  w = rng.standard_normal((3, 3, C1, Cout)).astype(np.float32)
  wp = ops.umma_filter_image(w, info['KC'], info['NPc'], info['n_split'], info['rowstack'], 'cuda')
  sc = torch.ones(Cout, device='cuda'); sh = torch.zeros(Cout, device='cuda')
  out = ops.conv3x3_block_umma(x, wp, Cout, sc, sh, pool=pool)
  return (x, wp, Cout, sc, sh, pool, out), info

def run(a):
  x, wp, Cout, sc, sh, pool, out = a
  ops.conv3x3_block_umma(x, wp, Cout, sc, sh, pool=pool, out=out)

def timed(fn, n):
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) * 1e3 / n

a, ia = make((32, 48, 48, 16, 32, 2))
b, ib = make((32, 12, 12, 64, 96, 2))
c, ic = make((1, 12, 12, 16, 16, 1))
print('A smem', ia['smem_bytes'], 'B smem', ib['smem_bytes'], 'C smem', ic['smem_bytes'], 'C grid', ic['grid'])
small = torch.zeros(1024, device='cuda')
for name, fn in (('A back-to-back', lambda: run(a)), ('B back-to-back', lambda: run(b)), ('C (tiny, 1 image) back-to-back', lambda: run(c)),
                 ('A,B alternating (per launch)', lambda: (run(a), run(b))), ('A + small torch kernel', lambda: (run(a), small.add_(1.0))),
                 ('small torch kernel alone', lambda: small.add_(1.0))):
  for _ in range(3): fn()
  t = timed(fn, 200)
  print('%-40s %.1f us per iteration' % (name, t))
