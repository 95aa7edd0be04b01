"""Per-CTA pipeline timeline (clock64 stamps) of single tcgen05 conv launches."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rec_attend_b200 import ops, _lib

LAYERS = {  # the KITTI-arch layers at the bench batch: (B, H, W, C1, C2, Cout, up, pool)
    'ctrl_L1': (32, 128, 256, 16, 0, 16, 1, 2), 'ctrl_L2': (32, 64, 128, 16, 0, 32, 1, 1),
    'ctrl_L3': (32, 64, 128, 32, 0, 32, 1, 2), 'ctrl_L4': (32, 32, 64, 32, 0, 64, 1, 1),
    'ctrl_L5': (32, 32, 64, 64, 0, 64, 1, 2), 'ctrl_L6': (32, 16, 32, 64, 0, 64, 1, 1),
    'ctrl_L7': (32, 16, 32, 64, 0, 64, 1, 2),
    'attn_L0': (32, 48, 48, 16, 0, 16, 1, 1), 'attn_L1': (32, 48, 48, 16, 0, 32, 1, 2),
    'attn_L2': (32, 24, 24, 32, 0, 32, 1, 1), 'attn_L3': (32, 24, 24, 32, 0, 64, 1, 2),
    'attn_L4': (32, 12, 12, 64, 0, 64, 1, 1), 'attn_L5': (32, 12, 12, 64, 0, 96, 1, 2),
    'dcnn_L0': (32, 6, 6, 96, 0, 64, 2, 1), 'dcnn_L1': (32, 12, 12, 64, 64, 64, 1, 1),
    'dcnn_L2': (32, 12, 12, 64, 64, 32, 2, 1), 'dcnn_L3': (32, 24, 24, 32, 32, 32, 1, 1),
    'dcnn_L4': (32, 24, 24, 32, 32, 16, 2, 1), 'dcnn_L5': (32, 48, 48, 16, 16, 16, 1, 1),
    'dcnn_L6': (32, 48, 48, 16, 16, 1, 1, 1), 'tiny': (1, 12, 12, 16, 0, 16, 1, 1)}
MODES = [('tma', None)] if os.environ.get('TL_TMA_ONLY') else [('tma', None), ('plain', '1')]
if len(sys.argv) > 1:
  LAYERS = {k: v for k, v in LAYERS.items() if k in sys.argv[1:]}
total = {m: 0.0 for m, _ in MODES}
names = ['start', 'setup', 'stage0', 'prod_done', 'mma_first', 'mma_issued', 'acc0_full', 'done']
dbg = torch.zeros(148 * 16, dtype=torch.int64, device='cuda')
for name, (B, H, W, C1, C2, Cout, up, pool) in LAYERS.items():
  info = ops.umma_plan_info(C1 + C2, Cout, H * up, W * up, pool, B)
  rng = np.random.default_rng(0)
  x1 = torch.from_numpy(rng.standard_normal((B, H, W, C1)).astype(np.float32)).cuda()
  x2 = torch.from_numpy(rng.standard_normal((B, H, W, C2)).astype(np.float32)).cuda() if C2 else None
  w = rng.standard_normal((3, 3, C1 + C2, Cout)).astype(np.float32)
  wp = ops.umma_filter_image(w, info['KC'], info['NPc'], info['n_split'], info['rowstack'], 'cuda')
  sc = torch.ones(Cout, device='cuda'); sh = torch.zeros(Cout, device='cuda')
  for mode, env in MODES:
   if env is None:
     os.environ.pop('RA_CONV_NO_TMA', None)
   else:
     os.environ['RA_CONV_NO_TMA'] = env
   for _ in range(3):
     out = ops.conv3x3_block_umma(x1, wp, Cout, sc, sh, pool=pool, x2=x2, upsample=up)
   torch.cuda.synchronize()
   _lib.call('ra_debug_conv_timeline', ctypes.c_void_p(dbg.data_ptr()))
   dbg.zero_()
   torch.cuda._sleep(150000)
   e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
   e0.record()
   ops.conv3x3_block_umma(x1, wp, Cout, sc, sh, pool=pool, x2=x2, upsample=up, out=out)
   e1.record(); torch.cuda.synchronize()
   _lib.call('ra_debug_conv_timeline', ctypes.c_void_p(0))
   d16 = dbg.cpu().numpy().reshape(148, 16)[:info['grid']]
   d = d16[:, :8]
   rel = (d - d[:, :1]).astype(np.float64)
   total[mode] += e0.elapsed_time(e1) * 1e3
   print('%s [%s]: %.1f us, plan %s' % (name, mode, e0.elapsed_time(e1) * 1e3, {k: info[k] for k in ('KC', 'TH', 'TW', 'n_split', 'n_mt', 'n_chunks', 'stages', 'w_resident', 'grid', 'rowstack', 'nbuf')}))
   print('   median cycles since CTA start: ' + ', '.join('%s=%d' % (n, np.median(rel[:, i])) for i, n in enumerate(names)))
   print('   max   cycles since CTA start: ' + ', '.join('%s=%d' % (n, rel[:, i].max()) for i, n in enumerate(names)))
   print('   CTA start spread (cycles): %d' % (d[:, 0].max() - d[:, 0].min()))
   wn = ['tma_wait_free_stage', 'conv_wait_box', 'mma_wait_chunk', 'mma_wait_acc', '-', '-', 'conv_busy']
   print('   median wait / busy cycles: ' + ', '.join('%s=%d' % (n, np.median(d16[:, 8 + i])) for i, n in enumerate(wn)))

print('sum over layers (us):', total)
