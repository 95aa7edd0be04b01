"""Phase timeline of paste_back_kernel (build csrc with NVCCFLAGS += -DRA_PB_PROF; debugging aid only): thread 0 of every
CTA stamps %globaltimer; the last launch of an eager KITTI forward is analysed: inside-box tiles vs constant tiles."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from rec_attend_b200 import _lib, config, synthetic  # noqa: E402
from rec_attend_b200.full_model import FullModel  # noqa: E402

opt = config.baseline_opt(2)
B = config.BASELINE_CONFIGS[2]['B']
batch = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_batch(opt, B).items()}
model = FullModel(opt).load_weights(synthetic.make_weights(opt))
model.forward(batch, use_graph=False)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
lib.ra_debug_pb_prof_clear()
model.forward(batch, use_graph=False)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * (4096 * 8))()
assert lib.ra_debug_pb_prof(out) == 0
t = np.array(list(out), np.int64).reshape(4096, 8)
t0 = t[:, 0][t[:, 0] > 0].min()
inside = t[:, 6] > 0
outside = (t[:, 7] > 0) & ~inside
print('CTAs: %d inside the box, %d constant tiles (last launch, B=%d)' % (inside.sum(), outside.sum(), B))
print('kernel span (first CTA start -> last CTA end): %.2f us' % ((max(t[:, 6].max(), t[:, 7].max()) - t0) / 1e3))
ti = t[inside]
names = ['griddepcontrol.wait', 'band / tap range', 'load P, wy', 'Fy^T P (t2)', 'band loop over taps', 'sigmoid + stores']
for k, n in enumerate(names):
  d = (ti[:, k + 1] - ti[:, k]) / 1e3
  print('  inside  %-22s mean %6.2f us  max %6.2f' % (n, d.mean(), d.max()))
print('  inside  total after wait       mean %6.2f us' % ((ti[:, 6] - ti[:, 1]).mean() / 1e3))
print('  inside  start offsets: mean %.2f us, last start %.2f us' % ((ti[:, 0] - t0).mean() / 1e3, (ti[:, 0] - t0).max() / 1e3))
to = t[outside]
print('  outside total after wait       mean %6.2f us' % ((to[:, 7] - to[:, 1]).mean() / 1e3))
print('  outside start offsets: mean %.2f us, last start %.2f us' % ((to[:, 0] - t0).mean() / 1e3, (to[:, 0] - t0).max() / 1e3))
