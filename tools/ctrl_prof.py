"""Phase timeline of controller_cluster_kernel (build csrc with NVCCFLAGS += -DRA_CTRL_PROF first; debugging aid only).
CTA 0 / thread 0 stamps %globaltimer at every phase boundary of the last launch of an eager KITTI forward."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from rec_attend_b200 import _lib, config, synthetic  # noqa: E402
from rec_attend_b200.full_model import FullModel  # noqa: E402

opt = config.baseline_opt(2)
B = int(os.environ.get('B', config.BASELINE_CONFIGS[2]['B']))
batch = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_batch(opt, B).items()}
model = FullModel(opt).load_weights(synthetic.make_weights(opt))
for _ in range(2):
  model.forward(batch, use_graph=False)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
out = (ctypes.c_ulonglong * 64)()
assert lib.ra_debug_ctrl_prof(out) == 0
t = np.array(list(out), np.int64)
names = {0: 'entry', 1: 'weight copies issued (cp.async)', 2: 'griddepcontrol.wait + copies landed + cluster.sync', 40: 'head done', 41: 'final cluster.sync'}
for it in range(5):
  names[3 + it * 5] = 'it%d read-out + sync' % it
  names[4 + it * 5] = 'it%d LSTM gates + sync' % it
  names[5 + it * 5] = 'it%d glimpse MLP0 + sync' % it
  names[6 + it * 5] = 'it%d glimpse MLP1 logits + sync' % it
  names[7 + it * 5] = 'it%d softmax' % it
print('it1 LSTM detail: loop %.2f | shuffle+store %.2f | syncthreads %.2f | gates %.2f | remote stores %.2f | cluster.sync %.2f us' % tuple((b - a) / 1e3 for a, b in [(t[8], t[42]), (t[42], t[43]), (t[43], t[44]), (t[44], t[45]), (t[45], t[46]), (t[46], t[9])]))
prev = t[0]
for i in sorted(names):
  if t[i] == 0:
    continue
  print('%-36s +%7.2f us   (t = %7.2f us)' % (names[i], (t[i] - prev) / 1e3, (t[i] - t[0]) / 1e3))
  prev = t[i]
