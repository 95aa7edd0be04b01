"""Where does the full-resolution parity error come from?  Per decode step: max |cuda - oracle| of the controller
output, the box parameters, attn_box and y_out (KITTI arch 256x512, T=20).  python tools/dbg_parity.py [arch H W T B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rec_attend_b200 as ra
from rec_attend_b200.full_model import FullModel
from oracle import model as OM

arch, H, W, T, B = (sys.argv[1:] + ['kitti', 256, 512, 20, 2])[:5] if len(sys.argv) > 1 else ('kitti', 256, 512, 20, 2)
H, W, T, B = int(H), int(W), int(T), int(B)
opt = ra.config.full_model_opt(arch, H, W, T)
batch = ra.synthetic.make_batch(opt, B, seed=1234)
w = ra.synthetic.make_weights(opt, seed=4321)
train = bool(os.environ.get('RA_DBG_TRAIN'))  # training-mode forward (batch-statistics BN)
with torch.no_grad():
  ref = OM.full_model_forward(opt, w, batch, phase_train=train)
model = FullModel(opt).load_weights(w)
out = model.forward(batch, phase_train=train)
torch.cuda.synchronize()
if train:
  new_w = model.export_weights()
  errs = sorted(((float(np.abs(new_w[k] - v.numpy()).max() / max(1e-12, np.abs(v.numpy()).max())), k)
                 for k, v in ref['ema_updates'].items()), reverse=True)
  print('worst EMA updates:', errs[:6])
print('train' if train else 'eval', 'mode', 'fp32 CUDA-core convs' if os.environ.get('RA_CONV_FP32') else 'tcgen05 3xTF32 convs', arch, H, W, T, B)
for k in ('ctrl_out', 'attn_ctr', 'attn_size', 'attn_lg_var', 'x_patch', 'y_out_patch', 'attn_box', 'y_out', 's_out'):
  if k not in ref or k not in out:
    continue
  a, b = out[k].float().cpu().numpy(), ref[k].numpy()
  per_t = [float(np.abs(a[:, t] - b[:, t]).max()) for t in range(T)]
  print('{:12s} scale {:9.3g}  max|d| per step: {}'.format(k, float(np.abs(b).max()), ' '.join('%.1e' % v for v in per_t)))
