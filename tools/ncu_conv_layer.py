"""One tcgen05 conv layer at the bench batch, a few launches - the target of an ncu --set full capture.
python tools/ncu_conv_layer.py [layer] (names of tools/f16_ab.py)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rec_attend_b200 import ops

LAYERS = {'ctrl_L1': (32, 128, 256, 16, 0, 16, 1, 2), 'ctrl_L2': (32, 64, 128, 16, 0, 32, 1, 1),
          'ctrl_L3': (32, 64, 128, 32, 0, 32, 1, 2), 'ctrl_L5': (32, 32, 64, 64, 0, 64, 1, 2),
          'dcnn_L5': (32, 48, 48, 16, 16, 16, 1, 1)}
name = sys.argv[1] if len(sys.argv) > 1 else 'ctrl_L1'
B, H, W, C1, C2, Cout, up, pool = LAYERS[name]
g = torch.Generator(device='cuda').manual_seed(0)
x1 = torch.randn((B, H, W, C1), device='cuda', generator=g).abs()
x2 = torch.randn((B, H, W, C2), device='cuda', generator=g) if C2 else None
w = np.random.default_rng(0).standard_normal((3, 3, C1 + C2, Cout)).astype(np.float32)
KC, NPc, nsp, nch, flags = ops.umma_plan(C1 + C2, Cout, H * up, W * up, pool, B, C2=C2)
wp = ops.umma_filter_image(w, KC, NPc, nsp, flags, 'cuda')
sc = torch.ones(Cout, device='cuda'); sh = torch.zeros(Cout, device='cuda')
for _ in range(4):
  out = ops.conv3x3_block_umma(x1, wp, Cout, sc, sh, pool=pool, x2=x2, upsample=up)
torch.cuda.synchronize()
print(name, 'plan', KC, NPc, nsp, nch, flags)
