"""A/B of the operand format of the tcgen05 conv layers: 3xTF32 (mode 0) against the fp16 hi / lo split (modes 1 / 2,
ra_conv3x3_umma_set_f16).  (1) single layers at the bench batch: time and error against an fp64 convolution;
(2) the KITTI eval forward at B = 32: ms per step, outputs against mode 0; (3) B = 2 against the CPU oracle.
Run on the GPU box: python tools/f16_ab.py [layers] [model] [oracle]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rec_attend_b200 as ra
from rec_attend_b200 import ops, _lib

what = set(sys.argv[1:]) or {'layers', 'model', 'oracle'}
MODES = [int(m) for m in os.environ.get('F16_MODES', '0,1,2').split(',')]

LAYERS = {  # name: (B, H, W, C1, C2, Cout, up, pool)
    'ctrl_L1': (32, 128, 256, 16, 0, 16, 1, 2),
    'ctrl_L2': (32, 64, 128, 16, 0, 32, 1, 1),
    'ctrl_L3': (32, 64, 128, 32, 0, 32, 1, 2),
    'ctrl_L4': (32, 32, 64, 32, 0, 64, 1, 1),
    'ctrl_L5': (32, 32, 64, 64, 0, 64, 1, 2),
    'ctrl_L6': (32, 16, 32, 64, 0, 64, 1, 1),
    'ctrl_L7': (32, 16, 32, 64, 0, 64, 1, 2),
    'attn_L1': (32, 48, 48, 16, 0, 32, 1, 2),
    'dcnn_L1': (32, 12, 12, 64, 64, 64, 1, 1),
    'dcnn_L4': (32, 24, 24, 32, 32, 16, 2, 1),
    'dcnn_L5': (32, 48, 48, 16, 16, 16, 1, 1),
}


def layer_ab():
  flush = torch.empty(64 << 20, device='cuda')
  for name, (B, H, W, C1, C2, Cout, up, pool) in LAYERS.items():
    rng = np.random.default_rng(1)
    # activations with a wide dynamic range (post-ReLU maps hold many small values): the fp16 parts must keep them
    x1n = (np.abs(rng.standard_normal((B, H, W, C1))) * np.exp(rng.uniform(-9, 1, (B, H, W, C1)))).astype(np.float32)
    x2n = rng.standard_normal((B, H, W, C2)).astype(np.float32) if C2 else None
    w = (rng.standard_normal((3, 3, C1 + C2, Cout)) / np.sqrt(9 * (C1 + C2))).astype(np.float32)
    x1 = torch.from_numpy(x1n).cuda()
    x2 = torch.from_numpy(x2n).cuda() if C2 else None
    sc = torch.ones(Cout, device='cuda'); sh = torch.zeros(Cout, device='cuda')
    ref = None
    if up == 1 and pool == 1:  # fp64 reference on the device (first 2 examples)
      xin = x1[:2] if x2 is None else torch.cat([x1[:2], x2[:2]], 3)
      ref = torch.nn.functional.conv2d(xin.double().permute(0, 3, 1, 2), torch.from_numpy(w).cuda().double().permute(3, 2, 0, 1),
                                       padding=1).permute(0, 2, 3, 1).relu()
    line = []
    for mode in MODES:
      ops.umma_set_f16(mode)
      info = ops.umma_plan_info(C1 + C2, Cout, H * up, W * up, pool, B)
      wp = ops.umma_filter_image(w, info['KC'], info['NPc'], info['n_split'], info['rowstack'], 'cuda')
      out = ops.conv3x3_block_umma(x1, wp, Cout, sc, sh, pool=pool, x2=x2, upsample=up)
      ts = []
      for _ in range(5):
        flush.zero_()
        torch.cuda._sleep(150000)  # queue the launch behind a busy GPU: the events see the kernel, not the host call
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv3x3_block_umma(x1, wp, Cout, sc, sh, pool=pool, x2=x2, upsample=up, out=out)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
      err = ''
      if ref is not None:
        err = ' err %.1e' % float((out[:2].double() - ref).abs().max() / ref.abs().max())
      line.append('mode %d: %6.1f us KC=%d flags=%d TH=%d TW=%d%s' % (mode, min(ts), info['KC'], info['rowstack'], info['TH'],
                                                                     info['TW'], err))
    print('%-8s %s' % (name, ' | '.join(line)), flush=True)
  ops.umma_set_f16(0)


def model_ab():
  from rec_attend_b200.full_model import FullModel
  opt = ra.config.baseline_opt(2)
  B = int(os.environ.get('F16_B', '32'))
  batch = {k: torch.from_numpy(v).cuda() for k, v in ra.synthetic.make_batch(opt, B, seed=1234).items()}
  w = ra.synthetic.make_weights(opt)
  keys = ['y_out', 's_out', 'attn_box', 'match', 'loss', 'iou_soft']
  base = None
  for mode in MODES:
    ops.umma_set_f16(mode)
    model = FullModel(opt).load_weights(w)
    for _ in range(4):
      out = model.forward(batch, outputs=keys)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
      out = model.forward(batch, outputs=keys)
    e1.record(); torch.cuda.synchronize()
    res = {k: out[k].detach().float().cpu().numpy().copy() for k in keys}
    msg = 'mode %d: %.3f ms per eval forward (KITTI 256x512 T=20 B=%d), loss %.7f' % (mode, e0.elapsed_time(e1) / 10, B,
                                                                                    float(res['loss']))
    if base is None:
      base = res
    else:
      msg += ' | vs mode 0: ' + ', '.join('%s %.1e' % (k, float(np.abs(res[k] - base[k]).max())) for k in keys)
    print(msg, flush=True)
    del model
  ops.umma_set_f16(0)


def oracle_ab():
  from rec_attend_b200.full_model import FullModel
  from oracle import model as OM
  opt = ra.config.baseline_opt(2)
  T = int(opt['timespan'])
  batch = ra.synthetic.make_batch(opt, 2, seed=1234)
  w = ra.synthetic.make_weights(opt, seed=4321)
  with torch.no_grad():
    ref = OM.full_model_forward(opt, w, batch)
  for mode in MODES:
    ops.umma_set_f16(mode)
    out = FullModel(opt).load_weights(w).forward(batch)
    torch.cuda.synchronize()
    line = []
    for k in ('ctrl_out', 'attn_box', 'y_out', 's_out', 'canvas'):
      if k in ref and k in out:
        a, b = out[k].float().cpu().numpy(), ref[k].numpy()
        line.append('%s %.1e' % (k, float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))))
    same = bool((out['match'].cpu().numpy() == ref['match'].numpy()).all())
    print('mode %d vs the oracle (B=2): %s | match equal: %s' % (mode, ', '.join(line), same), flush=True)
  ops.umma_set_f16(0)


if 'layers' in what:
  layer_ab()
if 'model' in what:
  model_ab()
if 'oracle' in what:
  oracle_ab()
