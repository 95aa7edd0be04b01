"""Generates tests/golden/nnlib_golden.npz by EXECUTING THE REFERENCE'S OWN nnlib.py (unmodified, imported from
/root/reference) - nn.cnn / nn.dcnn / nn.mlp / nn.lstm with batch_norm - over the numpy stand-in of
tests/golden/tf012_shim.  What this pins in oracle/model.py: the layer glue (conv + bias -> BN -> activation -> pool
order, one BN copy per call with the EMA shadows starting at zero, skip concatenation order and the
[kh,kw,Cout,Cin+skip] filter of the deconv head, the LSTM gate equations and state layout, the MLP).  The
convolution / pooling primitives themselves are the shim's reading of TensorFlow's 'SAME' semantics.
Run in the build container:  python tests/golden/make_nnlib_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'


def main():
  sys.path.insert(0, ROOT)
  sys.path.insert(0, REF)
  sys.path.insert(0, os.path.join(HERE, 'tf012_shim'))
  import tensorflow as tf
  import nnlib as nn  # the reference source file itself
  assert os.path.dirname(os.path.abspath(nn.__file__)) == REF, nn.__file__
  rng = np.random.default_rng(4242)
  f32 = np.float32
  out = {}
  B, H, W = 3, 16, 24
  n_copies = 2

  def conv_weights(chs, transposed_skip=None):
    ws = []
    for i in range(len(chs) - 1):
      cin = chs[i] + (transposed_skip[i] if transposed_skip else 0)
      shape = (3, 3, chs[i + 1], cin) if transposed_skip is not None else (3, 3, chs[i], chs[i + 1])
      d = {'w': (rng.standard_normal(shape) / np.sqrt(9 * cin)).astype(f32),
           'b': (rng.standard_normal(chs[i + 1]) * 0.1).astype(f32)}
      for t in range(n_copies):
        d['beta_%d' % t] = (rng.standard_normal(chs[i + 1]) * 0.5).astype(f32)
        d['gamma_%d' % t] = rng.uniform(0.5, 1.5, chs[i + 1]).astype(f32)
      ws.append(d)
    return ws

  def store_w(prefix, ws):
    for i, d in enumerate(ws):
      for k, v in d.items():
        out['%s_%d_%s' % (prefix, i, k)] = v

  # ---- nn.cnn: 3 layers, pools 1,2,2; called twice (BN copy 0 then copy 1), training then eval
  cch, cpool = [4, 6, 8, 8], [1, 2, 2]
  cw = conv_weights(cch)
  store_w('cnn_w', cw)
  x = rng.standard_normal((B, H, W, cch[0])).astype(f32)
  out['cnn_x'] = x
  for phase in (True, False):
    model = {}
    run = nn.cnn([3] * 3, cch, cpool, [tf.nn.relu] * 3, [True] * 3, phase_train=phase, wd=5e-5, scope='net', model=model,
                 init_weights=cw)
    for call in range(n_copies):
      h = run(x if call == 0 else x[:, ::-1].copy())
      for i, hi in enumerate(h):
        out['cnn_%s_call%d_h%d' % ('train' if phase else 'eval', call, i)] = np.asarray(hi)
    if phase:
      for i in range(3):
        for call in range(n_copies):
          for n in ('ema_mean', 'ema_var'):
            out['cnn_train_%d_%d_%s' % (i, call, n)] = np.asarray(model['net_%d_%d_%s' % (i, call, n)])

  # ---- nn.dcnn: 3 layers, unpool 2,1,2, skips on layers 1 and 2; training mode, one call
  dch, dpool, dskip_ch = [8, 6, 6, 2], [2, 1, 2], [0, 5, 3]
  dw = conv_weights(dch, transposed_skip=dskip_ch)
  store_w('dcnn_w', dw)
  xd = rng.standard_normal((B, 4, 6, dch[0])).astype(f32)
  sk1 = rng.standard_normal((B, 8, 12, 5)).astype(f32)
  sk2 = rng.standard_normal((B, 8, 12, 3)).astype(f32)
  out['dcnn_x'], out['dcnn_skip1'], out['dcnn_skip2'] = xd, sk1, sk2
  model = {}
  run = nn.dcnn([3] * 3, dch, dpool, [tf.nn.relu] * 3, [True] * 3, skip_ch=dskip_ch, phase_train=True, wd=5e-5,
                scope='dnet', model=model, init_weights=dw)
  h = run(xd, skip=[None, sk1, sk2])
  for i, hi in enumerate(h):
    out['dcnn_train_h%d' % i] = np.asarray(hi)

  # ---- nn.mlp: relu then softmax (the glimpse MLP form), and nn.lstm unrolled for 3 steps from a zero state
  mdims = [10, 7, 5]
  mw = [{'w': (rng.standard_normal((mdims[i], mdims[i + 1])) / np.sqrt(mdims[i])).astype(f32),
         'b': (rng.standard_normal(mdims[i + 1]) * 0.1).astype(f32)} for i in range(2)]
  for i, d in enumerate(mw):
    out['mlp_w_%d' % i], out['mlp_b_%d' % i] = d['w'], d['b']
  xm = rng.standard_normal((B, mdims[0])).astype(f32)
  out['mlp_x'] = xm
  hm = nn.mlp(mdims, [tf.nn.relu, tf.nn.softmax], init_weights=mw, scope='m')(xm)
  out['mlp_h0'], out['mlp_h1'] = np.asarray(hm[0]), np.asarray(hm[1])
  inp_dim, hid = 6, 9
  lw = {}
  for g in 'ifuo':
    lw['w_x' + g] = (rng.standard_normal((inp_dim, hid)) / np.sqrt(inp_dim)).astype(f32)
    lw['w_h' + g] = (rng.standard_normal((hid, hid)) / np.sqrt(hid)).astype(f32)
    lw['b_' + g] = (rng.standard_normal(hid) * 0.2).astype(f32)
  for k, v in lw.items():
    out['lstm_' + k] = v
  cell = nn.lstm(inp_dim, hid, wd=5e-5, scope='l', init_weights=lw)
  state = np.zeros((B, 2 * hid), f32)
  xs = rng.standard_normal((3, B, inp_dim)).astype(f32)
  out['lstm_x'] = xs
  for step in range(3):
    state, gi, gf, go = cell(xs[step], state)
    out['lstm_state_%d' % step] = np.asarray(state)
  path = os.path.join(HERE, 'nnlib_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, len(out), 'arrays', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
