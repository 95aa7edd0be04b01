"""oracle/model.py's layer functions against golden vectors produced by EXECUTING THE REFERENCE'S OWN nnlib.py
(tests/golden/make_nnlib_golden.py: nn.cnn / nn.dcnn / nn.mlp / nn.lstm imported unmodified over the numpy stand-in
of tests/golden/tf012_shim).  Pins the layer glue of the path to the reference's code: conv + bias -> BN -> ReLU ->
pool order, one BN copy per call, batch statistics and EMA update in training mode, EMA (zero-initialised) statistics
in eval mode, skip concatenation and the transposed-conv filter layout, the LSTM equations and state layout, the MLP.
CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import model as OM

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'nnlib_golden.npz'))
t = lambda k: torch.from_numpy(G[k])


def close(a, b, tol=1e-5):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  assert a.shape == b.shape, (a.shape, b.shape)
  scale = max(float(np.abs(b).max()), 1e-6)
  assert float(np.abs(a - b).max()) <= tol * scale, float(np.abs(a - b).max()) / scale


def _weights(prefix, scope, nlayers, copies, with_bn_last=True):
  w = {}
  for i in range(nlayers):
    w['%s_w_%d' % (scope, i)] = t('%s_%d_w' % (prefix, i))
    w['%s_b_%d' % (scope, i)] = t('%s_%d_b' % (prefix, i))
    for c in range(copies):
      k = '%s_%d_%d_' % (scope, i, c)
      w[k + 'beta'], w[k + 'gamma'] = t('%s_%d_beta_%d' % (prefix, i, c)), t('%s_%d_gamma_%d' % (prefix, i, c))
      n = w[k + 'beta'].shape[0]
      w[k + 'ema_mean'], w[k + 'ema_var'] = torch.zeros(n), torch.zeros(n)  # TF's EMA shadows of tensors start at 0
  return w


@pytest.mark.parametrize('phase', ['train', 'eval'])
def test_cnn_layers_bn_copies_and_ema(phase):
  w = _weights('cnn_w', 'net', 3, 2)
  x = t('cnn_x')
  for call in range(2):
    xin = x if call == 0 else torch.flip(x, [1])
    ema_out = {} if phase == 'train' else None
    with torch.no_grad():
      h = OM.run_cnn(xin, w, 'net', 3, [1, 2, 2], call, ema_out=ema_out)
    for i in range(3):
      close(h[i], G['cnn_%s_call%d_h%d' % (phase, call, i)])
    if phase == 'train':
      for i in range(3):
        for n in ('ema_mean', 'ema_var'):
          close(ema_out['net_%d_%d_%s' % (i, call, n)], G['cnn_train_%d_%d_%s' % (i, call, n)], tol=2e-5)
  # the two calls used different BN parameters: copy 1's outputs differ from what copy 0 would give on the same input
  with torch.no_grad():
    wrong = OM.run_cnn(torch.flip(x, [1]), w, 'net', 3, [1, 2, 2], 0, ema_out={} if phase == 'train' else None)
  assert float((wrong[2] - t('cnn_%s_call1_h2' % phase)).abs().max()) > 1e-2


def test_dcnn_skip_concat_and_transposed_filters():
  w = _weights('dcnn_w', 'dnet', 3, 1)
  assert w['dnet_w_1'].shape == (3, 3, 6, 6 + 5) and w['dnet_w_2'].shape == (3, 3, 2, 6 + 3)  # [kh,kw,Cout,Cin+skip]
  with torch.no_grad():
    h = OM.run_dcnn(t('dcnn_x'), w, 'dnet', 3, [2, 1, 2], 0, skip=[None, t('dcnn_skip1'), t('dcnn_skip2')], ema_out={})
  assert h[0].shape == (3, 8, 12, 6) and h[2].shape == (3, 16, 24, 2)
  for i in range(3):
    close(h[i], G['dcnn_train_h%d' % i])


def test_mlp_and_lstm():
  w = {'m_w_0': t('mlp_w_0'), 'm_b_0': t('mlp_b_0'), 'm_w_1': t('mlp_w_1'), 'm_b_1': t('mlp_b_1')}
  h = OM.run_mlp(t('mlp_x'), w, 'm', ['relu', 'softmax'])
  close(h[0], G['mlp_h0'])
  close(h[1], G['mlp_h1'])
  lw = {'l_' + k[len('lstm_'):]: t(k) for k in G.files if k.startswith('lstm_w_') or k.startswith('lstm_b_')}
  hid = lw['l_b_i'].shape[0]
  state = torch.zeros(G['lstm_x'].shape[1], 2 * hid)
  for step in range(3):
    state = OM.lstm_step(t('lstm_x')[step], state, lw, hid, scope='l')
    close(state, G['lstm_state_%d' % step])  # state = concat(c, h)
