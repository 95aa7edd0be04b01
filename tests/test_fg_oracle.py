"""CPU tests of the foreground / orientation FCN restatement (oracle.model.fg_model_forward, fg_model.py:11-245) and
of the host-side wiring tables (config.fg_model_opt / fg_skip_wiring).  The restatement is pinned to the
reference's own fg_model.py executed over the TF-0.12 stand-in (last test; its missing `image_ops_old` import is
aliased to the reference's image_ops.py)."""
import numpy as np
import pytest
import torch

import rec_attend_b200 as ra
from oracle import model as OM


def test_skip_wiring_follows_fg_model_py():
  # fg_model_train.py:436-437 defaults with --add_skip_conn: CNN mask picks x, h_cnn[5], h_cnn[7]; the DCNN mask hands
  # them out backwards: layer 2 <- h_cnn[7] (64 ch), layer 4 <- h_cnn[5] (32 ch), layer 10 <- x (3 ch)
  opt = ra.config.fg_model_opt('default', 64, 128, add_skip_conn=True)
  src, ch = ra.config.fg_skip_wiring(opt)
  assert src == [None, None, 8, None, 6, None, None, None, None, None, 0]
  assert ch == [0, 0, 64, 0, 32, 0, 0, 0, 0, 0, 3]
  assert ra.config.fg_skip_wiring(ra.config.fg_model_opt('default', 64, 128)) == ([None] * 11, [0] * 11)
  # run_kitti.sh:15-18
  src, ch = ra.config.fg_skip_wiring(ra.config.fg_model_opt('kitti', 64, 128))
  assert src == [None, 17, None, 13, None, 5, None, None, None, None, 0] and ch[1] == 256 and ch[3] == 128 and ch[5] == 96
  # run_cityscapes.sh passes one DCNN mask entry too many; nnlib.dcnn reads only the first nlayers
  src, ch = ra.config.fg_skip_wiring(ra.config.fg_model_opt('cityscapes', 64, 128))
  assert len(src) == 13 and src[1] == 17 and ch[1] == 512 and src[11] == 0 and src[12] is None
  with pytest.raises(ValueError):
    ra.config.fg_skip_wiring(ra.config.fg_model_opt('default', 64, 128, add_skip_conn=True, dcnn_skip_mask=[1] * 10))
  with pytest.raises(ValueError):
    ra.config.fg_skip_wiring(ra.config.fg_model_opt('default', 64, 128, add_skip_conn=True, dcnn_skip_mask=[0, 1]))
  # the oracle's own list agrees with the table (activations identified by shape)
  x = torch.zeros(1, 64, 128, 3)
  h = [torch.zeros(1, 1, 1, c) for c in opt['cnn_depth']]
  sk = OM.fg_skip_lists(opt, x, h)
  assert sk[2].shape[3] == 64 and sk[4].shape[3] == 32 and sk[10] is x and sk[1] is None


def test_weight_schema_and_shapes():
  opt = ra.config.fg_model_opt('kitti', 64, 128)
  w = ra.synthetic.make_fg_weights(opt)
  assert w['cnn_w_0'].shape == (3, 3, 3, 32) and w['cnn_w_17'].shape == (3, 3, 256, 512)
  assert w['dcnn_w_1'].shape == (3, 3, 256, 256 + 256)  # [kh, kw, Cout, Cin + skip] (nnlib.py:320-325)
  assert w['dcnn_w_10'].shape == (3, 3, 9, 32 + 3)
  assert 'cnn_17_0_ema_var' in w and 'dcnn_9_0_gamma' in w and 'dcnn_10_0_gamma' not in w  # last layer: no BN


@pytest.mark.parametrize('arch', ['default', 'kitti', 'cityscapes'])
def test_forward_outputs_and_loss_definitions(arch):
  H, W = (64, 128) if arch == 'cityscapes' else (32, 64)  # the Cityscapes FCN pools by 64
  opt = ra.config.fg_model_opt(arch, H, W)
  B = 2
  w = ra.synthetic.make_fg_weights(opt, seed=1)
  b = ra.synthetic.make_fg_batch(opt, B, seed=2)
  with torch.no_grad():
    r = OM.fg_model_forward(opt, w, b)
  nsc = opt['num_semantic_classes']
  assert r['y_out'].shape == (B, H, W, nsc)
  y, yg = r['y_out'], torch.from_numpy(b['y_gt'])
  if nsc == 1:
    assert torch.equal(r['y_out'], torch.sigmoid(r['logits'][..., :1]))
    inter = (y * yg).sum()
    assert float(r['iou_soft']) == pytest.approx(float(inter / (y.sum() + yg.sum() - inter + 1e-5)), rel=1e-6)
  else:
    assert torch.allclose(y.sum(3), torch.ones(B, H, W), atol=1e-5)
    assert float(r['y_out_hard'].sum(3).min()) >= 1.0  # one-hot of the maximum (ties may mark several)
  if opt['add_orientation']:
    assert torch.allclose(r['d_out'].sum(3), torch.ones(B, H, W), atol=1e-5)
    assert float(r['loss']) == pytest.approx(float(r['foreground_loss'] + r['orientation_ce']), rel=1e-6)
    assert 0.0 <= float(r['orientation_acc']) <= 1.0
    # pixels outside the foreground mask do not count: changing their labels changes nothing
    fg = (b['y_gt'][..., 1:].max(3) if nsc > 1 else b['y_gt'][..., 0]) > 0
    d2 = b['d_gt'].copy()
    d2[~fg] = np.eye(8, dtype=np.float32)[3]
    r2 = OM.fg_model_forward(opt, w, dict(b, d_gt=d2))
    assert float(r2['orientation_acc']) == float(r['orientation_acc'])
    assert float(r2['orientation_ce']) == pytest.approx(float(r['orientation_ce']), rel=1e-6)
  if opt['segm_loss_fn'] == 'iou':
    assert float(r['foreground_loss']) == -float(r['iou_soft'])
  else:
    assert float(r['foreground_loss']) > 0.0
  with pytest.raises(ValueError):
    OM.fg_model_forward(dict(opt, num_semantic_classes=nsc + 1), w, b)


def test_last_layer_has_no_bn_and_no_relu():
  opt = ra.config.fg_model_opt('default', 32, 64)
  w = ra.synthetic.make_fg_weights(opt, seed=3)
  b = ra.synthetic.make_fg_batch(opt, 1, seed=4)
  with torch.no_grad():
    r = OM.fg_model_forward(opt, w, b)
    assert float(r['logits'].min()) < 0.0  # a ReLU would clip these
    w2 = dict(w, dcnn_b_10=w['dcnn_b_10'] + np.float32(1.5))
    r2 = OM.fg_model_forward(opt, w2, b)
  assert torch.allclose(r2['logits'], r['logits'] + 1.5, atol=1e-5)  # the bias reaches the logits unscaled


import json  # noqa: E402
import os  # noqa: E402

from conftest import oracle_fp64  # noqa: E402

GF = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'fg_model_golden.npz'))


@pytest.mark.parametrize('name', sorted({k.split('/')[0] for k in GF.files}))
def test_fg_oracle_equals_the_reference_graph(name):
  """oracle.model.fg_model_forward against the reference's OWN fg_model.get_model(opt), executed unmodified over the
  numpy TF-0.12 stand-in (tests/golden/make_fg_model_golden.py; float64, training mode, the graph's own initial
  weights).  Pins the FCN restatement - skip wiring incl. run_cityscapes.sh's over-long mask, last layer without BN /
  activation, sigmoid / softmax heads, IoU / BCE / CE / orientation losses - to the reference's code."""
  meta = json.loads(str(GF[name + '/meta']))
  opt = ra.config.fg_model_opt(meta['arch'], meta['H'], meta['W'], **meta['overrides'])
  opt['cnn_depth'], opt['dcnn_depth'] = meta['cnn_depth'], meta['dcnn_depth']
  batch = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_fg_batch(opt, meta['B'], seed=meta['batch_seed']).items()}
  w = {k[len(name) + 3:]: GF[k].astype(np.float64) for k in GF.files if k.startswith(name + '/w/')}
  for k in list(w):  # TensorFlow's EMA shadows of tensors start at zero (irrelevant in training mode)
    if k.endswith('_gamma'):
      w[k[:-5] + 'ema_mean'] = np.zeros_like(w[k])
      w[k[:-5] + 'ema_var'] = np.zeros_like(w[k])
  O64 = oracle_fp64()
  torch.set_default_dtype(torch.float64)
  try:
    with torch.no_grad():
      o = O64.fg_model_forward(opt, w, batch, phase_train=True)
  finally:
    torch.set_default_dtype(torch.float32)
  for k in ('y_out', 'd_out', 'iou_soft', 'iou_hard', 'foreground_loss', 'orientation_ce', 'orientation_acc', 'loss'):
    key = '%s/%s' % (name, k)
    if key not in GF.files:
      assert k in ('d_out', 'orientation_ce', 'orientation_acc') and not opt['add_orientation']
      continue
    a, b = np.asarray(o[k].numpy(), np.float64), GF[key].astype(np.float64)
    tol = 1e-6 if GF[key].dtype == np.float32 else 1e-9
    assert float(np.abs(a - b).max()) <= tol * max(float(np.abs(b).max()), 1e-9), (k, float(np.abs(a - b).max()))
