"""Host logic of FgModel (layer tables, skip sources, BN folding, transposed-conv filter transform, output dict) on
the CPU: the two conv entry points and the head call are replaced by oracle-based stand-ins INSIDE THIS TEST, so the
Python wiring of rec_attend_b200/fg_model.py runs end to end without a GPU and must reproduce
oracle.model.fg_model_forward.  (The kernels themselves are tested on the GPU in tests/test_gpu_fg.py; nothing in the
product package can reach these stand-ins.)"""
import ctypes

import numpy as np
import pytest
import torch

import rec_attend_b200 as ra
from conftest import rel_err
from oracle import model as OM


def _fake_block(x, w, scale, shift, pool=1, relu=True, x2=None, upsample=1, add_to=None, out=None):
  """Contract of ops.conv3x3_block (ra_conv3x3_f32) restated with the oracle's convolutions."""
  xin = x if x2 is None else torch.cat([x, x2], 3)
  cout = w.shape[3]
  if upsample == 1:
    y = OM.conv2d_same(xin, w, torch.zeros(cout))
  else:  # conv-form filter of a transposed conv -> TF layout [kh,kw,Cout,Cin]
    wt = torch.from_numpy(np.ascontiguousarray(w.numpy()[::-1, ::-1].transpose(0, 1, 3, 2)))
    y = OM.conv2d_transpose_same(xin, wt, torch.zeros(cout), upsample)
  y = y * scale + shift
  if relu:
    y = torch.relu(y)
  if pool == 2:
    y = OM.max_pool_same(y, 2)
  out.copy_(y)
  return out


def _arr(p, n):
  return np.ctypeslib.as_array((ctypes.c_float * n).from_address(p.value)) if p.value else None


def _fake_call(name, *a):
  """Contract of ra_fg_head_f32 without ground truth (inference): sigmoid / softmax heads."""
  assert name == 'ra_fg_head_f32', name
  lgp, npix, nsc, nori, ygp, dgp, _, yop, dop, yhp, _, _, _ = a
  assert not ygp.value and not dgp.value
  lg = torch.from_numpy(_arr(lgp, npix * (nsc + nori)).reshape(1, 1, npix, nsc + nori).copy())
  y = torch.sigmoid(lg[..., :nsc]) if nsc == 1 else torch.softmax(lg[..., :nsc], 3)
  _arr(yop, npix * nsc)[:] = y.numpy().reshape(-1)
  hard = (y > 0.5).float() if nsc == 1 else (y == y.max(dim=3, keepdim=True)[0]).float()
  _arr(yhp, npix * nsc)[:] = hard.numpy().reshape(-1)
  if nori:
    _arr(dop, npix * nori)[:] = torch.softmax(lg[..., nsc:], 3).numpy().reshape(-1)


@pytest.mark.parametrize('arch,over', [('default', {'add_skip_conn': True}), ('default', {}), ('kitti', {}),
                                       ('cityscapes', {})])
def test_fg_model_wiring_on_cpu(monkeypatch, arch, over):
  from rec_attend_b200 import _lib, fg_model, ops
  monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
  monkeypatch.setattr(ops, 'conv3x3_block', _fake_block)
  monkeypatch.setattr(ops, '_chk', lambda *a: None)
  monkeypatch.setattr(ops, '_stream', lambda: ctypes.c_void_p(0))
  monkeypatch.setattr(_lib, 'call', _fake_call)
  monkeypatch.setenv('RA_CONV_FP32', '1')  # route every layer through the (replaced) fp32 entry point
  opt = ra.config.fg_model_opt(arch, 64, 128, **over)
  w = ra.synthetic.make_fg_weights(opt, seed=77)
  b = ra.synthetic.make_fg_batch(opt, 2, seed=5)
  with torch.no_grad():
    ref = OM.fg_model_forward(opt, w, b)
  model = fg_model.FgModel(opt, device='cpu').load_weights(w)
  out = model.forward({'x': b['x']})
  assert set(model.conv_kernels(2)) == {'fp32'}
  assert rel_err(out['logits'].numpy(), ref['logits'].numpy()) < 1e-4
  assert rel_err(out['y_out'].numpy(), ref['y_out'].numpy()) < 1e-4
  assert float((out['y_out_hard'] != ref['y_out_hard']).float().mean()) < 1e-3
  if opt['add_orientation']:
    assert rel_err(out['d_out'].numpy(), ref['d_out'].numpy()) < 1e-4
  assert 'loss' not in out
  packed = model.pack_outputs(out)
  assert packed['y_in'].shape == (2, 64, 128, opt['num_semantic_classes'])
