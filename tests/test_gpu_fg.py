"""GPU parity tests of the foreground / orientation FCN (fg_model.py): head + loss kernel against
the oracle's formulas, and FgModel.forward against oracle.model.fg_model_forward for the shipped architectures
(fg_model_train.py defaults, run_kitti.sh, run_cityscapes.sh) at reduced size.  Tolerance 1e-3;
statistics of thresholded / arg-max outputs are discontinuous and get the looser HARD_TOL after a flip-count check."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import model as OM

pytestmark = pytest.mark.gpu

TOL = 1e-3
HARD_TOL = 5e-3


def _g(a):
  return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _head_oracle(lg, nsc, nori, y_gt, d_gt, fn):
  """fg_model.py:174-239 on given logits (the tail of oracle.model.fg_model_forward)."""
  lg, y_gt = torch.from_numpy(lg), torch.from_numpy(y_gt)
  r = {}
  y = torch.sigmoid(lg[..., :nsc]) if nsc == 1 else torch.softmax(lg[..., :nsc], dim=3)
  r['y_out'] = y
  npix = float(lg.shape[0] * lg.shape[1] * lg.shape[2])
  if nsc == 1:
    hard, mask = (y > 0.5).float(), y_gt
    r['iou_soft'], r['iou_hard'] = OM.f_iou_all(y, y_gt), OM.f_iou_all(hard, y_gt)
    seg = OM.f_bce(y, y_gt).sum() / npix
  else:
    hard = (y == y.max(dim=3, keepdim=True)[0]).float()
    mask = y_gt[..., 1:].max(dim=3, keepdim=True)[0]
    r['iou_soft'], r['iou_hard'] = OM.f_iou_all(y[..., 1:], y_gt[..., 1:]), OM.f_iou_all(hard[..., 1:], y_gt[..., 1:])
    seg = OM.f_ce(y, y_gt).sum() / npix
  r['y_out_hard'] = hard
  r['segloss'] = seg
  r['foreground_loss'] = seg if fn == 'bce' else -r['iou_soft']
  r['loss'] = r['foreground_loss']
  if nori:
    d = torch.softmax(lg[..., nsc:], dim=3)
    dg = torch.from_numpy(d_gt)
    r['d_out'] = d
    r['orientation_ce'] = (OM.f_ce(d, dg) * mask).sum() / mask.sum()
    r['orientation_acc'] = ((d.argmax(3) == dg.argmax(3)).float() * mask[..., 0]).sum() / mask.sum()
    r['loss'] = r['loss'] + r['orientation_ce']
  return r


@pytest.mark.parametrize('nsc,nori,fn', [(1, 0, 'iou'), (1, 8, 'bce'), (9, 8, 'bce'), (3, 0, 'iou')])
def test_fg_head_kernel(cuda, nsc, nori, fn):
  import ctypes
  from rec_attend_b200 import _lib, ops
  rng = np.random.default_rng(3 + nsc + nori)
  B, H, W = 2, 37, 53  # odd sizes: the pass is per pixel
  lg = (rng.standard_normal((B, H, W, nsc + nori)) * 2.0).astype(np.float32)
  cls = rng.integers(0, max(nsc, 2), (B, H, W))
  y_gt = (cls > 0).astype(np.float32)[..., None] if nsc == 1 else np.eye(nsc, dtype=np.float32)[cls]
  d_gt = None
  if nori:
    d_gt = (np.eye(nori, dtype=np.float32)[rng.integers(0, nori, (B, H, W))] * (cls > 0)[..., None]).astype(np.float32)
  ref = _head_oracle(lg, nsc, nori, y_gt, d_gt, fn)
  npix = B * H * W
  y_out = torch.empty((B, H, W, nsc), device='cuda')
  y_hard = torch.empty((B, H, W, nsc), device='cuda')
  d_out = torch.empty((B, H, W, nori), device='cuda') if nori else None
  scal = torch.full((8,), -7.0, device='cuda')
  ws = torch.empty(int(_lib.lib().ra_fg_head_workspace()), device='cuda', dtype=torch.uint8)
  args = lambda yg, dg, out, w: (ops._p(_g(lg)), npix, nsc, nori, ops._p(yg), ops._p(dg), 1 if fn == 'bce' else 0,
                                 ops._p(y_out), ops._p(d_out), ops._p(y_hard), ops._p(out),
                                 ctypes.c_void_p(w.data_ptr() if w is not None else 0), ops._stream())
  _lib.call('ra_fg_head_f32', *args(_g(y_gt), None if d_gt is None else _g(d_gt), scal, ws))
  torch.cuda.synchronize()
  assert rel_err(y_out.cpu().numpy(), ref['y_out'].numpy()) < 1e-5
  if nori:
    assert rel_err(d_out.cpu().numpy(), ref['d_out'].numpy()) < 1e-5
  flips = float((y_hard.cpu() != ref['y_out_hard']).float().mean())
  assert flips <= 1e-3, flips
  s = scal.cpu().numpy()
  from rec_attend_b200.fg_model import FG_SCALARS
  for k, i in FG_SCALARS.items():
    if k in ref:
      tol = HARD_TOL if k in ('iou_hard', 'orientation_acc') else 1e-4
      assert abs(float(s[i]) - float(ref[k])) <= tol * max(1.0, abs(float(ref[k]))), (k, float(s[i]), float(ref[k]))
  # inference only: no ground truth, no scalars, outputs unchanged
  y_prev = y_out.clone()
  scal.fill_(-7.0)
  _lib.call('ra_fg_head_f32', *args(None, None, None, None))
  torch.cuda.synchronize()
  assert torch.equal(y_out, y_prev) and float(scal[0]) == -7.0
  with pytest.raises(_lib.RecAttendError):  # ground truth without the scalar output / workspace
    _lib.call('ra_fg_head_f32', *args(_g(y_gt), None if d_gt is None else _g(d_gt), None, None))
  if nori:
    with pytest.raises(_lib.RecAttendError):  # the orientation head needs d_gt next to y_gt
      _lib.call('ra_fg_head_f32', *args(_g(y_gt), None, scal, ws))


FG_CASES = [
    # name, arch, H, W, B, overrides, fp32 convs only
    ('default_skip_64x128', 'default', 64, 128, 2, {'add_skip_conn': True}, False),
    ('default_noskip_iou_32x64', 'default', 32, 64, 3, {}, False),
    ('kitti_64x128', 'kitti', 64, 128, 2, {}, False),
    ('kitti_64x128_fp32_convs', 'kitti', 64, 128, 2, {}, True),
    ('cityscapes_64x128', 'cityscapes', 64, 128, 1, {}, False),
]


@pytest.mark.parametrize('case', FG_CASES, ids=[c[0] for c in FG_CASES])
def test_fg_model_parity(cuda, case, monkeypatch):
  import rec_attend_b200 as ra
  from rec_attend_b200.fg_model import FgModel
  _, arch, H, W, B, over, fp32 = case
  if fp32:
    monkeypatch.setenv('RA_CONV_FP32', '1')
  else:
    monkeypatch.delenv('RA_CONV_FP32', raising=False)
  opt = ra.config.fg_model_opt(arch, H, W, **over)
  weights = ra.synthetic.make_fg_weights(opt, seed=77)
  batch = ra.synthetic.make_fg_batch(opt, B, seed=5)
  with torch.no_grad():
    ref = OM.fg_model_forward(opt, weights, batch)
  model = FgModel(opt).load_weights(weights)
  out = model.forward(batch)
  torch.cuda.synchronize()
  kinds = model.conv_kernels(B)
  assert len(kinds) == len(opt['cnn_depth']) + len(opt['dcnn_depth'])
  if fp32:
    assert set(kinds) == {'fp32'}
  else:
    assert kinds.count('umma') >= len(kinds) - 8, kinds  # only the 512-channel layers lack a tile plan
  for k in ('logits', 'y_out') + (('d_out',) if opt['add_orientation'] else ()):
    a, b = out[k].cpu().numpy(), ref[k].numpy()
    assert a.shape == b.shape, (k, a.shape, b.shape)
    assert rel_err(a, b) <= TOL, (k, rel_err(a, b))
  flips = float((out['y_out_hard'].cpu() != ref['y_out_hard']).float().mean())
  assert flips <= 2e-3, flips
  keys = ['iou_soft', 'iou_hard', 'foreground_loss', 'loss'] + (['orientation_ce', 'orientation_acc']
                                                                if opt['add_orientation'] else [])
  for k in keys:
    tol = HARD_TOL if k in ('iou_hard', 'orientation_acc') else TOL
    assert abs(float(out[k]) - float(ref[k])) <= tol * max(1.0, abs(float(ref[k]))), (k, float(out[k]), float(ref[k]))
  # inference-only call (fg_model_pack.py): same maps, no scalars; pack_outputs names them as the hot path's inputs
  inf = model.forward({'x': batch['x']})
  torch.cuda.synchronize()
  assert 'loss' not in inf and rel_err(inf['y_out'].cpu().numpy(), ref['y_out'].numpy()) <= TOL
  packed = model.pack_outputs(inf)
  assert packed['y_in'].shape == (B, H, W, opt['num_semantic_classes'])
  assert (packed['d_in'] is None) == (not opt['add_orientation'])


def test_fg_model_errors(cuda):
  import rec_attend_b200 as ra
  from rec_attend_b200 import _lib
  from rec_attend_b200.fg_model import FgModel
  opt = ra.config.fg_model_opt('default', 32, 64)
  with pytest.raises(_lib.RecAttendError):
    FgModel(dict(opt, dcnn_depth=opt['dcnn_depth'][:-1] + [2]))  # last channel must be num_semantic_classes
  with pytest.raises(_lib.RecAttendError):
    FgModel(dict(opt, inp_height=40))  # not divisible by the pooling factor
  m = FgModel(opt)
  with pytest.raises(_lib.RecAttendError):
    m.forward({'x': np.zeros((1, 32, 64, 3), np.float32)})  # load_weights() first
  m.load_weights(ra.synthetic.make_fg_weights(opt))
  with pytest.raises(_lib.RecAttendError):
    m.forward({'x': np.zeros((1, 32, 32, 3), np.float32)})
  with pytest.raises(_lib.RecAttendError):
    m.forward({'x': np.zeros((1, 32, 64, 3), np.float32)}, phase_train=True)
  w = ra.synthetic.make_fg_weights(opt)
  w['dcnn_w_3'] = w['dcnn_w_3'][:, :, :, :-1]
  with pytest.raises(_lib.RecAttendError):
    FgModel(opt).load_weights(w)
