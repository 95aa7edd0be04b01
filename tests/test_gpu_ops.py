"""GPU parity tests, operator level: every C-ABI entry point of librecattend_b200.so against
the CPU oracle (oracle/) on identical seeded inputs.  Integer/assignment results are
bit-exact; fp32 results within the tolerance written next to each assert."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import hungarian as OH
from oracle import model as OM

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'hungarian_kat.json')
TOL = 1e-4  # operator-level fp32 tolerance (summation order only); the model-level bar is 1e-3


@pytest.fixture(scope='module')
def ops(cuda):
  from rec_attend_b200 import ops as _ops
  return _ops


def _g(a):
  return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ----------------------------------------------------------------------------- Hungarian
def test_hungarian_golden_vectors(ops):
  for case in json.load(open(GOLD))['cases']:
    W = np.frombuffer(bytes.fromhex(case['W_f32_hex']), np.float32).reshape(case['shape'])
    M, cx, cy, st = ops.hungarian(_g(W))
    Mo, cxo, cyo = OH.hungarian(W)
    assert (M.cpu().numpy() == Mo).all(), case['name']
    assert (cx.cpu().numpy() == cxo).all() and (cy.cpu().numpy() == cyo).all(), case['name']
    assert int(st.max()) == 0
    if case['kind'] == 'known_answer':
      assert (M.cpu().numpy() == np.array(case['M'], np.float32)).all()


def test_hungarian_random_bit_exact(ops):
  from test_hungarian_oracle import _random_weights
  rng = np.random.default_rng(11)
  for it in range(60):
    kind = it % 5
    nx, ny = int(rng.integers(1, 34)), int(rng.integers(1, 34))
    if it % 3 == 0:
      ny = nx
    if it % 20 == 0:
      nx = ny = 64
    if it % 20 == 1:
      nx, ny = 64, 40
    B = 48
    W = _random_weights(rng, kind, B, nx, ny)
    M, cx, cy, st = ops.hungarian(_g(W))
    Mo, cxo, cyo, info = OH.hungarian(W, return_info=True)
    assert (M.cpu().numpy() == Mo).all(), (kind, nx, ny)
    assert (cx.cpu().numpy() == cxo).all() and (cy.cpu().numpy() == cyo).all(), (kind, nx, ny)
    assert ((st.cpu().numpy() & 1) == (info['status'] & 1)).all()


def test_hungarian_errors_and_edges(ops):
  from rec_attend_b200 import _lib
  with pytest.raises(ValueError):
    ops.hungarian(torch.zeros(3, device='cuda'))
  with pytest.raises(_lib.RecAttendError):
    ops.hungarian(torch.zeros(1, 65, 65, device='cuda'))
  M, cx, cy, st = ops.hungarian(torch.zeros(0, 4, 4, device='cuda'))
  assert M.shape == (0, 4, 4)
  # 1x1 and all-equal weights
  M, _, _, _ = ops.hungarian(_g(np.array([[0.3]], np.float32)))
  assert float(M[0, 0]) == 1.0
  W = np.full((2, 5, 5), 1e-5, np.float32)
  M, cx, cy, _ = ops.hungarian(_g(W))
  Mo, cxo, cyo = OH.hungarian(W)
  assert (M.cpu().numpy() == Mo).all()


def test_hungarian_host_entry_point(cuda):
  import ctypes
  from rec_attend_b200 import _lib
  rng = np.random.default_rng(5)
  W = rng.random((6, 7, 9)).astype(np.float32)
  M = np.zeros_like(W)
  cx = np.zeros((6, 7), np.float32)
  cy = np.zeros((6, 9), np.float32)
  st = np.zeros((6,), np.int32)
  vp = lambda a: ctypes.c_void_p(a.ctypes.data)
  _lib.call('ra_hungarian_f32_host', vp(W), 6, 7, 9, vp(M), vp(cx), vp(cy), vp(st))
  Mo, cxo, cyo = OH.hungarian(W)
  assert (M == Mo).all() and (cx == cxo[..., 0]).all() and (cy == cyo[:, 0]).all()


def test_segm_match(ops):
  rng = np.random.default_rng(2)
  B, T = 16, 20
  iou = (rng.random((B, T, T))**3).astype(np.float32)
  s_gt = np.zeros((B, T), np.float32)
  for b in range(B):
    s_gt[b, :rng.integers(0, T + 1)] = 1
  match, w, st = ops.f_segm_match(_g(iou), _g(s_gt), return_weights=True)
  wo = OM.segm_match_weights(torch.from_numpy(iou), torch.from_numpy(s_gt)).numpy()
  assert (w.cpu().numpy() == wo).all(), 'the fp32 matrix handed to the matcher must be bit-identical'
  mo = OM.f_segm_match(torch.from_numpy(iou), torch.from_numpy(s_gt)).numpy()
  assert (match.cpu().numpy() == mo).all()


# ----------------------------------------------------------------------------- conv blocks
CONV_CASES = [
    # B, H, W, C1, C2, Cout, up, pool, relu, add_to
    (2, 32, 64, 13, 0, 16, 1, 2, 1, False),
    (2, 64, 64, 1, 0, 16, 1, 2, 1, True),
    (1, 128, 128, 4, 0, 8, 1, 1, 1, False),
    (2, 24, 24, 32, 0, 64, 1, 2, 1, False),
    (2, 12, 12, 64, 0, 96, 1, 2, 1, False),
    (2, 6, 6, 96, 0, 64, 2, 1, 1, False),
    (2, 12, 12, 64, 64, 64, 1, 1, 1, False),
    (2, 24, 24, 32, 32, 16, 2, 1, 1, False),
    (2, 48, 48, 16, 13, 1, 1, 1, 1, False),
    (1, 16, 16, 12, 0, 16, 1, 1, 0, False),
    (1, 70, 34, 8, 0, 8, 1, 2, 1, False),
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv3x3_block(ops, case):
  B, H, W, C1, C2, Cout, up, pool, relu, use_add = case
  rng = np.random.default_rng(hash(case) % 2**31)
  x1 = rng.standard_normal((B, H, W, C1)).astype(np.float32)
  x2 = rng.standard_normal((B, H, W, C2)).astype(np.float32) if C2 else None
  Cin = C1 + C2
  scale = rng.uniform(0.5, 1.5, Cout).astype(np.float32)
  shift = rng.standard_normal(Cout).astype(np.float32)
  xin = torch.from_numpy(x1 if x2 is None else np.concatenate([x1, x2], 3))
  if up == 1:
    w = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    ref = OM.conv2d_same(xin, torch.from_numpy(w), torch.zeros(Cout))
    w_dev = w
  else:
    from rec_attend_b200.full_model import _deconv_to_conv
    wt = (rng.standard_normal((3, 3, Cout, Cin)) / np.sqrt(9 * Cin)).astype(np.float32)
    ref = OM.conv2d_transpose_same(xin, torch.from_numpy(wt), torch.zeros(Cout), up)
    w_dev = _deconv_to_conv(wt)
  add = None
  if use_add:
    add = rng.standard_normal(tuple(ref.shape)).astype(np.float32)
    ref = ref + torch.from_numpy(add)
  ref = ref * torch.from_numpy(scale) + torch.from_numpy(shift)
  if relu:
    ref = torch.relu(ref)
  if pool == 2:
    ref = OM.max_pool_same(ref, 2)
  out = ops.conv3x3_block(_g(x1), _g(w_dev), _g(scale), _g(shift), pool=pool, relu=bool(relu),
                          x2=None if x2 is None else _g(x2), upsample=up, add_to=None if add is None else _g(add))
  assert tuple(out.shape) == tuple(ref.shape)
  assert rel_err(out.cpu().numpy(), ref.numpy()) < TOL


def test_deconv_stride1_matches_transposed_conv(ops):
  from rec_attend_b200.full_model import _deconv_to_conv
  rng = np.random.default_rng(9)
  x = rng.standard_normal((2, 12, 12, 16)).astype(np.float32)
  wt = rng.standard_normal((3, 3, 8, 16)).astype(np.float32) / 12
  ref = OM.conv2d_transpose_same(torch.from_numpy(x), torch.from_numpy(wt), torch.zeros(8), 1)
  one, zero = torch.ones(8, device='cuda'), torch.zeros(8, device='cuda')
  out = ops.conv3x3_block(_g(x), _g(_deconv_to_conv(wt)), one, zero, pool=1, relu=False, upsample=1)
  assert rel_err(out.cpu().numpy(), ref.numpy()) < TOL


def test_concat_channels(ops):
  rng = np.random.default_rng(1)
  a, b, c = (rng.random((2, 5, 7, k)).astype(np.float32) for k in (3, 8, 1))
  out = ops.concat_channels(_g(a), _g(b), _g(c))
  assert (out.cpu().numpy() == np.concatenate([a, b, c], 3)).all()
  out = ops.concat_channels(_g(a))
  assert (out.cpu().numpy() == a).all()


# ----------------------------------------------------------------------------- controller
@pytest.mark.parametrize('arch,H,W', [('kitti', 64, 128), ('cvppp', 128, 128), ('cityscapes', 512, 1024)])
def test_controller_step(ops, arch, H, W):
  import rec_attend_b200 as ra
  from rec_attend_b200 import _lib
  opt = ra.config.full_model_opt(arch, H, W, 3)
  w = ra.synthetic.make_weights(opt, seed=77)
  wt = {k: torch.from_numpy(v) for k, v in w.items()}
  B = 3
  gh, gw = H // 32, W // 32
  P, Cf, Hd = gh * gw, 64, 256
  rng = np.random.default_rng(4)
  feat = np.maximum(rng.standard_normal((B, P, Cf)), 0).astype(np.float32)
  # oracle: the part of controller_step after the CNN
  state = torch.zeros(B, 2 * Hd)
  gmap = torch.ones(B, P, 1) / P
  gmaps = []
  f = torch.from_numpy(feat)
  for it in range(5):
    gmaps.append(gmap[:, :, 0])
    state = OM.lstm_step((f * gmap).sum(1), state, wt, Hd)
    h = state[:, Hd:]
    if it < 4:
      gmap = OM.run_mlp(h, wt, 'glimpse_mlp', ['relu', 'softmax'])[-1].unsqueeze(2)
  ctrl = OM.run_mlp(h, wt, 'ctrl_mlp', [None])[-1]
  flags = 0
  if opt['dynamic_var']:
    flags |= _lib.CTRL_DYNAMIC_VAR
  if opt['fixed_gamma']:
    flags |= _lib.CTRL_FIXED_GAMMA
  gates = 'ifou'
  h_out, ctrl_out, gm, box = ops.controller_step(
      _g(feat), _g(np.stack([w['ctrl_lstm_w_x' + g] for g in gates])),
      _g(np.stack([w['ctrl_lstm_w_h' + g] for g in gates])), _g(np.stack([w['ctrl_lstm_b_' + g] for g in gates])),
      _g(w['glimpse_mlp_w_0']), _g(w['glimpse_mlp_b_0']), _g(w['glimpse_mlp_w_1']), _g(w['glimpse_mlp_b_1']),
      _g(w['ctrl_mlp_w_0']), _g(w['ctrl_mlp_b_0']), H, W, 48, 48, flags)
  assert rel_err(h_out.cpu().numpy(), h.numpy()) < TOL
  assert rel_err(ctrl_out.cpu().numpy(), ctrl.numpy()) < TOL
  assert rel_err(gm.cpu().numpy(), torch.stack(gmaps, 1).numpy()) < TOL
  box = box.cpu().numpy()
  img = np.array([H, W], np.float32)
  ctr = (ctrl[:, 0:2].numpy() + 1) * img / 2
  size = np.exp(ctrl[:, 2:4].numpy()) * img
  assert rel_err(box[:, 0:2], ctr) < TOL and rel_err(box[:, 2:4], size) < TOL
  lgv = ctrl[:, 4:6].numpy() if opt['dynamic_var'] else np.log(size) - np.log(48.0)
  assert np.abs(box[:, 4:6] - lgv).max() < 1e-4
  g_attn = np.ones(B) if opt['fixed_gamma'] else np.exp(ctrl[:, 6].numpy())
  g_y = np.full(B, np.exp(2.0)) if opt['fixed_gamma'] else np.exp(ctrl[:, 8].numpy())
  assert rel_err(box[:, 6], g_attn) < TOL and rel_err(box[:, 7], np.exp(ctrl[:, 7].numpy())) < TOL
  assert rel_err(box[:, 8], g_y) < TOL
  assert rel_err(box[:, 9:11], ctr - size / 2) < TOL and rel_err(box[:, 11:13], ctr + size / 2) < TOL


# ----------------------------------------------------------------------------- attention
def _boxes(rng, B, H, W, dense=False, small=False):
  from rec_attend_b200 import _lib
  box = np.zeros((B, _lib.BOX_STRIDE), np.float32)
  box[:, 0] = rng.uniform(-0.1 * H, 1.1 * H, B)
  box[:, 1] = rng.uniform(-0.1 * W, 1.1 * W, B)
  box[:, 2] = rng.uniform(0.05 * H, (0.3 if small else 1.3) * H, B)
  box[:, 3] = rng.uniform(0.05 * W, (0.3 if small else 1.3) * W, B)
  box[:, 4:6] = rng.uniform(-1.0, 2.5, (B, 2)) if not dense else rng.uniform(5.0, 7.0, (B, 2))
  box[:, 6] = rng.uniform(0.5, 2.0, B)
  box[:, 7] = rng.uniform(20, 200, B)
  box[:, 8] = rng.uniform(5, 50, B)
  return box


@pytest.mark.parametrize('H,W,Cs,dense', [(64, 128, 12, False), (128, 128, 3, False), (32, 64, 12, True),
                                          (64, 64, 20, False), (128, 512, 12, 'small'), (96, 384, 3, 'small')])
def test_gaussian_filters_extract_paste(ops, H, W, Cs, dense):
  rng = np.random.default_rng(H * 1000 + W + Cs)
  B, F = 3, 48
  D = Cs + 1
  # 'small': boxes that leave most 8x128 paste-back tiles / extract column chunks outside every tap's band
  box = _boxes(rng, B, H, W, dense is True, small=(dense == 'small'))
  if dense == 'small':
    box[0, 0:2] = [-0.4 * H, 0.5 * W]  # one box (almost) entirely off the image
  bt = torch.from_numpy(box)
  fy, fx, band = ops.get_gaussian_filter(_g(box), H, W, F)
  fy_o = OM.get_gaussian_filter(bt[:, 0], bt[:, 2], bt[:, 4], H, F)  # [B,H,F]
  fx_o = OM.get_gaussian_filter(bt[:, 1], bt[:, 3], bt[:, 5], W, F)
  assert rel_err(fy.cpu().numpy(), fy_o.transpose(1, 2).numpy()) < TOL
  assert rel_err(fx.cpu().numpy(), fx_o.transpose(1, 2).numpy()) < TOL
  # glimpse
  static_idx = [0, 1, 2] + list(range(4, D))
  chan_map = torch.tensor(static_idx + [3], dtype=torch.int32, device='cuda')
  xs = rng.random((B, H, W, Cs)).astype(np.float32)
  canvas = rng.random((B, H, W)).astype(np.float32)
  full = np.zeros((B, H, W, D), np.float32)
  full[..., static_idx] = xs
  full[..., 3] = canvas
  patch = ops.extract_patch(_g(xs), _g(canvas), chan_map, _g(box), fy, fx, band)
  ref = bt[:, 6].view(-1, 1, 1, 1) * OM.extract_patch(torch.from_numpy(full), fy_o, fx_o, D)
  assert rel_err(patch.cpu().numpy(), ref.numpy()) < TOL
  # paste-back + canvas update, both overwrite modes
  P = np.maximum(rng.standard_normal((B, F, F)), 0).astype(np.float32)
  for dis, bnd in ((False, band), (True, band), (False, None), (True, None)):
    cv = _g(canvas.copy())
    T = 2
    attn_box = torch.zeros((B, T, H, W), device='cuda')
    y_out = torch.zeros((B, T, H, W), device='cuda')
    ops.paste_back(_g(P), _g(box), fy, fx, cv, attn_box=attn_box[:, 1], y_out=y_out[:, 1], out_bstride=T * H * W,
                   disable_overwrite=dis, band=bnd)
    y_ref = OM.extract_patch(torch.from_numpy(P).unsqueeze(3), fy_o.transpose(1, 2), fx_o.transpose(1, 2), 1)[..., 0]
    y_ref = torch.sigmoid(bt[:, 8].view(-1, 1, 1) * y_ref - 5.0)
    if dis:
      y_ref = y_ref * (1 - torch.from_numpy(canvas))
    ones = torch.ones(B, F, F, 1) * bt[:, 7].view(-1, 1, 1, 1)
    b_ref = torch.sigmoid(
        OM.extract_patch(ones, fy_o.transpose(1, 2), fx_o.transpose(1, 2), 1)[..., 0] - 5.0)
    assert rel_err(y_out[:, 1].cpu().numpy(), y_ref.numpy()) < TOL
    assert rel_err(attn_box[:, 1].cpu().numpy(), b_ref.numpy()) < TOL
    assert float(y_out[:, 0].abs().max()) == 0.0 and float(attn_box[:, 0].abs().max()) == 0.0
    assert rel_err(cv.cpu().numpy(), torch.maximum(torch.from_numpy(canvas), y_ref).numpy()) < TOL
  # attention box only (box model)
  only = torch.zeros((B, 1, H, W), device='cuda')
  for bnd in (None, band):
    only.zero_()
    ops.paste_back(None, _g(box), fy, fx, None, attn_box=only[:, 0], y_out=None, out_bstride=H * W, band=bnd)
    assert rel_err(only[:, 0].cpu().numpy(), b_ref.numpy()) < TOL


def test_score(ops):
  rng = np.random.default_rng(8)
  B, Hd, Cd, T = 5, 256, 3456, 4
  h = rng.standard_normal((B, Hd)).astype(np.float32)
  core = rng.standard_normal((B, 6, 6, 96)).astype(np.float32)
  w = (rng.standard_normal(Hd + Cd) / 60).astype(np.float32)
  bias = np.array([0.1], np.float32)
  s = torch.zeros((B, T), device='cuda')
  ops.score(_g(h), _g(core), _g(w), _g(bias), s[:, 2], T)
  ref = 1 / (1 + np.exp(-(np.concatenate([h, core.reshape(B, -1)], 1) @ w + bias[0])))
  assert rel_err(s[:, 2].cpu().numpy(), ref) < TOL and float(s[:, 0].abs().max()) == 0


# ----------------------------------------------------------------------------- loss side
def _masks(rng, B, T, H, W):
  import rec_attend_b200 as ra
  opt = {'inp_height': H, 'inp_width': W, 'timespan': T}
  b = ra.synthetic.make_batch(opt, B, seed=int(rng.integers(1 << 30)))
  return b['y_gt'], b['s_gt']


@pytest.mark.parametrize('W', [96, 90])  # float4 path / scalar path
def test_gt_box(ops, W):
  rng = np.random.default_rng(3)
  B, T, H = 3, 6, 64
  y_gt, _ = _masks(rng, B, T, H, W)
  y_gt[0, 2] = 0  # an empty mask in the middle (modellib.py:696-699)
  tl, br, box, rect, area = ops.get_gt_box(_g(y_gt), padding_ratio=0.2, min_padding=20.0)
  tlo, bro, boxo = OM.get_gt_box(torch.from_numpy(y_gt), padding_ratio=0.2, min_padding=20.0)
  assert (tl.cpu().numpy() == tlo.numpy()).all() and (br.cpu().numpy() == bro.numpy()).all()
  assert (box.cpu().numpy() == boxo.numpy()).all()
  assert (area.cpu().numpy() == y_gt.sum((2, 3))).all()


@pytest.mark.parametrize('T,H,W', [(8, 32, 48), (20, 64, 64), (32, 32, 64)])
def test_pairwise_iou(ops, T, H, W):
  rng = np.random.default_rng(T)
  B = 3
  y_gt, _ = _masks(rng, B, T, H, W)
  a = rng.random((B, T, H, W)).astype(np.float32)**2
  iou = ops.f_iou(_g(a), _g(y_gt))
  ref = OM.f_iou_pairwise(torch.from_numpy(a), torch.from_numpy(y_gt))
  assert rel_err(iou.cpu().numpy(), ref.numpy()) < TOL
  ih, dice = ops.f_iou(_g(a), _g(y_gt), hard_threshold=0.5, want_dice=True)
  ah = torch.from_numpy((a > 0.5).astype(np.float32))
  assert rel_err(ih.cpu().numpy(), OM.f_iou_pairwise(ah, torch.from_numpy(y_gt)).numpy()) < TOL
  assert rel_err(dice.cpu().numpy(), OM.f_dice_pairwise(ah, torch.from_numpy(y_gt)).numpy()) < TOL
  # rectangle form of b == the filled GT boxes
  tl, br, box, rect, _ = ops.get_gt_box(_g(y_gt), padding_ratio=0.2, min_padding=4.0)
  i1 = ops.f_iou(_g(a), box)
  i2 = ops.f_iou(_g(a), None, b_rect=rect)
  assert (i1 == i2).all()


def test_loss_block(ops):
  import rec_attend_b200 as ra
  from rec_attend_b200 import _lib
  rng = np.random.default_rng(6)
  B, T, H, W = 4, 6, 32, 48
  y_gt, s_gt = _masks(rng, B, T, H, W)
  y_out = rng.random((B, T, H, W)).astype(np.float32)
  attn_box = rng.random((B, T, H, W)).astype(np.float32)
  s_out = rng.random((B, T)).astype(np.float32)
  opt = ra.config.full_model_opt('kitti', H, W, T)
  model = {'y_out': torch.from_numpy(y_out), 'attn_box': torch.from_numpy(attn_box), 's_out': torch.from_numpy(s_out)}
  ref = OM.full_model_loss(opt, {}, model, torch.from_numpy(y_gt), torch.from_numpy(s_gt))
  g = lambda k: _g(ref[k].numpy())
  area = _g(y_gt.sum((2, 3)))
  scal = ops.loss_block(g('iou_soft_box_pairwise'), g('match_box'), g('iou_soft_pairwise'), g('match'),
                        g('iou_hard_pairwise'), _g(OM.f_dice_pairwise(torch.from_numpy((y_out > 0.5).astype(np.float32)),
                                                                       torch.from_numpy(y_gt)).numpy()),
                        _g(s_out), _g(s_gt), area, 1.0, 0.0).cpu().numpy()
  for i, k in enumerate(_lib.LOSS_NAMES):
    assert abs(float(scal[i]) - float(ref[k])) <= 1e-5 * max(1.0, abs(float(ref[k]))), (k, scal[i], float(ref[k]))


def test_box_gt_step(ops):
  rng = np.random.default_rng(12)
  B, T, H, W = 3, 5, 32, 64
  y_gt, _ = _masks(rng, B, T, H, W)
  tl, br, boxgt, rect, _ = ops.get_gt_box(_g(y_gt), padding_ratio=0.2, min_padding=10.0)
  attn = rng.random((B, 1, H, W)).astype(np.float32)
  noise = (rng.random((B, H, W)) * 0.3).astype(np.float32)
  canvas0 = (rng.random((B, H, W)) * 0.2).astype(np.float32)
  canvas = _g(canvas0.copy())
  iou_t = torch.zeros((B, T), device='cuda')
  grd = torch.zeros((B, T), device='cuda')
  ops.box_gt_step(_g(attn), H * W, rect, _g(y_gt), _g(noise), H * W, iou_t, T, grd, canvas)
  a = torch.from_numpy(attn)
  bg = boxgt.cpu()
  ref_iou = OM.f_inter(a, bg) / OM.f_union(a, bg)
  assert rel_err(iou_t.cpu().numpy(), ref_iou.numpy()) < TOL
  g = OM.f_greedy_match(ref_iou, torch.zeros(B, T))
  y_sel = (g.view(B, T, 1, 1) * torch.from_numpy(y_gt)).sum(1)
  y_sel = y_sel - y_sel * torch.from_numpy(noise)
  assert rel_err(canvas.cpu().numpy(), torch.maximum(y_sel, torch.from_numpy(canvas0)).numpy()) < TOL


def test_greedy_iou_box(ops):
  """opt['use_iou_box'] (full_model.py:750-754): modellib.f_iou_box + f_greedy_match on the box record, against the
  oracle's restatement; covers disjoint boxes (all-zero row -> every GT shares 1/T), exact ties and the canvas half."""
  import rec_attend_b200 as ra
  rng = np.random.default_rng(21)
  B, T, H, W = 6, 7, 32, 64
  stride = 16  # RA_BOX_STRIDE
  tl_gt = rng.uniform(0, 20, (B, T, 2)).astype(np.float32)
  br_gt = tl_gt + rng.uniform(2, 30, (B, T, 2)).astype(np.float32)
  ctr = rng.uniform(8, 28, (B, 2)).astype(np.float32)
  size = rng.uniform(4, 24, (B, 2)).astype(np.float32)
  ctr[0] = [500.0, 500.0]  # example 0: no overlap with any GT box
  tl_gt[1, 3], br_gt[1, 3] = tl_gt[1, 2], br_gt[1, 2]  # example 1: two identical GT boxes (a tie if they win)
  tl, br = ctr - size / 2.0, ctr + size / 2.0
  tl_gt[2, 4], br_gt[2, 4] = tl[2], br[2]  # example 2: GT box 4 equals the predicted box (IoU exactly 1)
  box = np.zeros((B, stride), np.float32)
  box[:, 0:2], box[:, 2:4], box[:, 9:11], box[:, 11:13] = ctr, size, tl, br
  iou_t = torch.full((B, 2, T), -1.0, device='cuda')
  grd = torch.zeros((B, T), device='cuda')
  ops.greedy_iou_box(_g(box), _g(tl_gt), _g(br_gt), iou_t[:, 1], 2 * T, grd)
  ref_iou = OM.f_iou_box(torch.from_numpy(tl).unsqueeze(1), torch.from_numpy(br).unsqueeze(1), torch.from_numpy(tl_gt),
                         torch.from_numpy(br_gt))
  ref_grd = OM.f_greedy_match(ref_iou, torch.zeros(B, T))
  assert np.array_equal(iou_t[:, 1].cpu().numpy(), ref_iou.numpy())  # same fp32 operations in the same order
  assert (iou_t[:, 0] == -1.0).all()  # the batch stride is honoured
  assert np.array_equal(grd.cpu().numpy(), ref_grd.numpy())
  assert np.allclose(grd[0].cpu().numpy(), 1.0 / T) and float(iou_t[2, 1, 4]) == 1.0
  # canvas half alone (box_model.py:497-503), with and without noise
  y_gt, _ = _masks(rng, B, T, H, W)
  noise = (rng.random((B, H, W)) * 0.3).astype(np.float32)
  canvas0 = (rng.random((B, H, W)) * 0.2).astype(np.float32)
  for nz in (noise, None):
    canvas = _g(canvas0.copy())
    ops.box_gt_canvas(grd, _g(y_gt), None if nz is None else _g(nz), H * W, canvas)
    y_sel = (ref_grd.view(B, T, 1, 1) * torch.from_numpy(y_gt)).sum(1)
    if nz is not None:
      y_sel = y_sel - y_sel * torch.from_numpy(nz)
    assert rel_err(canvas.cpu().numpy(), torch.maximum(y_sel, torch.from_numpy(canvas0)).numpy()) < TOL
  with pytest.raises(ra._lib.RecAttendError):
    ops.greedy_iou_box(_g(box), _g(tl_gt), _g(br_gt), iou_t[:, 1], T - 1, grd)  # batch stride shorter than a row


# ----------------------------------------------------------------------------- tcgen05 conv
UMMA_CASES = [
    # B, H, W, C1, C2, Cout, up, pool, relu   (the KITTI/Cityscapes-arch layers + odd shapes)
    (2, 128, 256, 16, 0, 16, 1, 2, 1),   # ctrl L1
    (2, 64, 128, 16, 0, 32, 1, 1, 1),    # ctrl L2
    (2, 64, 128, 32, 0, 32, 1, 2, 1),    # ctrl L3
    (2, 32, 64, 32, 0, 64, 1, 1, 1),     # ctrl L4
    (2, 32, 64, 64, 0, 64, 1, 2, 1),     # ctrl L5
    (3, 16, 32, 64, 0, 64, 1, 1, 1),     # ctrl L6
    (3, 16, 32, 64, 0, 64, 1, 2, 1),     # ctrl L7
    (2, 48, 48, 13, 0, 16, 1, 1, 1),     # attn L0 (Cin not a multiple of 4)
    (2, 48, 48, 16, 0, 32, 1, 2, 1),     # attn L1
    (2, 24, 24, 32, 0, 64, 1, 2, 1),     # attn L3
    (2, 12, 12, 64, 0, 96, 1, 2, 1),     # attn L5 (N = 96)
    (2, 6, 6, 96, 0, 64, 2, 1, 1),       # dcnn L0 (stride-2 transposed)
    (2, 12, 12, 64, 64, 64, 1, 1, 1),    # dcnn L1 (skip concat)
    (2, 12, 12, 64, 64, 32, 2, 1, 1),    # dcnn L2
    (2, 24, 24, 32, 32, 16, 2, 1, 1),    # dcnn L4
    (2, 48, 48, 16, 13, 1, 1, 1, 1),     # dcnn L6 (Cout = 1, odd skip)
    (1, 34, 70, 8, 0, 8, 1, 2, 0),       # ragged tiles, Cout = 8 (CVPPP-arch)
    (1, 512, 1024, 16, 0, 16, 1, 2, 1),  # Cityscapes-size ctrl L1
    (32, 16, 32, 64, 0, 64, 1, 2, 1),    # ctrl L7 at the bench batch (N split over CTAs)
    (32, 12, 12, 64, 64, 64, 1, 1, 1),   # dcnn L1 at the bench batch
    (7, 12, 12, 64, 0, 96, 1, 2, 1),     # odd batch, N = 96
    (5, 64, 128, 32, 0, 32, 1, 2, 1),    # several tiles per CTA (persistent loop, TMEM double buffering)
    (2, 48, 48, 16, 0, 16, 1, 1, 1),     # attn L0 on the glimpse padded to 16 channels (TMA feed)
    (2, 48, 48, 16, 16, 1, 1, 1, 1),     # dcnn L6 with the padded glimpse as skip input
    (32, 24, 24, 32, 32, 16, 2, 1, 1),   # dcnn L4 at the bench batch (TMA box of the low-resolution input)
    (1, 17, 22, 8, 4, 8, 2, 1, 0),       # transposed, odd height, ragged tiles
    (3, 10, 6, 4, 0, 20, 1, 1, 1),       # tile wider than the image (TMA box larger than the tensor)
]


@pytest.mark.parametrize('case', [UMMA_CASES[i] for i in (0, 4, 12, 13, 14, 21)])
def test_conv3x3_block_umma_plain_loads(ops, case, monkeypatch):
  """The same layers with the TMA feed switched off (the path layers with odd channel counts take)."""
  monkeypatch.setenv('RA_CONV_NO_TMA', '1')
  _run_umma_case(ops, case)


@pytest.mark.parametrize('case', UMMA_CASES)
def test_conv3x3_block_umma(ops, case):
  _run_umma_case(ops, case)


def _run_umma_case(ops, case):
  """One layer against the fp32 convolution of the oracle; returns the layout flags of the plan it ran."""
  B, H, W, C1, C2, Cout, up, pool, relu = case
  rng = np.random.default_rng(hash(case) % 2**31)
  x1 = rng.standard_normal((B, H, W, C1)).astype(np.float32)
  x2 = rng.standard_normal((B, H, W, C2)).astype(np.float32) if C2 else None
  Cin = C1 + C2
  scale = rng.uniform(0.5, 1.5, Cout).astype(np.float32)
  shift = rng.standard_normal(Cout).astype(np.float32)
  xin = torch.from_numpy(x1 if x2 is None else np.concatenate([x1, x2], 3))
  if up == 1:
    w = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    ref = OM.conv2d_same(xin, torch.from_numpy(w), torch.zeros(Cout))
  else:
    from rec_attend_b200.full_model import _deconv_to_conv
    wt = (rng.standard_normal((3, 3, Cout, Cin)) / np.sqrt(9 * Cin)).astype(np.float32)
    ref = OM.conv2d_transpose_same(xin, torch.from_numpy(wt), torch.zeros(Cout), up)
    w = _deconv_to_conv(wt)
  ref = ref * torch.from_numpy(scale) + torch.from_numpy(shift)
  if relu:
    ref = torch.relu(ref)
  if pool == 2:
    ref = OM.max_pool_same(ref, 2)
  KC, NPc, nsp, nch, rs = ops.umma_plan(Cin, Cout, H * up, W * up, pool, B, C2=C2)
  if rs & 2:  # fp16 hi / lo plan (test_conv3x3_block_umma_f16): the image is made on the device
    wp = ops.umma_filter_image(w, KC, NPc, nsp, rs, 'cuda')
    assert wp.numel() == nsp * nch * 9 * KC * NPc
  else:
    wp = ops.pack_umma_weights(w, KC, NPc, nsp, rs)
    assert wp.shape == (nsp, nch, 9, KC // 4, 2 * NPc, 4)
    wp = _g(wp)
  out = ops.conv3x3_block_umma(_g(x1), wp, Cout, _g(scale), _g(shift), pool=pool, relu=bool(relu),
                               x2=None if x2 is None else _g(x2), upsample=up)
  torch.cuda.synchronize()
  assert tuple(out.shape) == tuple(ref.shape)
  # 3xTF32 / fp16 hi + lo: ~2^-21 per product; 1e-5 of the tensor scale leaves margin
  assert rel_err(out.cpu().numpy(), ref.numpy()) < 1e-5
  return rs


@pytest.mark.parametrize('mode', [1, 2, 5])
@pytest.mark.parametrize('case', UMMA_CASES)
def test_conv3x3_block_umma_f16(ops, case, mode):
  """The fp16 hi / lo operand split (kind::f16, K = 16 per instruction; ra_conv3x3_umma_set_f16 / RA_UMMA_F16): same
  layers, same bound - hi = fp16(x) and lo' = fp16((x - hi) * 2^11) carry 11 + 11 significand bits like the two tf32 parts.
  Mode 5 = mode 1 + the correction half folded into the main half by the tensor core (A operand from tensor memory).
  Layers whose plan does not take the split (wide N, channel counts the TMA feed cannot serve) run as before."""
  B, H, W, C1, C2, Cout, up, pool, relu = case
  prev = ops.umma_set_f16(mode)
  try:
    rs = _run_umma_case(ops, case)
  finally:
    ops.umma_set_f16(prev)
  feedable = (C1 % 4) == 0 and (C2 % 4) == 0 and (C2 == 0 or C1 % 16 == 0)  # 16-channel TMA boxes
  if not feedable:
    assert not rs & 2, 'the fp16 split was planned for a layer the TMA feed cannot serve'
  if (mode & 3) == 2 and Cout <= 64 and (C1 + C2) >= 16 and feedable:
    assert rs & 2, 'mode 2 must take the fp16 split on every merged-mode layer with >= 16 input channels'


def test_umma_pack_f16(ops):
  """ra_umma_pack_f16 against its numpy model, bit for bit."""
  from rec_attend_b200 import params as PM
  rng = np.random.default_rng(5)
  v = (rng.standard_normal((2, 3, 9, 8, 24, 4)) * np.exp(rng.uniform(-12, 3, (2, 3, 9, 8, 24, 4)))).astype(np.float32)
  img = ops.umma_pack_f16(_g(v), 32, 24).cpu().numpy()
  ref = PM.pack_umma_f16_reference(v)
  assert np.array_equal(img.view(np.uint16).reshape(ref.shape), ref.view(np.uint16))


@pytest.mark.parametrize('bulk', [True, False])
@pytest.mark.parametrize('C0,pool,H,W,B', [(16, 2, 32, 64, 2), (8, 1, 24, 40, 2), (16, 1, 16, 16, 2), (8, 2, 64, 32, 2),
                                           (16, 2, 16, 320, 2),    # 3 segments per row, ragged last one
                                           (8, 2, 8, 520, 1),      # C0 = 8: 128 pooled pixels per item, 4-pixel tail
                                           (16, 1, 8, 300, 2),     # no pooling, 5 segments
                                           (16, 2, 256, 512, 3)])  # > 4 items per persistent CTA: the ring wraps
def test_canvas_conv(ops, C0, pool, H, W, B, bulk, monkeypatch):
  """Both kernels of ra_canvas_conv_f32: the cp.async.bulk pipeline (default) and the plain-load one."""
  if not bulk:
    monkeypatch.setenv('RA_CANVAS_NO_BULK', '1')
  rng = np.random.default_rng(C0 * 10 + pool)
  pre = rng.standard_normal((B, H, W, C0)).astype(np.float32)
  canvas = rng.random((B, H, W)).astype(np.float32)
  w = (rng.standard_normal((3, 3, 1, C0)) / 3).astype(np.float32)
  scale = rng.uniform(0.5, 1.5, C0).astype(np.float32)
  shift = rng.standard_normal(C0).astype(np.float32)
  ref = OM.conv2d_same(torch.from_numpy(canvas).unsqueeze(3), torch.from_numpy(w), torch.zeros(C0))
  ref = torch.relu((ref + torch.from_numpy(pre)) * torch.from_numpy(scale) + torch.from_numpy(shift))
  if pool == 2:
    ref = OM.max_pool_same(ref, 2)
  out = ops.canvas_conv(_g(pre), _g(canvas), _g(w), _g(scale), _g(shift), pool=pool)
  assert rel_err(out.cpu().numpy(), ref.numpy()) < TOL


# ----------------------------------------------------------------------------- post-processing
def _pp_golden():
  import os
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'postprocess_golden.npz'))
  for i in range(int(g['n_cases'])):
    p = 'c%d_' % i
    yield (g[p + 'y_out'], g[p + 's_out'], g[p + 'fg'] if p + 'fg' in g.files else None, float(g[p + 'thresh']),
           int(g[p + 'tiny']), g[p + 'dense'], g[p + 'conf'], g[p + 'area'])


def test_postprocess_reference_golden(ops):
  """ra_postprocess_f32 against outputs of the reference's own utils/postprocess.py (tests/golden)."""
  from rec_attend_b200 import postprocess as PP
  from oracle import postprocess as OP
  for y, s, fg, thresh, tiny, dense, conf, area in _pp_golden():
    out = PP.postprocess(_g(y), _g(s), thresh, fg=None if fg is None else _g(fg), remove_tiny=tiny, want_dense=True)
    assert (out['y_out_thresh'].cpu().numpy() == dense).all()
    assert (out['conf'].cpu().numpy() == conf).all()
    assert rel_err(out['area'].cpu().numpy(), area) < 1e-6
    binary_fg = fg is None or set(np.unique(fg)) <= {0.0, 1.0}
    if binary_fg:
      assert (out['area'].cpu().numpy() == area).all()
    assert (out['label'].cpu().numpy() == OP.label_map(dense)).all()


@pytest.mark.parametrize('B,T,H,W,tiny', [(3, 20, 64, 96, 0), (2, 32, 48, 64, 30), (4, 7, 30, 44, 15), (1, 64, 16, 16, 0)])
def test_postprocess_vs_oracle(ops, B, T, H, W, tiny):
  from rec_attend_b200 import postprocess as PP
  from oracle import postprocess as OP
  rng = np.random.default_rng(B * 100 + T)
  y = rng.random((B, T, H, W)).astype(np.float32)**3
  y[:, T // 2] = y[:, 0]  # exact ties: the first maximum must win
  s = rng.uniform(0.3, 1.0, (B, T)).astype(np.float32)
  s[:, T // 2] = s[:, 0]
  fg = (rng.random((B, H, W)) > 0.2).astype(np.float32)
  for use_fg in (False, True):
    dense, conf, area = OP.eval_chain(y, s, 0.3, fg if use_fg else None, tiny)
    out = PP.postprocess(_g(y), _g(s), 0.3, fg=_g(fg) if use_fg else None, remove_tiny=tiny, want_dense=True)
    assert (out['label'].cpu().numpy() == OP.label_map(dense)).all()
    assert (out['y_out_thresh'].cpu().numpy() == dense).all()
    assert (out['conf'].cpu().numpy() == conf).all() and (out['area'].cpu().numpy() == area).all()
    lab_only = PP.postprocess(_g(y), _g(s), 0.3, fg=_g(fg) if use_fg else None, remove_tiny=tiny)
    assert (lab_only['label'].cpu().numpy() == OP.label_map(dense)).all()


def test_postprocess_full_size_properties(ops):
  """BASELINE config-3 size (T=20, 256x512): size-independent properties instead of the (slow) numpy oracle."""
  from rec_attend_b200 import postprocess as PP
  B, T, H, W = 4, 20, 256, 512
  gen = torch.Generator(device='cuda').manual_seed(5)
  y = torch.rand((B, T, H, W), device='cuda', generator=gen)
  s = torch.rand((B, T), device='cuda', generator=gen) * 0.6 + 0.4
  out = PP.postprocess(y, s, 0.3, remove_tiny=0, want_dense=True)
  lab, dense, area = out['label'], out['y_out_thresh'], out['area']
  v = y * s.view(B, T, 1, 1)
  ref_lab = torch.where(v.max(dim=1).values.double() > 0.3, v.argmax(dim=1) + 1, torch.zeros_like(lab, dtype=torch.int64))
  # torch.argmax does not promise the first maximum on ties; random floats make ties vanishingly rare
  assert float((lab.long() == ref_lab).float().mean()) > 0.99999
  assert (dense.sum(dim=1) <= 1).all()                                         # one label per pixel
  assert (dense.sum(dim=(2, 3)) == area).all()                                  # areas are the dense sums
  assert int((lab > 0).sum()) == int(area.sum())                                # checksum of checksums
  onehot = torch.stack([(lab == t + 1) for t in range(T)], 1).float()
  assert (onehot == dense).all()
  # idempotence: the dense masks are a fixed point (confidence 1, same threshold)
  again = PP.postprocess(dense, torch.ones_like(s), 0.3, want_dense=True)
  assert (again['label'] == lab).all() and (again['y_out_thresh'] == dense).all()
  # removing everything below a huge threshold empties the map and zeroes the confidences
  none = PP.postprocess(y, s, 0.3, remove_tiny=H * W)
  assert int(none['label'].abs().sum()) == 0 and float(none['conf'].abs().sum()) == 0.0


def test_postprocess_errors(ops):
  from rec_attend_b200 import _lib, postprocess as PP
  y = torch.zeros((1, 65, 8, 8), device='cuda')
  with pytest.raises(_lib.RecAttendError):
    PP.postprocess(y, torch.zeros((1, 65), device='cuda'))          # T > 64
  with pytest.raises(_lib.RecAttendError):
    PP.postprocess(torch.zeros((1, 2, 3, 3), device='cuda'), torch.zeros((1, 2), device='cuda'))  # H*W % 4
  with pytest.raises(_lib.RecAttendError):
    PP.postprocess(torch.zeros((1, 2, 4, 4), device='cuda'), torch.zeros((1, 3), device='cuda'))
  out = PP.postprocess(torch.zeros((0, 2, 4, 4), device='cuda'), torch.zeros((0, 2), device='cuda'))
  assert tuple(out['label'].shape) == (0, 4, 4)


# ----------------------------------------------------------------------------- training-mode BN block
@pytest.mark.parametrize('B,H,W,C,pool', [(3, 16, 24, 16, 2), (2, 12, 12, 96, 2), (4, 48, 48, 16, 1), (1, 6, 10, 64, 1),
                                          (8, 128, 256, 16, 2), (3, 48, 48, 1, 1), (2, 10, 14, 6, 2)])
def test_batch_norm_train_block(ops, B, H, W, C, pool):
  """nnlib.batch_norm(phase_train=True) + ReLU + max-pool against the oracle (batch moments, EMA update)."""
  rng = np.random.default_rng(B * 100 + C)
  x = (rng.standard_normal((B, H, W, C)) * rng.uniform(0.5, 3.0, C) + rng.uniform(-4, 4, C)).astype(np.float32)
  p = {'gamma': rng.uniform(0.5, 1.5, C).astype(np.float32), 'beta': rng.standard_normal(C).astype(np.float32),
       'ema_mean': rng.standard_normal(C).astype(np.float32), 'ema_var': rng.uniform(0.5, 1.5, C).astype(np.float32)}
  pt = {k: torch.from_numpy(v) for k, v in p.items()}
  normed, mean, var, new_mean, new_var = OM.batch_norm_train(torch.from_numpy(x), pt)
  ref = torch.relu(normed)
  if pool == 2:
    ref = OM.max_pool_same(ref, 2)
  em, ev = _g(p['ema_mean'].copy()), _g(p['ema_var'].copy())
  y, bm, bv = ops.batch_norm_train_block(_g(x), _g(p['gamma']), _g(p['beta']), em, ev, pool=pool)
  assert rel_err(bm.cpu().numpy(), mean.numpy()) < 1e-5 and rel_err(bv.cpu().numpy(), var.numpy()) < 1e-4
  assert rel_err(em.cpu().numpy(), new_mean.numpy()) < 1e-5 and rel_err(ev.cpu().numpy(), new_var.numpy()) < 1e-4
  assert rel_err(y.cpu().numpy(), ref.numpy()) < 1e-4
  # without EMA buffers / without ReLU
  y2, _, _ = ops.batch_norm_train_block(_g(x), _g(p['gamma']), _g(p['beta']), None, None, pool=pool, relu=False)
  ref2 = OM.max_pool_same(normed, 2) if pool == 2 else normed
  assert rel_err(y2.cpu().numpy(), ref2.numpy()) < 1e-4


def test_conv_block_train_mode(ops):
  """A whole training-mode layer: tcgen05 conv + bias, batch-stat BN, ReLU, pool (nnlib.py:229-253)."""
  rng = np.random.default_rng(11)
  B, H, W, Cin, Cout, pool = 4, 32, 48, 16, 32, 2
  x = rng.standard_normal((B, H, W, Cin)).astype(np.float32)
  w = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
  b = (rng.standard_normal(Cout) * 0.1).astype(np.float32)
  p = {'gamma': rng.uniform(0.5, 1.5, Cout).astype(np.float32), 'beta': rng.standard_normal(Cout).astype(np.float32),
       'ema_mean': np.zeros(Cout, np.float32), 'ema_var': np.ones(Cout, np.float32)}
  KC, NPc, nsp, _, rs = ops.umma_plan(Cin, Cout, H, W, 1, B)
  wp = ops.umma_filter_image(w, KC, NPc, nsp, rs, 'cuda')
  y, bm, bv = ops.conv3x3_block_train(_g(x), wp, _g(b), _g(p['gamma']), _g(p['beta']), pool=pool)
  raw = OM.conv2d_same(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b))
  normed, mean, var, _, _ = OM.batch_norm_train(raw, {k: torch.from_numpy(v) for k, v in p.items()})
  ref = OM.max_pool_same(torch.relu(normed), 2)
  assert rel_err(bm.cpu().numpy(), mean.numpy()) < 1e-5 and rel_err(y.cpu().numpy(), ref.numpy()) < 1e-4


def test_batch_norm_train_errors(ops):
  from rec_attend_b200 import _lib
  with pytest.raises(_lib.RecAttendError):
    ops.batch_norm_train_block(torch.zeros((1, 4, 4, 300), device='cuda'), torch.ones(300, device='cuda'),
                               torch.zeros(300, device='cuda'))  # C > 256


# ----------------------------------------------------------------------------- augmentation
@pytest.mark.parametrize('H,W,pad,off,vf,hf,tr,with_d', [(16, 16, 4, (4, 4), False, False, False, False),
                                                         (16, 16, 4, (0, 7), True, False, True, False),
                                                         (12, 20, 3, (6, 1), True, True, False, False),
                                                         (24, 24, 8, (16, 0), False, True, True, False),
                                                         (12, 20, 5, (2, 9), False, False, False, True)])
def test_random_transformation(ops, H, W, pad, off, vf, hf, tr, with_d):
  rng = np.random.default_rng(H + W + pad)
  B, T = 2, 3
  x = rng.random((B, H, W, 3)).astype(np.float32)
  y = (rng.random((B, T, H, W)) > 0.5).astype(np.float32)
  d = rng.random((B, H, W, 8)).astype(np.float32) if with_d else None
  c = rng.random((B, H, W, 2)).astype(np.float32) if with_d else None
  ref = OM.random_transformation(torch.from_numpy(x), pad, off, vf, hf, tr, y=torch.from_numpy(y),
                                 d=None if d is None else torch.from_numpy(d), c=None if c is None else torch.from_numpy(c))
  out = ops.random_transformation(_g(x), pad, off, vf, hf, tr, y=_g(y), d=None if d is None else _g(d),
                                  c=None if c is None else _g(c))
  assert sorted(out) == sorted(ref)
  for k in ref:
    assert (out[k].cpu().numpy() == ref[k].numpy()).all(), k
  if off == (pad, pad) and not (vf or hf or tr):
    assert (out['x'].cpu().numpy() == x).all()  # the eval-mode identity (image_ops.py:70-80,106-112)


def test_random_transformation_errors(ops):
  from rec_attend_b200 import _lib
  x = torch.zeros((1, 8, 12, 3), device='cuda')
  with pytest.raises(_lib.RecAttendError):
    ops.random_transformation(x, 2, (1, 1), transpose=True)            # H != W
  with pytest.raises(_lib.RecAttendError):
    ops.random_transformation(x, 2, (5, 0))                            # offset > 2*padding
  with pytest.raises(_lib.RecAttendError):
    ops.random_transformation(x, 2, (1, 1), hflip=True, d=torch.zeros((1, 8, 12, 8), device='cuda'))


def test_rowstack_mode_matches_fp32_conv(cuda):
  """The opt-in row-stacked-taps mode of the tcgen05 convolution (RA_UMMA_ROWSTACK=2: three kx taps share one read of
  the A operand, lanes combined in the epilogue) against the CUDA-core fp32 convolution on every narrow KITTI layer
  shape, pooled / transposed / skip-concatenated / odd-sized.  The mode is read once per process: subprocess."""
  import os
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  env = dict(os.environ, RA_UMMA_ROWSTACK='2')
  out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'rowstack_ab.py')], env=env, capture_output=True,
                       text=True, timeout=600)
  assert out.returncode == 0, out.stderr[-2000:]
  rows = [l for l in out.stdout.splitlines() if l.startswith('mode 2 ') and ' err ' in l]
  assert len(rows) >= 12 and sum("'rowstack': 1" in l for l in rows) >= 10
  for l in rows:
    assert float(l.split(' err ')[1].split()[0]) < 1e-5, l


@pytest.mark.parametrize('B,T,H,W', [(3, 5, 16, 32), (7, 20, 32, 64), (2, 32, 64, 64), (32, 20, 64, 128), (1, 127, 8, 32)])
def test_iou_soft_hard_tensor_core(cuda, B, T, H, W):
  """ra_pairwise_iou_umma_f32 (tcgen05 K-major GEMM over the pixels, hi / lo split of the soft masks) against the
  oracle's f_iou_pairwise / f_dice_pairwise and against the CUDA-core kernel."""
  from rec_attend_b200 import ops
  rng = np.random.default_rng(B * 100 + T)
  a = rng.random((B, T, H, W)).astype(np.float32) ** 3
  g = (rng.random((B, T, H, W)) > 0.7).astype(np.float32)
  g[:, -1] = 0.0  # an empty ground-truth slot
  res = ops.f_iou_soft_hard(_g(a), _g(g), hard_threshold=0.5)
  assert res is not None
  soft, hard, dice = [t.cpu().numpy() for t in res]
  at, gt = torch.from_numpy(a), torch.from_numpy(g)
  ref_soft = OM.f_iou_pairwise(at, gt).numpy()
  ah = (at > 0.5).float()
  ref_hard = OM.f_iou_pairwise(ah, gt).numpy()
  ref_dice = OM.f_dice_pairwise(ah, gt).numpy()
  assert rel_err(soft, ref_soft) < 2e-5
  assert rel_err(hard, ref_hard) < 2e-5
  assert rel_err(dice, ref_dice) < 2e-5
  if T <= 64:
    old = ops.f_iou(_g(a), _g(g)).cpu().numpy()
    assert rel_err(soft, old) < 2e-5
  # H*W % 32 != 0: outside the kernel's range, the caller falls back to f_iou
  assert ops.f_iou_soft_hard(_g(a[:, :, :H - 1, :W - 1].copy()), _g(g[:, :, :H - 1, :W - 1].copy())) is None


def test_conv_chain_equals_layer_by_layer(cuda, request):
  """ra_conv3x3_umma_chain_*: the six attention-CNN layers and the first deconv layers of the KITTI patch network (skip
  connections, transposed conv, pooling) in ONE persistent launch with grid barriers between the layers == the same
  layers launched one by one (bit for bit: the same tile plans and kernels)."""
  from rec_attend_b200 import ops
  prev_f16 = ops.umma_set_f16(0)  # the chain kernel runs the 3xTF32 variant only
  request.addfinalizer(lambda: ops.umma_set_f16(prev_f16))
  rng = np.random.default_rng(11)
  B = 8
  spec = [  # C1, C2 (skip = index of an earlier output or None), Cout, up, pool, size_in
      (16, None, 16, 1, 1), (16, None, 32, 1, 2), (32, None, 32, 1, 1), (32, None, 64, 1, 2), (64, None, 64, 1, 1),
      (64, None, 96, 1, 2), (96, None, 64, 2, 1), (64, 4, 64, 1, 1), (64, 3, 32, 2, 1), (32, 2, 32, 1, 1)]
  x0 = _g(rng.standard_normal((B, 48, 48, 16)).astype(np.float32))
  acts = [x0]
  layers = []
  for C1, skip, Cout, up, pool in spec:
    x = acts[-1]
    x2 = acts[skip + 1] if skip is not None else None
    Bx, H, W, _ = x.shape
    Cin = C1 + (0 if x2 is None else x2.shape[3])
    assert x.shape[3] == C1 and (x2 is None or tuple(x2.shape[1:3]) == (H, W))
    KC, NPc, nsp, _, rs = ops.umma_plan(Cin, Cout, H * up, W * up, pool, B)
    w = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    out = torch.empty((B, H * up // pool, W * up // pool, Cout), device='cuda')
    layers.append({'x': x, 'x2': x2, 'wpack': _g(ops.pack_umma_weights(w, KC, NPc, nsp, rs)), 'Cout': Cout,
                   'scale': _g(rng.uniform(0.5, 1.5, Cout).astype(np.float32)),
                   'shift': _g(rng.standard_normal(Cout).astype(np.float32) * 0.1), 'pool': pool, 'relu': True,
                   'upsample': up, 'out': out})
    acts.append(out)
  chain = ops.ConvChain(layers)
  chain.run()
  torch.cuda.synchronize()
  got = [L['out'].clone() for L in layers]
  for L in layers:
    L['out'].zero_()
  for L in layers:
    ops.conv3x3_block_umma(L['x'], L['wpack'], L['Cout'], L['scale'], L['shift'], pool=L['pool'], relu=True, x2=L['x2'],
                           upsample=L['upsample'], out=L['out'])
  torch.cuda.synchronize()
  for i, L in enumerate(layers):
    assert torch.equal(got[i], L['out']), i
  chain.run()  # a second run (barrier counter reset) gives the same again
  torch.cuda.synchronize()
  assert torch.equal(got[-1], layers[-1]['out'])


@pytest.mark.gpu
@pytest.mark.parametrize('n_in,n_out,R,bias', [(64, 256, 3200, True), (256, 256, 3200, False), (320, 9, 640, True),
                                               (17, 130, 100, True), (256, 1, 640, True), (64, 64, 63, False)])
def test_outer_sum_row_chunks(cuda, n_in, n_out, R, bias):
  """ra_outer_sum_ex_f32 (rows split over chunks of CTAs, partials added in chunk order) against a float64 product,
  against the single-chunk entry, and bit-identical from run to run."""
  from rec_attend_b200 import _lib, ops, train as TR
  g = torch.Generator().manual_seed(n_in * 1000 + n_out + R)
  a_stride, d_stride = n_in + 3, n_out + 5  # rows live inside wider records, like the controller tape
  A = torch.randn((R, a_stride), generator=g).cuda()
  D = torch.randn((R, d_stride), generator=g).cuda()
  ref = A[:, :n_in].double().t() @ D[:, :n_out].double()
  dW = torch.empty((n_in, n_out), device='cuda')
  db = torch.empty((n_out,), device='cuda') if bias else None
  TR._outer_sum(A, a_stride, n_in, D, d_stride, n_out, R, dW, db)
  assert rel_err(dW.cpu().numpy(), ref.cpu().numpy()) < 1e-5
  if bias:
    assert rel_err(db.cpu().numpy(), D[:, :n_out].double().sum(0).cpu().numpy()) < 1e-5
  dW2 = torch.empty_like(dW)
  TR._outer_sum(A, a_stride, n_in, D, d_stride, n_out, R, dW2, None)
  assert torch.equal(dW, dW2)
  one = torch.empty_like(dW)
  _lib.call('ra_outer_sum_f32', ops._p(A), a_stride, n_in, ops._p(D), d_stride, n_out, R, ops._p(one), ops._p(None),
            ops._stream())
  assert rel_err(one.cpu().numpy(), ref.cpu().numpy()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize('B,T,H,W', [(2, 20, 64, 128), (3, 32, 40, 72), (2, 40, 32, 64), (1, 7, 33, 50)])
def test_box_gt_step_many_boxes(ops, B, T, H, W):
  """ra_box_gt_step_f32 with up to 32 GT boxes (single-pass kernel: per-column / per-row rectangle masks, sums in
  registers), more than 32 (multi-pass fallback) and ragged sizes, against the oracle's f_inter / f_union on the filled
  GT boxes."""
  rng = np.random.default_rng(100 + T)
  y_gt, _ = _masks(rng, B, T, H, W)
  tl, br, boxgt, rect, _ = ops.get_gt_box(_g(y_gt), padding_ratio=0.2, min_padding=4.0)
  attn = rng.random((B, 1, H, W)).astype(np.float32)
  canvas = torch.zeros((B, H, W), device='cuda')
  iou_t = torch.zeros((B, T), device='cuda')
  grd = torch.zeros((B, T), device='cuda')
  ops.box_gt_step(_g(attn), H * W, rect, _g(y_gt), None, 0, iou_t, T, grd, canvas)
  a = torch.from_numpy(attn)
  bg = boxgt.cpu()
  ref_iou = OM.f_inter(a, bg) / OM.f_union(a, bg)
  assert rel_err(iou_t.cpu().numpy(), ref_iou.numpy()) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize('B,H,W,Cout,bias', [(3, 16, 24, 16, True), (2, 9, 11, 8, False), (5, 32, 64, 16, True),
                                             (2, 7, 13, 64, True), (2, 8, 8, 6, True)])
def test_bwd_weight_single_input_channel(cuda, B, H, W, Cout, bias):
  """Weight gradient with ONE input channel (the canvas channel of the first controller layer): the bandwidth-bound
  register kernel (Cout % 4 == 0; ragged widths, image borders) and the tiled fallback (Cout = 6) against autograd."""
  from rec_attend_b200 import ops
  rng = np.random.default_rng(B * 100 + W)
  x = rng.standard_normal((B, H, W, 1)).astype(np.float32)
  g = rng.standard_normal((B, H, W, Cout)).astype(np.float32)
  xt, wt = torch.from_numpy(x), torch.randn(3, 3, 1, Cout, requires_grad=True)
  bt = torch.zeros(Cout, requires_grad=True)
  y = OM.conv2d_same(xt, wt, bt)
  gw, gb = torch.autograd.grad(y, [wt, bt], torch.from_numpy(g))
  dw, db = ops.conv3x3_bwd_weight(_g(x), _g(g), want_db=bias)
  assert rel_err(dw.cpu().numpy(), gw.numpy()) < 1e-5
  if bias:
    assert rel_err(db.cpu().numpy(), gb.numpy()) < 1e-5
  dw2, _ = ops.conv3x3_bwd_weight(_g(x), _g(g), want_db=False)
  assert torch.equal(dw, dw2)  # fixed summation order


@pytest.mark.gpu
@pytest.mark.parametrize('C', [4, 8, 16, 32, 64, 128, 96, 12, 1])
def test_batch_norm_train_channel_counts(ops, C):
  """Batch moments for every channel-count class of the partial reduction (float4 + shuffle path for C / 4 dividing 32,
  the general path for 96 / 12 / 1 channels) against float64 moments."""
  rng = np.random.default_rng(C)
  B, H, W = 3, 20, 28
  x = (rng.standard_normal((B, H, W, C)) * 2.0 + rng.standard_normal(C) * 3.0).astype(np.float32)
  gamma = rng.uniform(0.5, 1.5, C).astype(np.float32)
  beta = rng.standard_normal(C).astype(np.float32)
  y, mean, var = ops.batch_norm_train_block(_g(x), _g(gamma), _g(beta), pool=1, relu=False)
  x64 = x.astype(np.float64).reshape(-1, C)
  m64, v64 = x64.mean(0), x64.var(0)
  assert rel_err(mean.cpu().numpy(), m64) < 1e-5
  assert rel_err(var.cpu().numpy(), v64) < 1e-5
  ref = (x64 - m64) / np.sqrt(v64 + 1e-3) * gamma + beta
  assert rel_err(y.cpu().numpy().reshape(-1, C), ref) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('B,H,W,C,pool', [(32, 48, 48, 16, 1), (4, 24, 24, 32, 2), (3, 12, 12, 64, 1), (2, 16, 32, 64, 2),
                                          (5, 6, 6, 8, 1)])
def test_batch_norm_fused_small_map_kernel(ops, monkeypatch, B, H, W, C, pool):
  """The opt-in one-launch batch norm (RA_BN_FUSED=1: slice resident in shared memory, two grid barriers) against the
  three-launch path and float64 moments; EMA shadows moved alike."""
  rng = np.random.default_rng(B * 1000 + C)
  x = (rng.standard_normal((B, H, W, C)) * 1.5 + rng.standard_normal(C)).astype(np.float32)
  gamma = rng.uniform(0.5, 1.5, C).astype(np.float32)
  beta = rng.standard_normal(C).astype(np.float32)
  em0, ev0 = rng.standard_normal(C).astype(np.float32), rng.uniform(0.5, 1.5, C).astype(np.float32)
  res = {}
  for mode in ('three', 'fused'):
    if mode == 'fused':
      monkeypatch.setenv('RA_BN_FUSED', '1')
    else:
      monkeypatch.delenv('RA_BN_FUSED', raising=False)
    em, ev = _g(em0.copy()), _g(ev0.copy())
    y, mean, var = ops.batch_norm_train_block(_g(x), _g(gamma), _g(beta), em, ev, pool=pool, relu=True)
    torch.cuda.synchronize()
    res[mode] = [t.cpu().numpy() for t in (y, mean, var, em, ev)]
  for a, b in zip(res['three'], res['fused']):
    assert rel_err(b, a) < 1e-5
  x64 = x.astype(np.float64).reshape(-1, C)
  assert rel_err(res['fused'][1], x64.mean(0)) < 1e-5 and rel_err(res['fused'][2], x64.var(0)) < 1e-5
