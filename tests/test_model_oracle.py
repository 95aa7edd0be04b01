"""CPU semantic tests of the model oracle (oracle/model.py).  TensorFlow 0.12 cannot run here; the
oracle is pinned to the reference's own source through the shim-executed golden vectors (tests/test_*_golden.py), and
every TF-specific kernel semantic that pinning takes on trust is checked here against an independent
numpy formulation."""
import numpy as np
import pytest
import torch

from oracle import model as OM


def _conv_same_numpy(x, w):
  B, H, W, Ci = x.shape
  Co = w.shape[3]
  xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
  y = np.zeros((B, H, W, Co), np.float64)
  for ky in range(3):
    for kx in range(3):
      y += np.einsum('bhwc,cd->bhwd', xp[:, ky:ky + H, kx:kx + W, :], w[ky, kx])
  return y


def test_conv_same_padding_is_symmetric_one():
  rng = np.random.default_rng(0)
  x = rng.standard_normal((2, 6, 8, 3)).astype(np.float32)
  w = rng.standard_normal((3, 3, 3, 5)).astype(np.float32)
  y = OM.conv2d_same(torch.from_numpy(x), torch.from_numpy(w), torch.zeros(5)).numpy()
  assert np.abs(y - _conv_same_numpy(x, w)).max() < 1e-4


def _conv_transpose_numpy(x, w, stride):
  """Gradient of a SAME stride-s 3x3 conv w.r.t. its input == tf.nn.conv2d_transpose (nnlib.py:372-376).
  Forward SAME padding for k=3: s=1 -> 1 before; s=2 (even input) -> 0 before, 1 after."""
  B, M, N, Ci = x.shape
  Co = w.shape[2]
  H, W = M * stride, N * stride
  pad_before = 1 if stride == 1 else 0
  y = np.zeros((B, H, W, Co), np.float64)
  for iy in range(M):
    for ix in range(N):
      for ky in range(3):
        for kx in range(3):
          oy, ox = iy * stride + ky - pad_before, ix * stride + kx - pad_before
          if 0 <= oy < H and 0 <= ox < W:
            y[:, oy, ox, :] += x[:, iy, ix, :] @ w[ky, kx].T  # w [kh,kw,Cout,Cin]
  return y


def test_conv2d_transpose_crops_at_the_end():
  rng = np.random.default_rng(1)
  for stride in (1, 2):
    x = rng.standard_normal((2, 4, 5, 3)).astype(np.float32)
    w = rng.standard_normal((3, 3, 4, 3)).astype(np.float32)
    y = OM.conv2d_transpose_same(torch.from_numpy(x), torch.from_numpy(w), torch.zeros(4), stride).numpy()
    assert y.shape == (2, 4 * stride, 5 * stride, 4)
    assert np.abs(y - _conv_transpose_numpy(x, w, stride)).max() < 1e-4


def test_deconv_as_conv_over_zero_inserted_input():
  """The identity the CUDA path relies on: conv2d_transpose(stride s) == SAME-style conv with the
  flipped/swapped filter over the zero-inserted input, window starting s before the output pixel."""
  rng = np.random.default_rng(2)
  for s in (1, 2):
    x = rng.standard_normal((1, 3, 4, 2)).astype(np.float32)
    wt = rng.standard_normal((3, 3, 5, 2)).astype(np.float32)
    ref = _conv_transpose_numpy(x, wt, s)
    wc = wt[::-1, ::-1].transpose(0, 1, 3, 2)  # [k,k,Cin,Cout]
    H, W = 3 * s, 4 * s
    v = np.zeros((1, H, W, 2))
    v[:, ::s, ::s] = x
    vp = np.pad(v, ((0, 0), (s, 2), (s, 2), (0, 0)))
    y = np.zeros((1, H, W, 5))
    for ky in range(3):
      for kx in range(3):
        y += np.einsum('bhwc,cd->bhwd', vp[:, ky:ky + H, kx:kx + W], wc[ky, kx])
    assert np.abs(y - ref).max() < 1e-5


def test_batch_norm_eval_uses_ema_and_eps_1e3():
  x = torch.tensor([[[[2.0, -1.0]]]])
  p = {'gamma': torch.tensor([2.0, 0.5]), 'beta': torch.tensor([0.1, -0.2]), 'ema_mean': torch.tensor([1.0, 0.0]),
       'ema_var': torch.tensor([4.0, 1.0])}
  y = OM.batch_norm_eval(x, p).numpy().ravel()
  ref = (np.array([2.0, -1.0]) - [1.0, 0.0]) / np.sqrt(np.array([4.0, 1.0]) + 1e-3) * [2.0, 0.5] + [0.1, -0.2]
  assert np.abs(y - ref).max() < 1e-6


def test_gaussian_filter_formula():  # SURVEY §10, modellib.py:581-612
  f = OM.get_gaussian_filter(torch.tensor([10.0]), torch.tensor([23.0]), torch.tensor([0.5]), 32, 48).numpy()[0]
  mu = 10.0 + 24.0 / 48 * (np.arange(48) - 23.5)
  var = np.exp(0.5)
  ref = np.exp(-0.5 * (np.arange(32)[:, None] - mu[None])**2 / var) / np.sqrt(var) / np.sqrt(2 * np.pi)
  assert f.shape == (32, 48) and np.abs(f - ref).max() < 1e-6


def test_extract_patch_is_FyT_X_Fx():
  rng = np.random.default_rng(3)
  x = rng.random((2, 7, 9, 3)).astype(np.float32)
  fy = rng.random((2, 7, 4)).astype(np.float32)
  fx = rng.random((2, 9, 5)).astype(np.float32)
  p = OM.extract_patch(torch.from_numpy(x), torch.from_numpy(fy), torch.from_numpy(fx), 3).numpy()
  ref = np.einsum('byi,byxd,bxj->bijd', fy, x, fx)
  assert np.abs(p - ref).max() < 1e-4


def test_union_eps_is_per_pixel():
  a = torch.zeros(1, 1, 4, 5)
  b = torch.zeros(1, 1, 4, 5)
  assert abs(float(OM.f_union(a, b)) - 20 * 1e-5) < 1e-9
  iou = OM.f_iou_pairwise(torch.ones(1, 2, 4, 5), torch.ones(1, 3, 4, 5))
  assert iou.shape == (1, 2, 3) and abs(float(iou[0, 0, 0]) - 20 / (20 + 20e-5)) < 1e-6


def test_gt_box_padding_and_empty_mask():  # modellib.py:663-701
  y = torch.zeros(1, 2, 40, 50)
  y[0, 0, 10:21, 5:31] = 1  # rows 10..20, cols 5..30
  tl, br, box = OM.get_gt_box(y, padding_ratio=0.2, min_padding=4.0)
  # size = (10, 25); pad = max(0.2*size, 4) = (4, 5)
  assert tl[0, 0].tolist() == [6.0, 0.0] and br[0, 0].tolist() == [24.0, 35.0]
  assert float(box[0, 0].sum()) == (24 - 6 + 1) * (35 - 0 + 1)
  assert tl[0, 1].tolist() == [0.0, 0.0] and br[0, 1].tolist() == [8.0, 8.0]  # empty mask -> (0,0)-(2*min_pad)
  assert float(box[0, 1].sum()) == 0.0  # the filled box is drawn BEFORE the fix-up


def test_segm_match_rounding_and_masking():  # modellib.py:395-415, SURVEY §9.14
  iou = torch.tensor([[[0.4999995, 0.2], [0.1, 0.3]]])
  s_gt = torch.tensor([[1.0, 0.0]])
  w = OM.segm_match_weights(iou, s_gt).numpy()
  assert abs(w[0, 0, 0] - (0.5 + 1e-5)) < 1e-7  # floor(x*1e6 + 0.5): half-up
  assert np.allclose(w[0, :, 1], 1e-5) and np.allclose(w[0, 1, :], 1e-5)
  m = OM.f_segm_match(iou, s_gt).numpy()
  assert m.sum() == 1 and m[0, 0, 0] == 1


def test_conf_loss_cummin_cummax():  # modellib.py:316-339
  s = torch.tensor([[0.9, 0.2, 0.6]])
  match = torch.zeros(1, 3, 3)
  match[0, 0, 1] = 1
  ref = -(np.log(0.9 + 1e-5)) - np.log(1 - 0.6 + 1e-5) - np.log(1 - 0.6 + 1e-5)
  assert abs(float(OM.f_conf_loss(s, match)) - ref / 3) < 1e-6


def test_full_forward_shapes_and_skip_wiring():
  import rec_attend_b200 as ra
  for arch, nsc in (('kitti', 1), ('cityscapes', 9), ('cvppp', 1)):
    opt = ra.config.full_model_opt(arch, 32, 64, 3)
    batch = ra.synthetic.make_batch(opt, 2)
    w = ra.synthetic.make_weights(opt)
    D = 4 if arch == 'cvppp' else 12 + nsc
    assert w['attn_cnn_w_0'].shape[2] == D
    if arch != 'cvppp':  # every deconv layer >= 1 has a skip
      assert w['attn_dcnn_w_6'].shape == (3, 3, 1, 16 + D)
      assert w['attn_dcnn_w_1'].shape == (3, 3, 64, 64 + 64)
    else:
      assert w['attn_dcnn_w_6'].shape == (3, 3, 1, 8)
    m = OM.full_model_forward(opt, w, batch)
    assert m['y_out'].shape == (2, 3, 32, 64) and m['x_patch'].shape == (2, 3, 48, 48, D)
    assert m['ctrl_rnn_glimpse_map'].shape == (2, 3, 5, 1, 2)
    assert torch.isfinite(m['loss'])
    # masks can only rise from sigmoid(-5) and the canvas is their running max
    assert float(m['y_out'].min()) >= 0.0066
    assert torch.allclose(m['canvas'][..., 0], m['y_out'].max(dim=1)[0])


def test_box_model_forward_shapes():
  import rec_attend_b200 as ra
  opt = ra.config.box_model_opt(32, 64, 3)
  batch = ra.synthetic.make_batch(opt, 2)
  w = ra.synthetic.make_weights(opt, model='box')
  m = OM.box_model_forward(opt, w, batch)
  assert m['attn_box'].shape == (2, 3, 32, 64) and m['s_out'].shape == (2, 3)
  assert m['match_box'].shape == (2, 3, 3) and torch.isfinite(m['loss'])


def test_batch_norm_train_matches_torch_and_ema_rule():
  """oracle.batch_norm_train (nnlib.py:65-128, phase_train=True): biased batch moments, eps 1e-3, EMA decay 0.9."""
  import torch.nn.functional as F
  g = torch.Generator().manual_seed(0)
  x = torch.randn((3, 5, 7, 8), generator=g) * 2 + 1
  p = {'gamma': torch.rand(8, generator=g) + 0.5, 'beta': torch.randn(8, generator=g),
       'ema_mean': torch.randn(8, generator=g), 'ema_var': torch.rand(8, generator=g) + 0.5}
  normed, mean, var, nm, nv = OM.batch_norm_train(x, p)
  ref = F.batch_norm(x.permute(0, 3, 1, 2), None, None, p['gamma'], p['beta'], training=True, eps=1e-3).permute(0, 2, 3, 1)
  assert torch.allclose(normed, ref, atol=1e-5)
  assert torch.allclose(mean, x.reshape(-1, 8).mean(0), atol=1e-6)
  assert torch.allclose(var, x.reshape(-1, 8).var(0, unbiased=False), atol=1e-5)
  assert torch.allclose(nm, 0.9 * p['ema_mean'] + 0.1 * mean, atol=1e-6)
  assert torch.allclose(nv, 0.9 * p['ema_var'] + 0.1 * var, atol=1e-6)


def test_random_transformation_oracle_semantics():
  """image_ops.py:9-113: offset == padding without flips is the identity (the eval path); flips and transpose
  compose in the reference's order (crop, vflip, hflip, transpose)."""
  g = torch.Generator().manual_seed(1)
  x = torch.rand((2, 6, 6, 3), generator=g)
  y = torch.rand((2, 4, 6, 6), generator=g)
  r = OM.random_transformation(x, 2, (2, 2), y=y)
  assert torch.equal(r['x'], x) and torch.equal(r['y'], y)
  r = OM.random_transformation(x, 2, (0, 4), y=y)           # shifted crop: zeros come in from the padding
  assert torch.equal(r['x'][:, 2:, :4], x[:, :4, 2:]) and float(r['x'][:, :2].abs().sum()) == 0.0
  assert float(r['x'][:, :, 4:].abs().sum()) == 0.0
  r = OM.random_transformation(x, 2, (2, 2), vflip=True, hflip=True, transpose=True, y=y)
  assert torch.equal(r['x'], torch.flip(x, [1, 2]).permute(0, 2, 1, 3))
  assert torch.equal(r['y'], torch.flip(y, [2, 3]).permute(0, 1, 3, 2))


def test_knob_oracle_semantics():
  """Scheduled sampling (full_model.py:589-625,744-785,826-843) in the oracle: switches off == plain training
  forward; box switch on == the matched noisy GT box drives the glimpse; mask switch on == the canvas is written
  from the matched GT mask."""
  import rec_attend_b200 as ra
  opt = ra.config.full_model_opt('kitti', 32, 64, 3, use_knob=True)
  B, T = 3, 3
  batch = ra.synthetic.make_batch(opt, B, seed=3)
  w = ra.synthetic.make_weights(opt)
  draws = ra.synthetic.make_knob_draws(opt, B, global_step=0, seed=1)
  assert (draws['gt_knob_box'] == 1).all(), 'at step 0 the box knob probability is 1 (knob_base = 1)'
  p = ra.synthetic.knob_probability(opt, 8000 + 1500, 'knob_segm_offset')
  assert p[0] == pytest.approx(0.5) and p[1] == pytest.approx(min(1.0, 0.5 * (1 + np.log(4.0))))
  off = dict(draws, gt_knob_box=np.zeros((B, T), np.float32), gt_knob_segm=np.zeros((B, T), np.float32))
  with torch.no_grad():
    plain = OM.full_model_forward(dict(opt, use_knob=False), w, batch, phase_train=True)
    r_off = OM.full_model_forward(opt, w, batch, phase_train=True, draws=off)
    r_on = OM.full_model_forward(opt, w, batch, phase_train=True, draws=draws)
  for k in ('y_out', 'attn_box', 'attn_ctr', 's_out'):
    assert torch.equal(plain[k], r_off[k]), k
  assert torch.allclose(plain['iou_soft_box_pairwise'], r_off['iou_soft_box_pairwise'], atol=1e-6)
  assert float(plain['loss']) == pytest.approx(float(r_off['loss']), abs=1e-5)
  # box switch on: centre / size of step 0 are those of a (noisy) GT box, not the controller's
  tl_n, br_n, _ = OM.get_gt_box(torch.from_numpy(batch['y_gt']), padding_ratio=torch.from_numpy(draws['gt_box_pad']),
                                center_shift_ratio=torch.from_numpy(draws['gt_box_ctr_shift']),
                                min_padding=opt['padding'] + 4)
  ctr_n = (tl_n + br_n) / 2.0
  d = (r_on['attn_ctr'][:, 0].unsqueeze(1) - ctr_n).abs().sum(2).min(1)[0]
  assert float(d.max()) < 1e-4
  assert not torch.equal(r_on['attn_ctr'], plain['attn_ctr'])
  # the attention box OUTPUT of step 0 is still the controller's own box (computed before the mix, :738-741)
  assert torch.equal(r_on['attn_box'][:, 0], plain['attn_box'][:, 0])
  # mask switch: with noise 0 and the switch on, the final canvas contains every matched GT mask
  seg = dict(draws, gt_knob_segm=np.ones((B, T), np.float32), gt_segm_noise=np.zeros_like(draws['gt_segm_noise']))
  with torch.no_grad():
    r_seg = OM.full_model_forward(opt, w, batch, phase_train=True, draws=seg)
  assert float(r_seg['canvas'].max()) == 1.0


def test_iou_box_coordinate_form():
  """modellib.f_iou_box (modellib.py:206-238): strict overlap test, no eps; used by the knob / box-model greedy match
  when opt['use_iou_box'] (full_model.py:750-754, box_model.py:487-491)."""
  import rec_attend_b200 as ra
  tl_a = torch.tensor([[[0.0, 0.0]]])
  br_a = torch.tensor([[[4.0, 6.0]]])  # 4 x 6 box
  tl_b = torch.tensor([[[2.0, 3.0], [4.0, 0.0], [0.0, 0.0], [10.0, 10.0]]])
  br_b = torch.tensor([[[6.0, 9.0], [8.0, 6.0], [4.0, 6.0], [12.0, 12.0]]])
  iou = OM.f_iou_box(tl_a, br_a, tl_b, br_b)
  # overlap 2x3 = 6 of 24 + 24 - 6; touching edge (y1 == y2) is NOT an overlap; identical boxes; disjoint
  assert torch.allclose(iou, torch.tensor([[6.0 / 42.0, 0.0, 1.0, 0.0]]))
  # the flag switches both model oracles onto it
  opt = ra.config.full_model_opt('cityscapes', 32, 64, 2, use_knob=True)
  assert opt['use_iou_box']
  B = 2
  batch = ra.synthetic.make_batch(opt, B, seed=3)
  w = ra.synthetic.make_weights(opt)
  draws = ra.synthetic.make_knob_draws(opt, B, global_step=0, seed=1)
  with torch.no_grad():
    r_box = OM.full_model_forward(opt, w, batch, phase_train=True, draws=draws)
    r_soft = OM.full_model_forward(dict(opt, use_iou_box=False), w, batch, phase_train=True, draws=draws)
  tl_gt, br_gt, _ = OM.get_gt_box(torch.from_numpy(batch['y_gt']), padding_ratio=opt['attn_box_padding_ratio'],
                                  center_shift_ratio=0.0, min_padding=opt['padding'] + 4)
  # step 0's box is the controller's own in both runs; the per-step IoU rows are the coordinate form in one, soft in the other
  tl0, br0 = _pre_mix_box(opt, w, batch)
  want = OM.f_iou_box(tl0.unsqueeze(1), br0.unsqueeze(1), tl_gt, br_gt)
  assert torch.allclose(r_box['iou_soft_box_pairwise'][:, 0], want, atol=1e-6)
  assert not torch.allclose(r_box['iou_soft_box_pairwise'][:, 0], r_soft['iou_soft_box_pairwise'][:, 0], atol=1e-4)
  bopt = ra.config.box_model_opt(32, 64, 2, use_iou_box=True)
  bb = ra.synthetic.make_batch(bopt, B, seed=3)
  bw = ra.synthetic.make_weights(bopt, model='box')
  with torch.no_grad():
    rb = OM.box_model_forward(bopt, bw, bb)
  tl_g, br_g, _ = OM.get_gt_box(torch.from_numpy(bb['y_gt']), padding_ratio=bopt['attn_box_padding_ratio'],
                                center_shift_ratio=0.0)
  want = OM.f_iou_box(rb['attn_top_left'][:, 0].unsqueeze(1), rb['attn_bot_right'][:, 0].unsqueeze(1), tl_g, br_g)
  assert torch.allclose(rb['iou_soft_box_pairwise'][:, 0], want, atol=1e-6)


def _pre_mix_box(opt, w, batch):
  """Step-0 box of the controller before the knob mixes a GT box in = the plain training-mode forward's step-0 box."""
  with torch.no_grad():
    plain = OM.full_model_forward(dict(opt, use_knob=False), w, batch, phase_train=True)
  return plain['attn_top_left'][:, 0], plain['attn_bot_right'][:, 0]

