"""The hand-assembled backward pass (oracle/backward_manual.py: the per-block formulas of the CUDA building blocks
chained over the decode loop - the executable spec of the GPU assembly) against autograd through the same forward
(oracle/grads.py).  Run in float64 so that the comparison tests the CHAINING (skip routing, channel order, filter
gradient accumulation, per-step BN copies, step independence), not fp32 conditioning.  CPU only."""
import numpy as np
import pytest
import torch

import rec_attend_b200 as ra
from conftest import oracle_fp64
from oracle import backward_manual as BM
from oracle import grads as OG


@pytest.mark.parametrize('arch,H,W,knob', [('cvppp', 64, 64, False), ('kitti', 64, 64, False), ('kitti', 64, 64, True),
                                          ('cityscapes', 64, 64, True)])  # Cityscapes: use_iou_box, fixed gamma
def test_manual_backward_equals_autograd(arch, H, W, knob):
  T, B = 2, 2
  opt = ra.config.full_model_opt(arch, H, W, T, use_knob=knob)
  batch = ra.synthetic.make_batch(opt, B, seed=21)
  weights = ra.synthetic.make_weights(opt, seed=4321)
  d64 = None
  if knob:  # scheduled sampling with both branches of both switches present
    draws = ra.synthetic.make_knob_draws(opt, B, global_step=9000, seed=3)
    draws['gt_knob_box'][:, 0], draws['gt_knob_box'][:, 1] = [1, 0], [0, 1]
    draws['gt_knob_segm'][:, 0] = [0, 1]
    d64 = {k: np.asarray(v, np.float64) for k, v in draws.items()}
  O64 = oracle_fp64()
  w64 = {k: np.asarray(v, np.float64) for k, v in weights.items()}
  b64 = {k: np.asarray(v, np.float64) for k, v in batch.items()}
  torch.set_default_dtype(torch.float64)
  try:
    ref, _ = OG.full_model_grads(opt, w64, b64, draws=d64, include_weight_decay=False, model_module=O64,
                                 dtype=torch.float64)
    got, out = BM.full_model_backward(opt, w64, b64, draws=d64, dtype=np.float64, model_module=O64)
  finally:
    torch.set_default_dtype(torch.float32)
  assert set(got) == set(ref), sorted(set(ref) ^ set(got))[:8]
  worst = {}
  for k in ref:
    assert got[k].shape == ref[k].shape, (k, got[k].shape, ref[k].shape)
    scale = max(float(np.abs(ref[k]).max()), 1e-9)
    worst[k] = float(np.abs(got[k] - ref[k]).max()) / scale
  bad = {k: v for k, v in worst.items() if v > 1e-6 and not ('_cnn_b_' in k or k.startswith('attn_dcnn_b_'))}
  assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:8]
  # conv biases sit in front of a batch-statistics BN: both gradients are pure round-off
  for k in ref:
    if '_cnn_b_' in k or k.startswith('attn_dcnn_b_'):
      assert float(np.abs(ref[k]).max()) < 1e-9 and float(np.abs(got[k]).max()) < 1e-9, k


def test_block_functions_match_autograd_in_isolation():
  """Each block function is the numpy twin of one C-ABI entry point; a quick stand-alone check of two of them (the
  others are covered by the end-to-end comparison above and, on the GPU, by tests/test_gpu_backward.py)."""
  from oracle import model as OM
  rng = np.random.default_rng(0)
  B, T, H, W = 2, 4, 10, 12
  a = torch.rand(B, T, H, W, dtype=torch.float64, requires_grad=True)
  g = (torch.rand(B, T, H, W, dtype=torch.float64) > 0.6).double()
  match = np.zeros((B, T, T))
  for b in range(B):
    for n, m in enumerate(rng.permutation(T)[:3]):
      match[b, n, m] = 1.0
  inter = torch.einsum('bnhw,bmhw->bnm', a, g)
  U = a.sum((2, 3)).unsqueeze(2) + g.sum((2, 3)).unsqueeze(1) - inter + H * W * 1e-5
  mt = torch.from_numpy(match)
  L = -(((inter / U * mt).sum((1, 2)) / mt.sum((1, 2)).clamp(min=1.0)).sum() / B)
  ga, = torch.autograd.grad(L, [a])
  assert np.allclose(BM.iou_loss_bwd(a.detach().numpy(), g.numpy(), match), ga.numpy(), atol=1e-12)
  s = torch.rand(B, T, dtype=torch.float64, requires_grad=True)
  Lc = OM.f_conf_loss(s, mt.float().double())
  gs, = torch.autograd.grad(Lc, [s])
  assert np.allclose(BM.conf_loss_bwd(s.detach().numpy(), match, 1.0), gs.numpy(), atol=1e-12)


def test_iou_box_coordinate_gradient():
  """modellib.f_iou_box (the use_iou_box box loss of the scheduled-sampling mode) differentiated by hand vs autograd,
  on boxes that overlap their targets partially, fully and not at all."""
  from oracle import model as OM
  rng = np.random.default_rng(4)
  B, M = 5, 6
  ctr = torch.tensor(rng.uniform(20, 40, (B, 2)), requires_grad=True)
  size = torch.tensor(rng.uniform(10, 30, (B, 2)), requires_grad=True)
  tl_gt = rng.uniform(0, 40, (B, M, 2))
  br_gt = tl_gt + rng.uniform(5, 40, (B, M, 2))
  tl_gt[0, 0], br_gt[0, 0] = [0.0, 0.0], [100.0, 100.0]   # contains the box
  tl_gt[1, 1], br_gt[1, 1] = [200.0, 200.0], [210.0, 210.0]  # disjoint
  wgt = rng.standard_normal((B, M))
  iou = OM.f_iou_box((ctr - size / 2).unsqueeze(1), (ctr + size / 2).unsqueeze(1), torch.from_numpy(tl_gt),
                     torch.from_numpy(br_gt))
  assert float((iou.detach() > 0).double().mean()) > 0.5 and float(iou.detach()[1, 1]) == 0.0
  gc, gs = torch.autograd.grad((iou * torch.from_numpy(wgt)).sum(), [ctr, size])
  dctr, dsize = BM.iou_box_coord_bwd(ctr.detach().numpy(), size.detach().numpy(), tl_gt, br_gt, wgt)
  assert np.allclose(dctr, gc.numpy(), atol=1e-12) and np.allclose(dsize, gs.numpy(), atol=1e-12)
  assert float(np.abs(gc.numpy()).max()) > 1e-4
