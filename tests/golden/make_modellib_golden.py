"""Generates tests/golden/modellib_golden.npz by EXECUTING THE REFERENCE'S OWN modellib.py (unmodified, imported
from /root/reference) on seeded inputs, with tests/golden/tf012_shim standing in for TensorFlow 0.12 (see its README).
Run in the build container (needs /root/reference):  python tests/golden/make_modellib_golden.py
The fixtures travel with the repo; tests/test_modellib_golden.py compares oracle/model.py with them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'


def main():
  sys.path.insert(0, ROOT)
  sys.path.insert(0, REF)
  sys.path.insert(0, os.path.join(HERE, 'tf012_shim'))  # shadows tensorflow / utils
  import modellib as M  # the reference source file itself
  assert os.path.dirname(os.path.abspath(M.__file__)) == REF, M.__file__

  rng = np.random.default_rng(20261017)
  B, T, H, W, F, D = 3, 5, 24, 40, 6, 4
  f32 = np.float32
  out = {}

  def put(name, value):
    out[name] = np.asarray(value)

  # ---- masks, scores
  yy, xx = np.mgrid[0:H, 0:W]
  y_gt = np.zeros((B, T, H, W), f32)
  s_gt = np.zeros((B, T), f32)
  for b in range(B):
    for t in range(T if b == 0 else 3):
      cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(3, 9)
      y_gt[b, t] = ((yy - cy)**2 + (xx - cx)**2 <= r * r)
      s_gt[b, t] = 1.0
  y_out = np.clip(0.7 * y_gt[:, rng.permutation(T)] + 0.3 * rng.random((B, T, H, W)), 0, 1).astype(f32)
  s_out = rng.uniform(0.05, 0.95, (B, T)).astype(f32)
  put('in_y_gt', y_gt), put('in_s_gt', s_gt), put('in_y_out', y_out), put('in_s_out', s_out)

  iou = M.f_iou(y_out, y_gt, T, pairwise=True)
  put('f_iou_pairwise', iou)
  put('f_iou_aligned', M.f_iou(y_out, y_gt))
  put('f_dice_pairwise', M.f_dice(y_out, y_gt, T, pairwise=True))
  put('f_inter', M.f_inter(y_out, y_gt)), put('f_union', M.f_union(y_out, y_gt))
  put('f_iou_all', M.f_iou_all(y_out, y_gt))
  put('f_coverage_weight', M.f_coverage_weight(y_gt))
  put('f_weighted_coverage', M.f_weighted_coverage(iou, y_gt))
  match = M.f_segm_match(iou, s_gt)
  put('f_segm_match', match)
  cnt = np.maximum(match.sum(axis=(1, 2)), 1).astype(f32)
  put('f_unweighted_coverage', M.f_unweighted_coverage(iou, cnt))
  put('f_conf_loss', M.f_conf_loss(s_out, match, T))
  put('f_cum_min', M.f_cum_min(s_out, T)), put('f_cum_max', M.f_cum_max(s_out, T))
  put('f_count_acc', M.f_count_acc(s_out, s_gt))
  put('f_dic', M.f_dic(s_out, s_gt)), put('f_dic_abs', M.f_dic(s_out, s_gt, abs=True))
  put('f_bce', M.f_bce(y_out, y_gt)), put('f_ce', M.f_ce(y_out, y_gt))
  score = rng.random((B, T)).astype(f32)
  score[1, 2] = score[1, 4] = 2.0  # a tie at the maximum
  put('in_score', score)
  put('f_greedy_match', M.f_greedy_match(score, np.zeros((B, T), f32)))
  put('get_identity_match', M.get_identity_match(B, T, s_gt))

  # ---- boxes
  tl_a = rng.uniform(0, 20, (B, T, 2)).astype(f32)
  br_a = (tl_a + rng.uniform(2, 20, (B, T, 2))).astype(f32)
  tl_b = rng.uniform(0, 20, (B, T, 2)).astype(f32)
  br_b = (tl_b + rng.uniform(2, 20, (B, T, 2))).astype(f32)
  put('in_tl_a', tl_a), put('in_br_a', br_a), put('in_tl_b', tl_b), put('in_br_b', br_b)
  put('f_iou_box', M.f_iou_box(tl_a, br_a, tl_b, br_b))
  for name, kw in (('default', {}), ('padded', {'padding_ratio': 0.2, 'min_padding': 8.0}),
                   ('shifted', {'padding_ratio': 0.1, 'center_shift_ratio': 0.15, 'min_padding': 3.0})):
    tl, br, box = M.get_gt_box(y_gt, **kw)
    put('get_gt_box_%s_tl' % name, tl), put('get_gt_box_%s_br' % name, br), put('get_gt_box_%s_box' % name, box)
  ctr, sz, lg_var, lg_gamma, box, tl, br = M.get_gt_attn(y_gt, F, F, padding_ratio=0.2, min_padding=8.0)
  put('get_gt_attn_ctr', ctr), put('get_gt_attn_size', sz), put('get_gt_attn_lg_var', lg_var)
  put('get_gt_attn_lg_gamma', lg_gamma)

  # ---- attention
  ctr_norm = rng.uniform(-0.6, 0.6, (B, 2)).astype(f32)
  lg_size = rng.uniform(-1.5, -0.3, (B, 2)).astype(f32)
  put('in_ctr_norm', ctr_norm), put('in_lg_size', lg_size)
  c, s = M.get_unnormalized_attn(ctr_norm, lg_size, H, W)
  put('get_unnormalized_center', c), put('get_unnormalized_size', s)
  put('get_normalized_center', M.get_normalized_center(c, H, W)), put('get_normalized_size', M.get_normalized_size(s, H, W))
  lgv = M.get_normalized_var(s, F, F)
  put('get_normalized_var', lgv), put('get_normalized_gamma', M.get_normalized_gamma(s, F, F))
  tl, br = M.get_box_coord(c, s)
  put('get_box_coord_tl', tl), put('get_box_coord_br', br)
  f_y = M.get_gaussian_filter(c[:, 0], s[:, 0], lgv[:, 0], H, F)
  f_x = M.get_gaussian_filter(c[:, 1], s[:, 1], lgv[:, 1], W, F)
  put('get_gaussian_filter_y', f_y), put('get_gaussian_filter_x', f_x)
  x = rng.random((B, H, W, D)).astype(f32)
  put('in_x', x)
  patch = M.extract_patch(x, f_y, f_x, D)
  put('extract_patch', patch)
  p1 = rng.random((B, F, F, 1)).astype(f32)
  put('in_patch1', p1)
  put('paste_back', M.extract_patch(p1, np.transpose(f_y, [0, 2, 1]), np.transpose(f_x, [0, 2, 1]), 1))

  path = os.path.join(HERE, 'modellib_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, len(out), 'arrays', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
