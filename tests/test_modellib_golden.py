"""oracle/model.py against golden vectors produced by EXECUTING THE REFERENCE'S OWN modellib.py
(tests/golden/make_modellib_golden.py: /root/reference/modellib.py imported unmodified, with the numpy stand-in of
tests/golden/tf012_shim for the ~30 TensorFlow-0.12 ops it uses).  This pins the math library of the path - IoU /
DICE / coverage / matching / confidence loss / greedy match / GT boxes / box maths / Gaussian filters / glimpse and
paste-back - to the reference's code; TensorFlow's own kernel semantics (one numpy line per op in the shim) are the
only thing taken on trust.  CPU only; the fixtures travel with the repo."""
import os

import numpy as np
import pytest
import torch

import rec_attend_b200 as ra
from oracle import model as OM

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'modellib_golden.npz'))
t = lambda k: torch.from_numpy(G[k])
TOL = 2e-6


def close(a, b, tol=TOL):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  assert a.shape == b.shape, (a.shape, b.shape)
  scale = max(float(np.abs(b).max()), 1e-6)
  assert float(np.abs(a - b).max()) <= tol * scale, float(np.abs(a - b).max()) / scale


def test_overlap_scores():
  y_out, y_gt = t('in_y_out'), t('in_y_gt')
  close(OM.f_iou_pairwise(y_out, y_gt), G['f_iou_pairwise'])
  close(OM.f_inter(y_out, y_gt), G['f_inter'])
  close(OM.f_union(y_out, y_gt), G['f_union'])
  close(OM.f_inter(y_out, y_gt) / OM.f_union(y_out, y_gt), G['f_iou_aligned'])
  close(OM.f_dice_pairwise(y_out, y_gt), G['f_dice_pairwise'])
  close(OM.f_iou_all(y_out, y_gt), G['f_iou_all'])
  close(OM.f_bce(y_out, y_gt), G['f_bce'])
  close(OM.f_ce(y_out, y_gt), G['f_ce'])
  close(OM.f_iou_box(t('in_tl_a'), t('in_br_a'), t('in_tl_b'), t('in_br_b')), G['f_iou_box'])


def test_matching_coverage_and_confidence_loss():
  y_gt, s_gt, s_out = t('in_y_gt'), t('in_s_gt'), t('in_s_out')
  iou = t('f_iou_pairwise')
  match = OM.f_segm_match(iou, s_gt)
  assert np.array_equal(match.numpy(), G['f_segm_match'])  # masking, rounding, +1e-5, Hungarian, masking again
  assert G['f_segm_match'].sum() == s_gt.sum()
  close(OM.f_coverage_weight(y_gt), G['f_coverage_weight'])
  close(OM.f_weighted_coverage(iou, y_gt), G['f_weighted_coverage'])
  cnt = torch.clamp(match.sum(dim=(1, 2)), min=1.0)
  close(OM.f_unweighted_coverage(iou, cnt), G['f_unweighted_coverage'])
  close(OM.f_cum_min(s_out), G['f_cum_min'])
  close(OM.f_cum_max(s_out), G['f_cum_max'])
  close(OM.f_conf_loss(s_out, match), G['f_conf_loss'])
  close(OM.f_count_acc(s_out, s_gt), G['f_count_acc'])
  close(OM.f_dic(s_out, s_gt, False), G['f_dic'])
  close(OM.f_dic(s_out, s_gt, True), G['f_dic_abs'])
  g = OM.f_greedy_match(t('in_score'), torch.zeros_like(t('in_score')))
  assert np.array_equal(g.numpy(), G['f_greedy_match']) and G['f_greedy_match'][1, 2] == 0.5  # ties share 1/k


@pytest.mark.parametrize('name,kw', [('default', {}), ('padded', {'padding_ratio': 0.2, 'min_padding': 8.0}),
                                     ('shifted', {'padding_ratio': 0.1, 'center_shift_ratio': 0.15, 'min_padding': 3.0})])
def test_ground_truth_boxes(name, kw):
  tl, br, box = OM.get_gt_box(t('in_y_gt'), **kw)
  close(tl, G['get_gt_box_%s_tl' % name])
  close(br, G['get_gt_box_%s_br' % name])
  assert np.array_equal(box.numpy(), G['get_gt_box_%s_box' % name])
  empty = G['in_s_gt'] == 0  # empty masks: top-left corner box, EMPTY filled box (computed before the fix-up)
  assert (G['get_gt_box_%s_box' % name][empty].sum() == 0) and (G['get_gt_box_%s_tl' % name][empty] == 0).all()
  if name == 'padded':  # get_gt_attn = the same box as centre / size
    close((tl + br) / 2.0, G['get_gt_attn_ctr'])
    close(br - tl, G['get_gt_attn_size'])


def test_box_maths_filters_glimpse_and_paste_back():
  B = G['in_ctr_norm'].shape[0]
  H, W = G['in_y_gt'].shape[2:]
  F = G['get_gaussian_filter_y'].shape[2]
  opt = dict(ra.config.full_model_opt('cvppp', H, W, 2), filter_height=F, filter_width=F)
  assert not opt['fixed_var'] and not opt['dynamic_var'] and not opt['squash_ctrl_params']
  ctrl_out = torch.zeros(B, 9)
  ctrl_out[:, 0:2], ctrl_out[:, 2:4] = t('in_ctr_norm'), t('in_lg_size')
  p = OM.box_params(opt, ctrl_out)
  close(p['ctr'], G['get_unnormalized_center'])
  close(p['size'], G['get_unnormalized_size'])
  close(p['lg_var'], G['get_normalized_var'], tol=1e-5)
  close(OM.top_left_pred(p['ctr'], p['size']), G['get_box_coord_tl'])
  close(OM.bot_right_pred(p['ctr'], p['size']), G['get_box_coord_br'])
  f_y = OM.get_gaussian_filter(p['ctr'][:, 0], p['size'][:, 0], p['lg_var'][:, 0], H, F)
  f_x = OM.get_gaussian_filter(p['ctr'][:, 1], p['size'][:, 1], p['lg_var'][:, 1], W, F)
  close(f_y, G['get_gaussian_filter_y'], tol=2e-5)
  close(f_x, G['get_gaussian_filter_x'], tol=2e-5)
  close(OM.extract_patch(t('in_x'), t('get_gaussian_filter_y'), t('get_gaussian_filter_x'), G['in_x'].shape[3]),
        G['extract_patch'], tol=1e-5)
  close(OM.extract_patch(t('in_patch1'), t('get_gaussian_filter_y').transpose(1, 2),
                         t('get_gaussian_filter_x').transpose(1, 2), 1), G['paste_back'], tol=1e-5)


GI = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'image_ops_golden.npz'))


@pytest.mark.parametrize('name', sorted({k.split('/')[0] for k in GI.files}))
def test_random_transformation_equals_the_reference(name):
  """oracle.model.random_transformation against the reference's own image_ops.random_transformation executed over the
  shim with chosen draws (tests/golden/make_image_ops_golden.py): crop offsets, flips, transpose, the orientation
  mode (no flips) and the eval-mode identity (centre crop whatever the draws)."""
  H, W, pad, oy, ox, hf, vf, tr, with_d, train = [int(v) for v in GI[name + '/params']]
  g = lambda k: torch.from_numpy(GI['%s/%s' % (name, k)])
  d = g('in_d') if with_d else None
  c = g('in_c') if with_d else None
  if train:
    r = OM.random_transformation(g('in_x'), pad, (oy, ox), vflip=bool(vf), hflip=bool(hf), transpose=bool(tr),
                                 y=g('in_y'), d=d, c=c)
  else:  # phase_train = False: the centre slices, no flips (image_ops.py:70-80,106-112)
    r = OM.random_transformation(g('in_x'), pad, (pad, pad), y=g('in_y'), d=d, c=c)
  for k in ('x', 'y') + (('d', 'c') if with_d else ()):
    ref = GI['%s/out_%s' % (name, k)]
    assert tuple(r[k].shape) == ref.shape, (k, tuple(r[k].shape), ref.shape)
    assert np.array_equal(r[k].numpy(), ref), k
