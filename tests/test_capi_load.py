"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol that include/rec_attend_b200.h declares; host-side logic (configs, weight packing, plans)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  hdr = open(os.path.join(ROOT, 'include', 'rec_attend_b200.h')).read()
  hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
  return sorted(set(re.findall(r'\b(ra_[a-z0-9_]+)\s*\(', hdr)))


def test_library_exports_every_declared_symbol():
  from rec_attend_b200 import _lib
  lib = _lib.lib()
  names = _declared()
  assert len(names) >= 18
  for n in names:
    assert hasattr(lib, n), 'librecattend_b200.so does not export ' + n
  assert sorted(_lib.EXPORTED) == names, 'ctypes table and header disagree'
  assert lib.ra_version() >= 100


def test_no_cpu_fallback_without_gpu():
  import torch
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  import rec_attend_b200 as ra
  from rec_attend_b200 import _lib
  from rec_attend_b200.full_model import FullModel
  assert _lib.lib().ra_device_count() == 0
  with pytest.raises(_lib.RecAttendError):
    FullModel(ra.config.baseline_opt(0))


def test_baseline_configs_and_input_depths():
  import rec_attend_b200 as ra
  names = [c['name'] for c in ra.config.BASELINE_CONFIGS]
  assert names[2] == 'kitti_256x512_T20_B32'
  assert ra.config.input_depths(ra.config.baseline_opt(1))[0] == 4
  assert ra.config.input_depths(ra.config.baseline_opt(2))[0] == 13
  assert ra.config.input_depths(ra.config.baseline_opt(3))[0] == 21
  opt = ra.config.baseline_opt(2)
  assert opt['disable_overwrite'] is False and opt['dynamic_var'] and not opt['fixed_gamma']  # run_kitti.sh:68-111
  assert ra.synthetic.dcnn_skip_channels(opt) == [0, 64, 64, 32, 32, 16, 13]
  assert ra.synthetic.dcnn_skip_channels(ra.config.baseline_opt(1)) == [0] * 7


def test_synthetic_batch_contract():
  import rec_attend_b200 as ra
  opt = ra.config.full_model_opt('kitti', 32, 64, 6)
  b = ra.synthetic.make_batch(opt, 3)
  assert b['x'].shape == (3, 32, 64, 3) and b['d_in'].shape == (3, 32, 64, 8) and b['y_in'].shape == (3, 32, 64, 1)
  area = b['y_gt'].sum((2, 3))
  assert (np.diff(area, axis=1) <= 0).all(), 'masks sorted by area, descending (ins_seg_dataset.py:169-172)'
  assert ((area > 0) == (b['s_gt'] > 0)).all()
  assert b['y_gt'].sum(1).max() <= 1.0, 'instances are disjoint'


def test_umma_f16_plans():
  """The operand-format switch and the 16-channel-box feasibility rule of the fp16 hi / lo split (host logic only)."""
  from rec_attend_b200 import ops
  prev = ops.umma_set_f16(1)
  try:
    assert ops.umma_set_f16(-1) == 1
    assert ops.umma_plan(16, 16, 128, 256, 2, 32)[4] & 2          # narrow layer: fp16 split
    assert ops.umma_plan(32, 32, 24, 24, 1, 32, C2=16)[4] & 2     # concatenation on a 16-channel boundary
    assert not ops.umma_plan(16, 16, 48, 48, 1, 32, C2=8)[4] & 2  # 8 + 8: a box would straddle the two inputs
    assert not ops.umma_plan(13, 16, 48, 48, 1, 32)[4] & 2        # no TMA feed at all
    assert not ops.umma_plan(64, 256, 12, 12, 1, 32)[0:5][4] & 2 or ops.umma_plan(64, 256, 12, 12, 1, 32)[1] <= 64
    ops.umma_set_f16(0)
    assert ops.umma_plan(16, 16, 128, 256, 2, 32)[4] & 2 == 0
  finally:
    ops.umma_set_f16(prev)
  with pytest.raises(Exception):
    ops.pack_umma_weights(np.zeros((3, 3, 16, 16), np.float32), 16, 16, 1, 2)  # fp16 images are made on the device


def test_umma_plan_and_weight_packing():
  from rec_attend_b200 import ops
  prev = ops.umma_set_f16(0)  # the host-packed (3xTF32) image
  try:
    _umma_plan_and_weight_packing(ops)
  finally:
    ops.umma_set_f16(prev)


def _umma_plan_and_weight_packing(ops):
  KC, NPc, nsp, nch, rs = ops.umma_plan(64, 96, 12, 12, 2, 32)
  assert KC in (8, 16, 32) and NPc * nsp == 96 and nch == 64 // KC and not rs & 1  # 6 * NPc > 256: no row stacking
  rng = np.random.default_rng(0)
  w = rng.standard_normal((3, 3, 13, 40)).astype(np.float32)
  for B in (1, 32):
    KC, NPc, nsp, nch, rs = ops.umma_plan(13, 40, 48, 48, 1, B)
    assert NPc * nsp == 48 and NPc % 16 == 0
    wp = ops.pack_umma_weights(w, KC, NPc, nsp)
    assert wp.shape == (nsp, nch, 9, KC // 4, 2 * NPc, 4)
    if rs:  # narrow split: the plan wants the row-stacked image
      assert ops.pack_umma_weights(w, KC, NPc, nsp, rs).shape == (nsp, nch, 3, KC // 4, 6 * NPc, 4)
    hi, lo = wp[:, :, :, :, :NPc], wp[:, :, :, :, NPc:]
    # hi + lo reproduces w exactly; hi has at most 11 significant mantissa bits
    rec = (hi + lo).transpose(2, 1, 3, 5, 0, 4).reshape(9, nch * KC, nsp * NPc)
    assert (rec[:, :13, :40] == w.reshape(9, 13, 40)).all() and (rec[:, 13:] == 0).all() and (rec[:, :, 40:] == 0).all()
    assert (np.ascontiguousarray(hi).view(np.uint32) & np.uint32(0x1FFF) == 0).all()
  # the row-stacked image (narrow layers, RA_UMMA_ROWSTACK): the three kx taps of a filter row along N
  KC, NPc, nsp, nch, rs = ops.umma_plan(16, 16, 128, 256, 2, 32)
  assert NPc == 16 and rs in (0, 1, 2)  # 2: the fp16 hi / lo split (default operand format of narrow layers)
  w16 = rng.standard_normal((3, 3, 16, 16)).astype(np.float32)
  st = ops.pack_umma_weights(w16, KC, NPc, nsp, 1)
  assert st.shape == (nsp, nch, 3, KC // 4, 6 * NPc, 4)
  rec = st[..., :3 * NPc, :] + st[..., 3 * NPc:, :]  # hi + lo: [ns, nch, ky, pl, kx*NPc + co, 4]
  rec = rec.reshape(nsp, nch, 3, KC // 4, 3, NPc, 4).transpose(2, 4, 1, 3, 6, 0, 5).reshape(3, 3, nch * KC, nsp * NPc)
  assert (rec[:, :, :16, :16] == w16).all()
  with pytest.raises(Exception):
    ops.umma_plan(8, 8, 7, 7, 1, 1)  # odd output width is not supported


def test_fold_bn_and_deconv_filter_transform():
  from rec_attend_b200.full_model import _deconv_to_conv, _fold_bn
  w = {'n_0_0_gamma': np.array([2.0], np.float32), 'n_0_0_beta': np.array([0.5], np.float32),
       'n_0_0_ema_mean': np.array([1.0], np.float32), 'n_0_0_ema_var': np.array([3.0], np.float32)}
  sc, sh = _fold_bn(w, 'n', 0, 1, np.array([0.25], np.float32))
  inv = 2.0 / np.sqrt(3.0 + 1e-3)
  assert abs(sc[0, 0] - inv) < 1e-6 and abs(sh[0, 0] - (0.5 - 1.0 * inv + 0.25 * inv)) < 1e-6
  wt = np.arange(3 * 3 * 2 * 5, dtype=np.float32).reshape(3, 3, 2, 5)
  wc = _deconv_to_conv(wt)
  assert wc.shape == (3, 3, 5, 2) and wc[0, 2, 4, 1] == wt[2, 0, 1, 4]
