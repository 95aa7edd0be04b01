"""Host logic of the optimiser <-> device-image plumbing (rec_attend_b200/params.py): every device weight image is a
permutation (with zero padding) of the reference's weight dict, described by integer source-index arrays; the numpy
model of ra_param_gather_f32 must rebuild each image from the flat trainable bucket bit for bit.  CPU only."""
import numpy as np

import rec_attend_b200 as ra
from rec_attend_b200 import ops, params as PM
from rec_attend_b200.optim import FlatParams


def _weights():
  opt = ra.config.full_model_opt('kitti', 64, 128, 2)
  return opt, ra.synthetic.make_weights(opt, seed=7)


def test_pack_umma_matches_value_packing_and_roundtrips_through_codes():
  opt, w = _weights()
  lay = PM.AllLayout(w)
  flat_all = np.concatenate([np.asarray(w[k], np.float32).reshape(-1) for k in lay.keys])
  fp = FlatParams(w)
  tmap = PM.to_train_map(lay, fp)
  flat_train = fp.flatten(w)
  for key, KC, NPc, nsp in (('ctrl_cnn_w_1', 16, 16, 1), ('attn_cnn_w_0', 8, 16, 1), ('attn_dcnn_w_1', 16, 32, 2)):
    wi = PM.wi_of(w, lay, key)
    if key == 'attn_cnn_w_0':
      wi = wi.map(lambda a: PM.pad_cin(a, 16))
    if key.startswith('attn_dcnn'):
      wi = wi.map(PM.deconv_to_conv)
    packed = PM.pack_umma(wi, KC, NPc, nsp)
    assert np.array_equal(packed.val, ops.pack_umma_weights(wi.val, KC, NPc, nsp))  # same image as the value path
    # the index array really names the source of every element
    src = np.where(wi.idx > 0, flat_all[np.maximum(wi.idx - 1, 0)], 0.0)
    assert np.array_equal(src, wi.val)
    code = PM.encode(packed.idx, packed.kind, tmap)
    assert code is not None and code.dtype == np.int32
    assert np.array_equal(PM.gather_reference(flat_train, code).reshape(packed.val.shape), packed.val)


def test_non_trainable_sources_are_left_alone_and_frozen_keys_drop_out():
  opt, w = _weights()
  lay = PM.AllLayout(w)
  fp = FlatParams(w, frozen=('ctrl_cnn_w_2',))
  tmap = PM.to_train_map(lay, fp)
  assert PM.encode(lay.index('ctrl_cnn_0_1_ema_mean'), None, tmap) is None
  assert PM.encode(lay.index('ctrl_cnn_w_2'), None, tmap) is None
  code = PM.encode(lay.index('ctrl_cnn_w_3'), None, tmap)
  off, shape = fp.layout['ctrl_cnn_w_3']
  assert np.array_equal(code >> 2, np.arange(off + 1, off + 1 + int(np.prod(shape))))


def test_transforms_commute_with_indices():
  opt, w = _weights()
  lay = PM.AllLayout(w)
  flat_all = np.concatenate([np.asarray(w[k], np.float32).reshape(-1) for k in lay.keys])
  wi = PM.wi_of(w, lay, 'attn_dcnn_w_3').map(PM.deconv_to_conv).map(PM.flip_transpose)
  assert np.array_equal(flat_all[wi.idx - 1], wi.val)
  st = PM.WI.stack([PM.wi_of(w, lay, 'ctrl_lstm_w_x' + g) for g in 'ifou'])
  assert st.val.shape[0] == 4 and np.array_equal(flat_all[st.idx - 1], st.val)
  hi = PM.tf32_hi(wi.val)
  assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.abs(wi.val - hi).max() <= np.abs(wi.val).max() * 2.0**-11
