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


def test_f16_filter_image_source_and_numpy_model():
  """fp16 hi / lo filter images (the default operand format of the narrow conv layers): the registered fp32 source
  rebuilds from the trainable bucket through the gather codes, and the numpy model of ra_umma_pack_f16 - hi = fp16(w),
  lo' = fp16((w - hi) * 2^11), rows [hi | lo'] per 8-channel plane - reproduces w to 2^-21 of its magnitude (2^-24 + the
  fp16 subnormal step / 2^11 in absolute terms for tiny weights)."""
  opt, w = _weights()
  lay = PM.AllLayout(w)
  fp = FlatParams(w)
  tmap = PM.to_train_map(lay, fp)
  flat_train = fp.flatten(w)
  for key, KC, NPc, nsp in (('ctrl_cnn_w_1', 16, 16, 1), ('ctrl_cnn_w_3', 32, 32, 1), ('attn_dcnn_w_1', 32, 32, 2)):
    wi = PM.wi_of(w, lay, key)
    if key.startswith('attn_dcnn'):
      wi = wi.map(PM.deconv_to_conv)
    src = PM.umma_f16_source(wi, KC, NPc, nsp)
    Cin, Cout = wi.val.shape[2], wi.val.shape[3]
    nch = (Cin + KC - 1) // KC
    assert src.val.shape == (nsp, nch, 9, KC // 4, NPc, 4) and src.kind is None
    code = PM.encode(src.idx, None, tmap)
    assert np.array_equal(PM.gather_reference(flat_train, code).reshape(src.val.shape), src.val)
    img = PM.pack_umma_f16_reference(src.val)
    assert img.shape == (nsp, nch, 9, KC // 8, 2 * NPc, 8) and img.dtype == np.float16
    hi, lo = img[..., :NPc, :].astype(np.float64), img[..., NPc:, :].astype(np.float64)
    # back to [.., KC/4, NPc, 4]: halves 0-3 of plane k are fp32 plane 2k, halves 4-7 plane 2k + 1
    rec = (hi + lo / 2048.0).reshape(nsp, nch, 9, KC // 8, NPc, 2, 4).swapaxes(-3, -2).reshape(src.val.shape)
    err = np.abs(rec - src.val.astype(np.float64))
    assert (err <= np.abs(src.val) * 2.0 ** -21 + 2.0 ** -35).all()
  # dynamic range: values down to the fp16 subnormals keep an absolute error of 2^-36, large ones 2^-22 relative
  x = (np.random.default_rng(3).standard_normal((1, 1, 9, 4, 16, 4)) *
       np.exp(np.random.default_rng(4).uniform(-20, 8, (1, 1, 9, 4, 16, 4)))).astype(np.float32)
  img = PM.pack_umma_f16_reference(x)
  rec = (img[..., :16, :].astype(np.float64) + img[..., 16:, :].astype(np.float64) / 2048.0)
  rec = rec.reshape(1, 1, 9, 2, 16, 2, 4).swapaxes(-3, -2).reshape(x.shape)
  assert (np.abs(rec - x) <= np.abs(x) * 2.0 ** -21 + 2.0 ** -35).all()
