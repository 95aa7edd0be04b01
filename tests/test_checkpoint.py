"""Weight-file / checkpoint interchange: key schema of full_model_read.py / box_model_read.py,
the box_model -> full_model pretrained hand-off of full_model.py:271-518 and box_model.py:182-330, the freeze flags,
the reference's initialisers and the run-folder Saver protocol (utils/saver.py).  Host logic only — runs on CPU."""
import os

import numpy as np
import pytest

import rec_attend_b200 as ra
from rec_attend_b200 import checkpoint as ck
from rec_attend_b200 import optim


def _opts(T=3):
  full = ra.config.full_model_opt('kitti', 64, 128, T)
  box = ra.config.box_model_opt(64, 128, T)
  return full, box


def test_weight_file_keys_follow_the_reference_readers():
  full, box = _opts()
  T = 3
  kb = ck.weight_file_keys(box, 'box')
  kf = ck.weight_file_keys(full, 'full')
  # box_model_read.py:31-52: ctrl_cnn (w, b, then beta/gamma per step), ctrl_mlp, glimpse_mlp, score_mlp, the 12 LSTM tensors
  assert kb[:4] == ['ctrl_cnn_w_0', 'ctrl_cnn_b_0', 'ctrl_cnn_0_0_beta', 'ctrl_cnn_0_0_gamma']
  assert len(kb) == 8 * (2 + 2 * T) + 2 * 1 + 2 * 2 + 2 + 12
  assert kb[-12:] == ['ctrl_lstm_' + w for w in ck.LSTM_KEYS]
  # full_model_read.py:58-70 appends the attention CNN / DCNN with their per-step BN
  assert kf[:len(kb)] == kb
  assert len(kf) == len(kb) + (6 + 7) * (2 + 2 * T)
  assert 'attn_dcnn_6_2_gamma' in kf and not any('ema' in k for k in kf)  # EMA shadows are not exported
  # every exported key exists in the model's own schema, with nothing trainable left out
  w = ra.synthetic.make_weights(full)
  assert set(kf) == set(optim.trainable_keys(w))
  wb = ra.synthetic.make_weights(box, model='box')
  assert set(kb) == set(optim.trainable_keys(wb))
  with pytest.raises(ck.CheckpointError):
    ck.weight_file_keys(full, 'fg')


def test_save_load_roundtrip_npz_and_h5_gate(tmp_path):
  full, _ = _opts()
  w = ra.synthetic.make_weights(full)
  fn = str(tmp_path / 'weights.npz')
  ck.save_weights(fn, w, ck.weight_file_keys(full))
  r = ck.load_weights(fn)
  assert sorted(r) == sorted(ck.weight_file_keys(full))
  for k in r:
    assert r[k].dtype == np.float32 and np.array_equal(r[k], w[k]), k
  assert r['attn_dcnn_w_1'].shape == (3, 3, 64, 128)  # [kh, kw, Cout, Cin + skip] (nnlib.py:320-325)
  with pytest.raises(ck.CheckpointError):
    ck.save_weights(fn, w, ['no_such_key'])
  with pytest.raises(ck.CheckpointError):
    ck.load_weights(str(tmp_path / 'absent.npz'))
  if ck._h5py() is None:  # this image: the .h5 container is refused loudly, not silently renamed
    with pytest.raises(ck.CheckpointError):
      ck.save_weights(str(tmp_path / 'weights.h5'), w)
  else:
    h5 = ck.save_weights(str(tmp_path / 'weights.h5'), w, ck.weight_file_keys(full))
    assert np.array_equal(ck.load_weights(h5)['ctrl_lstm_w_xi'], w['ctrl_lstm_w_xi'])


def test_reference_init_distributions():
  full, _ = _opts()
  w = ck.reference_init(full, seed=3)
  assert set(w) == set(ra.synthetic.make_weights(full))
  a = w['ctrl_lstm_w_hi']
  assert abs(float(a.std()) - 0.0088) < 5e-4 and float(np.abs(a).max()) <= 0.02  # sigma 0.01 truncated at 2 sigma
  assert float(np.abs(w['ctrl_cnn_b_0']).max()) <= 0.02 and float(np.abs(w['ctrl_cnn_b_0']).max()) > 0  # biases too
  assert (w['ctrl_lstm_b_f'] == 1).all() and (w['ctrl_lstm_b_i'] == 0).all() and (w['ctrl_lstm_b_o'] == 0).all()
  assert (w['attn_cnn_2_1_gamma'] == 1).all() and (w['attn_cnn_2_1_beta'] == 0).all()
  assert (w['ctrl_cnn_0_0_ema_mean'] == 0).all() and (w['ctrl_cnn_0_0_ema_var'] == 0).all()


def test_box_to_full_handoff_and_freeze_flags(tmp_path):
  """run_kitti.sh:62-66,110: box_model_read -> weights file -> full_model --pretrain_ctrl_net."""
  full, box = _opts()
  wb = ra.synthetic.make_weights(box, seed=11, model='box')
  fn = ck.save_weights(str(tmp_path / 'box_weights.npz'), wb, ck.weight_file_keys(box, 'box'))
  init = ck.reference_init(full, seed=5)
  new, frozen = ck.apply_pretrained(dict(full, pretrain_ctrl_net=fn), init)
  assert frozen == []
  for k in ck.weight_file_keys(box, 'box'):
    if k.startswith('score_mlp'):
      assert np.array_equal(new[k], init[k]), k  # only pretrain_net feeds the score MLP (full_model.py:464)
    else:
      assert np.array_equal(new[k], wb[k]), k
  for k in init:
    if k.startswith(('attn_cnn', 'attn_dcnn')) or k.endswith(('_ema_mean', '_ema_var')):
      assert np.array_equal(new[k], init[k]), k  # untouched: attention nets and every EMA shadow
  assert new['ctrl_cnn_w_0'] is not wb['ctrl_cnn_w_0'] and init['ctrl_cnn_w_0'].std() < 0.011  # inputs not modified

  # freeze flags: w / b only (BN stays trainable); freeze_ctrl_rnn also freezes the glimpse MLP (:363)
  o = dict(full, freeze_ctrl_cnn=True, freeze_ctrl_rnn=True, freeze_attn_net=True)
  _, frozen = ck.apply_pretrained(o, init, pretrain_ctrl_net=wb)
  assert 'ctrl_cnn_w_7' in frozen and 'ctrl_lstm_b_f' in frozen and 'glimpse_mlp_w_1' in frozen
  assert 'attn_dcnn_b_6' in frozen and 'attn_cnn_w_0' in frozen
  assert not any(k.endswith(('_beta', '_gamma')) for k in frozen) and 'ctrl_mlp_w_0' not in frozen
  flat = optim.FlatParams(new, frozen)
  assert set(flat.keys) == set(optim.trainable_keys(new)) - set(frozen)
  assert flat.numel == sum(new[k].size for k in flat.keys)

  # pretrain_net wins over the two partial files and also feeds the attention nets and the score MLP
  wf = ra.synthetic.make_weights(full, seed=12)
  new2, _ = ck.apply_pretrained(full, init, pretrain_net=wf, pretrain_ctrl_net=wb)
  for k in ck.weight_file_keys(full):
    assert np.array_equal(new2[k], wf[k]), k
  # a file that lacks a needed key / has another shape fails like the reference's KeyError, by name
  bad = {k: v for k, v in wb.items() if k != 'ctrl_lstm_w_xu'}
  with pytest.raises(ck.CheckpointError, match='ctrl_lstm_w_xu'):
    ck.apply_pretrained(full, init, pretrain_ctrl_net=bad)
  short = ra.synthetic.make_weights(ra.config.box_model_opt(64, 128, 2), model='box')  # trained with T = 2
  with pytest.raises(ck.CheckpointError, match='ctrl_cnn_0_2_beta'):
    ck.apply_pretrained(full, init, pretrain_ctrl_net=short)
  wrong = dict(wb, ctrl_mlp_w_0=np.zeros((256, 7), np.float32))
  with pytest.raises(ck.CheckpointError, match='shape'):
    ck.apply_pretrained(full, init, pretrain_ctrl_net=wrong)


def test_box_model_pretrained_cnn_prefixes():
  """box_model.py:182-219: first layers of the controller CNN from a file with `attn_cnn_`, `cnn_` or `ctrl_cnn_` keys."""
  _, box = _opts()
  T = box['timespan']
  init = ck.reference_init(box, seed=1, model='box')
  src = ra.synthetic.make_weights(box, seed=2, model='box')
  for prefix in ('attn_', '', 'ctrl_'):
    f = {}
    for ii in range(3):  # a 3-layer pretrained CNN
      f['{}cnn_w_{}'.format(prefix, ii)] = src['ctrl_cnn_w_%d' % ii]
      f['{}cnn_b_{}'.format(prefix, ii)] = src['ctrl_cnn_b_%d' % ii]
      for tt in range(T):
        for w in ('beta', 'gamma'):
          f['{}cnn_{}_{}_{}'.format(prefix, ii, tt, w)] = src['ctrl_cnn_%d_%d_%s' % (ii, tt, w)]
    new, frozen = ck.apply_pretrained(box, init, pretrain_cnn=f, model='box')
    assert frozen == sorted(['ctrl_cnn_w_0', 'ctrl_cnn_b_0', 'ctrl_cnn_w_1', 'ctrl_cnn_b_1', 'ctrl_cnn_w_2',
                             'ctrl_cnn_b_2'])  # freeze_pretrain_cnn defaults to True (box_model.py:47-50)
    assert np.array_equal(new['ctrl_cnn_w_2'], src['ctrl_cnn_w_2'])
    assert np.array_equal(new['ctrl_cnn_1_2_gamma'], src['ctrl_cnn_1_2_gamma'])
    assert np.array_equal(new['ctrl_cnn_w_3'], init['ctrl_cnn_w_3'])  # layers beyond the file keep their init
    assert np.array_equal(new['ctrl_lstm_w_xi'], init['ctrl_lstm_w_xi'])  # pretrain_cnn feeds the CNN only
  _, frozen = ck.apply_pretrained(dict(box, freeze_pretrain_cnn=False), init, pretrain_cnn=f, model='box')
  assert frozen == []
  new, frozen = ck.apply_pretrained(box, init, pretrain_net=src, model='box')  # a whole box model: everything but the score MLP
  assert np.array_equal(new['ctrl_lstm_w_ho'], src['ctrl_lstm_w_ho']) and len(frozen) == 16
  assert np.array_equal(new['score_mlp_w_0'], init['score_mlp_w_0'])


def test_saver_folder_protocol(tmp_path):
  full, _ = _opts()
  folder = str(tmp_path / 'results' / 'full_model-001')
  sv = ck.Saver(folder, model_opt=dict(full, base_learn_rate=np.float32(1e-3)))
  assert os.path.exists(os.path.join(folder, 'model_opt.yaml'))
  with pytest.raises(ck.CheckpointError, match='No checkpoint'):
    sv.get_latest_ckpt()
  w = ra.synthetic.make_weights(full)
  n = optim.FlatParams(w).numel
  for step in (1000, 2000, 3000):
    m = np.full(n, step, np.float32)
    sv.save(ck.pack_state(w, m, m * 2, step), step)
  files = sorted(f for f in os.listdir(folder) if f.startswith('model.ckpt'))
  assert files == ['model.ckpt-2000.npz', 'model.ckpt-3000.npz']  # max_to_keep = 2 (utils/saver.py:9)
  info = ck.Saver(folder).get_ckpt_info()
  assert info['step'] == 3000 and info['model_id'] == 'full_model-001'
  assert info['model_opt']['ctrl_cnn_depth'] == full['ctrl_cnn_depth'] and info['model_opt']['timespan'] == 3
  w2, m2, v2, step = ck.unpack_state(ck.Saver(folder).restore())
  assert step == 3000 and float(m2[0]) == 3000.0 and float(v2[-1]) == 6000.0
  assert set(w2) == set(w) and np.array_equal(w2['ctrl_cnn_0_0_ema_var'], w['ctrl_cnn_0_0_ema_var'])
  older = ck.Saver(folder).restore(os.path.join(folder, 'model.ckpt-2000.npz'))
  assert int(older['global_step']) == 2000
