"""Generates tests/golden/reference_fd_golden.npz: CENTRAL DIFFERENCES OF THE REFERENCE GRAPH'S OWN LOSS.
The reference's full_model.get_model(opt) (unmodified source over tests/golden/tf012_shim, float64, training mode) is
evaluated at w +- eps for a handful of weight entries; (L+ - L-) / 2 eps is the true derivative of the reference's
total loss (data terms + weight decay) with nothing held back.  tests/test_train_step_oracle.py compares these numbers
with torch.autograd through the oracle (stop_canvas_grad=False, the setting in which autograd computes that same true
derivative), which pins the gradient oracle - and everything checked against it: oracle/backward_manual.py and the
CUDA backward blocks - to the reference's code.  (TensorFlow's autodiff itself cannot run here; the one thing this
cannot see is WHERE the reference stops gradients: full_model.py:846-848, two lines, restated by inspection.)
Run:  python tests/golden/make_reference_fd_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
PROBES = ['ctrl_cnn_w_0', 'ctrl_cnn_w_5', 'ctrl_cnn_2_1_gamma', 'ctrl_cnn_5_0_beta', 'ctrl_lstm_w_xi', 'ctrl_lstm_w_hf',
          'ctrl_lstm_b_o', 'glimpse_mlp_w_0', 'glimpse_mlp_w_1', 'ctrl_mlp_w_0', 'ctrl_mlp_b_0', 'attn_cnn_w_1',
          'attn_cnn_3_0_gamma', 'attn_dcnn_w_2', 'attn_dcnn_6_1_beta', 'score_mlp_w_0']
CASES = [('cvppp', 'cvppp', {'use_knob': False}), ('kitti', 'kitti', {'use_knob': False})]
EPS = 1e-9


def main():
  os.environ['TF012_SHIM_DTYPE'] = 'float64'
  sys.path.insert(0, ROOT)
  sys.path.insert(0, REF)
  sys.path.insert(0, os.path.join(HERE, 'tf012_shim'))
  import h5py
  import tensorflow as tf
  import full_model as FM
  import rec_attend_b200 as ra
  assert os.path.dirname(os.path.abspath(FM.__file__)) == REF
  out = {}
  H = W = 64
  T = B = 2
  for name, arch, over in CASES:
    over = dict(over, ctrl_rnn_hid_dim=32, ctrl_mlp_dim=32, stop_canvas_grad=False)
    opt = ra.config.full_model_opt(arch, H, W, T, **over)
    batch = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_batch(opt, B, seed=21).items()}
    w0 = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_weights(opt, seed=4321).items()}

    def loss_at(w):
      h5py.REGISTRY['weights.h5'] = w
      feed = [('x', batch['x']), ('y_gt', batch['y_gt']), ('s_gt', batch['s_gt'])]
      if opt.get('add_d_out', False):
        feed += [('d_in', batch['d_in']), ('y_in', batch['y_in'])]
      feed.append(('phase_train', True))
      tf.reset(feed, seed=7)
      return float(np.asarray(FM.get_model(dict(opt, pretrain_net='weights.h5'))['loss']))

    rng = np.random.default_rng(5)
    rows = []
    for key in PROBES:
      a = w0[key]
      for _ in range(2):
        idx = tuple(int(rng.integers(0, s)) for s in a.shape)
        eps = EPS * max(1.0, abs(float(a[idx])))
        wp, wm = dict(w0), dict(w0)
        ap, am = a.copy(), a.copy()
        ap[idx] += eps
        am[idx] -= eps
        wp[key], wm[key] = ap, am
        rows.append({'key': key, 'idx': idx, 'fd': (loss_at(wp) - loss_at(wm)) / (2 * eps)})
    out[name + '/meta'] = np.array(json.dumps({'arch': arch, 'H': H, 'W': W, 'T': T, 'B': B, 'overrides': over,
                                               'batch_seed': 21, 'weight_seed': 4321, 'rows': rows}))
    print(name, len(rows), 'probes', [round(r['fd'], 4) for r in rows[:6]])
  path = os.path.join(HERE, 'reference_fd_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
