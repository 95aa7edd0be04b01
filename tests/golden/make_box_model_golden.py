"""Generates tests/golden/box_model_golden.npz by EXECUTING THE REFERENCE'S OWN box_model.get_model(opt)
(/root/reference/box_model.py + nnlib.py + modellib.py + image_ops.py, unmodified) over the numpy stand-in of
tests/golden/tf012_shim, in float64 and training mode (batch-statistics BN; the per-step canvas noise of
box_model.py:501-502 is logged and replayed into the oracle).  Pins oracle.model.box_model_forward - BASELINE
configs[4]'s graph - to the reference's code.   Run:  python tests/golden/make_box_model_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
CASES = [('kitti_box', 64, 64, 3, 3, {}), ('kitti_box_iou_box', 64, 128, 2, 3, {'use_iou_box': True})]
SMALL = {'ctrl_rnn_hid_dim': 32, 'ctrl_mlp_dim': 32}
KEEP = ['attn_box', 's_out', 'attn_ctr', 'attn_size', 'attn_top_left', 'attn_bot_right', 'attn_top_left_gt',
        'attn_bot_right_gt', 'match_box', 'loss', 'box_loss', 'conf_loss', 'ctrl_rnn_glimpse_map']


def main():
  os.environ['TF012_SHIM_DTYPE'] = 'float64'
  sys.path.insert(0, ROOT)
  sys.path.insert(0, REF)
  sys.path.insert(0, os.path.join(HERE, 'tf012_shim'))
  import h5py
  import tensorflow as tf
  import box_model as BM  # the reference source file itself
  import rec_attend_b200 as ra
  assert os.path.dirname(os.path.abspath(BM.__file__)) == REF, BM.__file__
  out = {}
  for name, H, W, T, B, over in CASES:
    over = dict(SMALL, **over)
    opt = ra.config.box_model_opt(H, W, T, **over)
    batch = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_batch(opt, B, seed=21).items()}
    w = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_weights(opt, seed=4321, model='box').items()}
    h5py.REGISTRY['weights.h5'] = w
    feed = [('x', batch['x']), ('y_gt', batch['y_gt']), ('s_gt', batch['s_gt']), ('d_in', batch['d_in']),
            ('y_in', batch['y_in']), ('phase_train', True)]
    tf.reset(feed, seed=3)
    model = BM.get_model(dict(opt, pretrain_net='weights.h5'))
    assert not tf.FEED
    out[name + '/meta'] = np.array(json.dumps({'H': H, 'W': W, 'T': T, 'B': B, 'overrides': over, 'batch_seed': 21,
                                               'weight_seed': 4321}))
    out[name + '/weights_checksum'] = np.float64(sum(float(np.abs(v).sum()) for v in w.values()))
    for k in KEEP:
      out['%s/%s' % (name, k)] = np.asarray(model[k], np.float32 if k == 'attn_box' else np.float64)
    noise = [r['value'] for r in tf.RANDOM_LOG if r['shape'] == (B, H, W, 1)]
    assert len(noise) == T
    out[name + '/draw_canvas_noise'] = np.asarray(np.stack([n[..., 0] for n in noise], 1), np.float32)
    print(name, 'loss', float(np.asarray(model['loss'])))
  path = os.path.join(HERE, 'box_model_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, len(out), 'arrays', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
