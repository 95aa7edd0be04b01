"""Generates tests/golden/fg_model_golden.npz by EXECUTING THE REFERENCE'S OWN fg_model.get_model(opt)
(/root/reference/fg_model.py + nnlib.py + modellib.py + image_ops.py, unmodified) over the numpy stand-in of
tests/golden/tf012_shim, float64, training mode.  fg_model.py imports `image_ops_old`, which the reference does not
ship; the shim aliases it to the reference's image_ops.py (same function and signature) - the only repair needed.
The graph initialises its own weights (no pretrained path), so they are read back from the model dict and stored
(float32, lossless) together with the outputs; channel counts are scaled down (structure, pools and skip masks of the
shipped architectures kept) so that the fixture stays small.   Run:  python tests/golden/make_fg_model_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
CASES = [('default_skip', 'default', 64, 64, 2, {'add_skip_conn': True}, 4), ('default_iou', 'default', 32, 64, 2, {}, 8),
         ('kitti', 'kitti', 64, 64, 2, {}, 16), ('cityscapes', 'cityscapes', 64, 64, 1, {}, 32)]
KEEP = ['y_out', 'd_out', 'iou_soft', 'iou_hard', 'foreground_loss', 'orientation_ce', 'orientation_acc', 'loss']


def scaled(opt, f):
  o = dict(opt)
  o['cnn_depth'] = [max(2, d // f) for d in opt['cnn_depth']]
  o['dcnn_depth'] = [max(2, d // f) for d in opt['dcnn_depth'][:-1]] + [opt['dcnn_depth'][-1]]
  return o


def main():
  os.environ['TF012_SHIM_DTYPE'] = 'float64'
  sys.path.insert(0, ROOT)
  sys.path.insert(0, REF)
  sys.path.insert(0, os.path.join(HERE, 'tf012_shim'))
  import tensorflow as tf
  import fg_model as FG  # the reference source file itself
  import rec_attend_b200 as ra
  assert os.path.dirname(os.path.abspath(FG.__file__)) == REF, FG.__file__
  out = {}
  for name, arch, H, W, B, over, f in CASES:
    opt = scaled(ra.config.fg_model_opt(arch, H, W, **over), f)
    batch = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_fg_batch(opt, B, seed=5).items()}
    feed = [(None, batch['x']), (None, batch['y_gt']), (None, True)]  # placeholder order: x, y_gt, phase_train[, d_gt]
    if opt['add_orientation']:
      feed.append((None, batch['d_gt']))
    tf.reset(feed, seed=11)
    model = FG.get_model(opt)
    assert not tf.FEED
    out[name + '/meta'] = np.array(json.dumps({'arch': arch, 'H': H, 'W': W, 'B': B, 'overrides': over,
                                               'cnn_depth': opt['cnn_depth'], 'dcnn_depth': opt['dcnn_depth'],
                                               'batch_seed': 5}))
    nw = 0
    for k, v in model.items():
      if k.startswith(('cnn_', 'dcnn_')) and not k.endswith(('ema_mean', 'ema_var')):
        a = np.asarray(v, np.float64)
        assert np.array_equal(a, a.astype(np.float32).astype(np.float64)), k
        out['%s/w/%s' % (name, k)] = a.astype(np.float32)
        nw += a.size
    for k in KEEP:
      if k in model:
        out['%s/%s' % (name, k)] = np.asarray(model[k], np.float32 if k in ('y_out', 'd_out') else np.float64)
    print(name, 'loss', float(np.asarray(model['loss'])), 'weights', nw)
  path = os.path.join(HERE, 'fg_model_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, len(out), 'arrays', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
