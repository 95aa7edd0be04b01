"""Generates tests/golden/full_model_golden.npz by EXECUTING THE REFERENCE'S OWN full_model.get_model(opt) -
/root/reference/full_model.py, nnlib.py, modellib.py and image_ops.py imported unmodified - eagerly over the numpy
stand-in of tests/golden/tf012_shim (placeholders return the queued batch, pretrained weights come from an in-memory
h5py stand-in, random draws are logged so that the oracle can be fed the same numbers, the optimiser is a stub).
The outputs pin oracle.model.full_model_forward - the whole T-step decode and the matching loss block in training
mode (batch-statistics BN), with and without scheduled sampling - to the reference's code.  TensorFlow's kernel
semantics (the shim's one-liners) are the only thing taken on trust.
Run in the build container:  python tests/golden/make_full_model_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'

# name, arch, H, W, T, B, overrides, global_step   (small hidden sizes keep the fixtures small)
CASES = [
    ('cvppp', 'cvppp', 64, 64, 3, 2, {'use_knob': False}, 0),
    ('kitti', 'kitti', 64, 64, 2, 2, {'use_knob': False}, 0),
    ('cityscapes', 'cityscapes', 64, 64, 2, 2, {'use_knob': False}, 0),
    ('kitti_knob', 'kitti', 64, 64, 3, 3, {'use_knob': True}, 9500),
    ('cityscapes_knob_iou_box', 'cityscapes', 64, 64, 2, 3, {'use_knob': True}, 9500),
    # phase_train = False: EMA statistics (TensorFlow's shadows of a fresh graph are ZERO, so every BN layer multiplies
    # by gamma / sqrt(1e-3) - numerically wild but exactly defined in float64), centre crop, knob terms switched off
    ('kitti_eval', 'kitti', 64, 64, 2, 2, {'use_knob': True, 'phase_train': False}, 9500),
]
SMALL = {'ctrl_rnn_hid_dim': 32, 'ctrl_mlp_dim': 32}
BIG = ('y_out', 'attn_box', 'y_out_patch', 'ctrl_rnn_glimpse_map')  # stored as float32 (compared at 1e-6)
KEEP = ['y_out', 's_out', 'attn_box', 'y_out_patch', 'match', 'match_box', 'loss', 'box_loss', 'segm_loss',
        'conf_loss', 'iou_soft', 'iou_hard', 'wt_cov_soft', 'unwt_cov_soft', 'wt_cov_hard', 'unwt_cov_hard', 'dice',
        'count_acc', 'dic', 'dic_abs', 'attn_ctr', 'attn_size', 'attn_top_left', 'attn_bot_right',
        'ctrl_rnn_glimpse_map', 'learn_rate']


def main():
  # double precision: the training-mode decode loop amplifies fp32 round-off ~10x per step, which would blur a
  # float32 comparison after two steps; in float64 the oracle must agree with the reference graph to ~1e-9
  os.environ['TF012_SHIM_DTYPE'] = 'float64'
  sys.path.insert(0, ROOT)
  sys.path.insert(0, REF)
  sys.path.insert(0, os.path.join(HERE, 'tf012_shim'))
  import h5py
  import tensorflow as tf
  import full_model as FM  # the reference source file itself
  import rec_attend_b200 as ra
  assert os.path.dirname(os.path.abspath(FM.__file__)) == REF, FM.__file__
  out = {}
  for name, arch, H, W, T, B, over, step in CASES:
    over = dict(SMALL, **over)
    phase = bool(over.pop('phase_train', True))
    opt = ra.config.full_model_opt(arch, H, W, T, **over)
    batch = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_batch(opt, B, seed=21).items()}
    w = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_weights(opt, seed=4321).items()}
    h5py.REGISTRY['weights.h5'] = w
    ropt = dict(opt, pretrain_net='weights.h5')
    feed = [('x', batch['x']), ('y_gt', batch['y_gt']), ('s_gt', batch['s_gt'])]
    if opt.get('add_d_out', False):
      feed += [('d_in', batch['d_in']), ('y_in', batch['y_in'])]
    feed.append(('phase_train', phase))
    tf.reset(feed, seed=7)
    tf.VARIABLE_OVERRIDES['global_step'] = float(step)
    model = FM.get_model(ropt)
    assert not tf.FEED, 'unused placeholders'
    out[name + '/meta'] = np.array(json.dumps({'arch': arch, 'H': H, 'W': W, 'T': T, 'B': B, 'overrides': over,
                                               'global_step': step, 'batch_seed': 21, 'weight_seed': 4321,
                                               'phase_train': phase}))
    out[name + '/weights_checksum'] = np.float64(sum(float(np.abs(v).sum()) for v in w.values()))
    for k in KEEP:  # (the glimpse x_patch is left out to keep the fixture small: y_out_patch is computed from it)
      out['%s/%s' % (name, k)] = np.asarray(model[k], np.float32 if k in BIG else np.float64)
    # the random draws of the graph, in call order (image_ops first, then the scheduled-sampling draws)
    logs = [r for r in tf.RANDOM_LOG if np.asarray(r['value']).dtype.kind == 'f']
    if over['use_knob'] and phase:
      by_shape = lambda shp: [r['value'] for r in logs if r['shape'] == shp]
      f32 = lambda a: np.asarray(a, np.float32)  # lossless: the shim draws in float32
      out[name + '/draw_box_pad'] = f32(by_shape((B, T, 1))[0])
      out[name + '/draw_ctr_shift'] = f32(by_shape((B, T, 2))[0])
      out[name + '/draw_knob_box_u'] = f32(by_shape((B, T, 1))[1])
      out[name + '/draw_knob_segm_u'] = f32(by_shape((B, T, 1))[2])
      noise = by_shape((B, H, W, 1))
      assert len(noise) == T, len(noise)
      out[name + '/draw_segm_noise'] = f32(np.stack([n[..., 0] for n in noise], 1))  # [B,T,H,W]
    print(name, 'loss', float(np.asarray(model['loss'])), 'random draws', len(tf.RANDOM_LOG))
  path = os.path.join(HERE, 'full_model_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, len(out), 'arrays', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
