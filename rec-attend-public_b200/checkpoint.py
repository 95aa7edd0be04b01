"""Weight-file / checkpoint interchange of the hot path — host-side numpy only, no kernels.

What the reference does around the path, restated on the flat weight dict this package already uses
(`FullModel.load_weights` / `export_weights`, reference key schema and TensorFlow layouts):

* ``full_model_read.py:14-82`` / ``box_model_read.py:13-58`` export a trained graph into a flat ``weights.h5``:
  conv / mlp ``{net}_{w,b}_{i}``, per-(layer, step) BN ``{net}_{i}_{t}_{beta,gamma}``, the twelve ``ctrl_lstm_*``
  tensors.  The BN EMA shadows are NOT exported.  -> `weight_file_keys`, `save_weights`, `load_weights`.
* ``full_model.py:271-284,315-326,355-363,387-395,421-434,464-470,504-518`` / ``box_model.py:182-219,250-330`` read
  such files back as the initial values of a new graph (the two-stage hand-off box_model -> full_model of
  run_*.sh), with the ``freeze_*`` flags deciding which tensors stay trainable (``nnlib.py:185,455,527``).
  -> `apply_pretrained`.
* ``nnlib.py:41-62,80-86,519-610`` initialise whatever is not read from a file.  -> `reference_init`.
* ``utils/saver.py:12-108`` keeps ``model_opt.yaml`` + the latest two ``model.ckpt-<step>`` files of a run folder.
  -> `Saver` (same folder protocol; the TF checkpoint container is replaced by ``.npz`` holding the weights, the EMA
  shadows and the Adam slots).

Containers: ``.h5`` when h5py is importable (it is not in this image), otherwise ``.npz`` with the same keys.
"""
import fnmatch
import os

import numpy as np

LSTM_KEYS = ['w_xi', 'w_hi', 'b_i', 'w_xf', 'w_hf', 'b_f', 'w_xu', 'w_hu', 'b_u', 'w_xo', 'w_ho', 'b_o']
kModelOptFilename = 'model_opt.yaml'  # utils/saver.py:7
kMaxToKeep = 2  # utils/saver.py:9


class CheckpointError(RuntimeError):
  pass


# ----------------------------------------------------------------------------- key schema
def _cnn_keys(net, nlayers, timespan, with_bn=True):
  keys = []
  for ii in range(nlayers):
    keys += ['{}_w_{}'.format(net, ii), '{}_b_{}'.format(net, ii)]
    if with_bn:
      for tt in range(timespan):
        keys += ['{}_{}_{}_{}'.format(net, ii, tt, w) for w in ('beta', 'gamma')]
  return keys


def _mlp_keys(net, nlayers):
  return [k for ii in range(nlayers) for k in ('{}_w_{}'.format(net, ii), '{}_b_{}'.format(net, ii))]


def weight_file_keys(opt, model='full'):
  """The keys ``full_model_read.read`` (full_model_read.py:32-70) or ``box_model_read.read``
  (box_model_read.py:31-52) put into weights.h5, in the reference's order."""
  T = opt['timespan']
  use_bn = bool(opt.get('use_bn', True))
  keys = _cnn_keys('ctrl_cnn', len(opt['ctrl_cnn_filter_size']), T, use_bn)
  keys += _mlp_keys('ctrl_mlp', opt['num_ctrl_mlp_layers'])
  keys += _mlp_keys('glimpse_mlp', opt['num_glimpse_mlp_layers'])
  keys += _mlp_keys('score_mlp', 1)
  keys += ['ctrl_lstm_' + w for w in LSTM_KEYS]
  if model == 'full':
    keys += _cnn_keys('attn_cnn', len(opt['attn_cnn_filter_size']), T, use_bn)
    keys += _cnn_keys('attn_dcnn', len(opt['attn_dcnn_filter_size']), T, use_bn)
  elif model != 'box':
    raise CheckpointError("model must be 'full' or 'box'")
  return keys


# ----------------------------------------------------------------------------- containers
def _h5py():
  try:
    import h5py  # noqa: WPS433 (optional dependency, absent in this image)
    return h5py
  except ImportError:
    return None


def save_weights(fname, weights, keys=None):
  """``full_model_read.save`` (full_model_read.py:85-91): one dataset per key.  keys=None writes every entry of
  `weights` (incl. EMA shadows); pass `weight_file_keys(opt)` for exactly the reference's file."""
  keys = list(weights) if keys is None else list(keys)
  missing = [k for k in keys if k not in weights]
  if missing:
    raise CheckpointError('weights lack {} of the requested keys, first: {}'.format(len(missing), missing[0]))
  arrays = {k: np.ascontiguousarray(np.asarray(weights[k], np.float32)) for k in keys}
  if fname.endswith(('.h5', '.hdf5')):
    h5py = _h5py()
    if h5py is None:
      raise CheckpointError('h5py is not installed: use a .npz file name (same keys, same layouts)')
    with h5py.File(fname, 'w') as h5f:
      for k, v in arrays.items():
        h5f[k] = v
  else:
    with open(fname, 'wb') as f:  # np.savez would append ".npz" to other extensions
      np.savez(f, **arrays)
  return fname


def load_weights(fname):
  """Read a flat weight file (``h5f[key][:]`` of full_model.py:274-283) into a dict of fp32 arrays."""
  if not os.path.exists(fname):
    raise CheckpointError('weight file {} does not exist'.format(fname))
  if fname.endswith(('.h5', '.hdf5')):
    h5py = _h5py()
    if h5py is None:
      raise CheckpointError('h5py is not installed: cannot read {}; convert it to .npz'.format(fname))
    with h5py.File(fname, 'r') as h5f:
      return {k: np.asarray(h5f[k][:], np.float32) for k in h5f}
  with np.load(fname) as z:
    return {k: np.asarray(z[k], np.float32) for k in z.files}


def _as_dict(src):
  if src is None or isinstance(src, dict):
    return src
  return load_weights(src)


# ----------------------------------------------------------------------------- initialisers
def _trunc_normal(rng, shape, std=0.01):
  """tf.truncated_normal_initializer(stddev=0.01), nnlib.py:53-54: values beyond two sigma are redrawn."""
  v = rng.standard_normal(shape)
  bad = np.abs(v) > 2.0
  while bad.any():
    v[bad] = rng.standard_normal(int(bad.sum()))
    bad = np.abs(v) > 2.0
  return (v * std).astype(np.float32)


def reference_init(opt, seed=0, model='full'):
  """The reference's own initial values for every tensor of the schema (shapes from `synthetic.make_weights`):
  truncated normal, sigma 0.01, for conv / mlp weights AND biases (`weight_variable` default, nnlib.py:53-56,
  200-201,462-463) and the LSTM matrices; LSTM biases 0 except b_f = 1 (nnlib.py:546,566,586,606); BN beta 0,
  gamma 1 (nnlib.py:88-91); EMA shadows 0 (tf.train.ExponentialMovingAverage over tensors starts at zero).
  TensorFlow's random stream is not reproducible, so only the distribution matches.  NOTE: with these values every
  decode step emits the same box — parity tests use `synthetic.make_weights` instead."""
  from . import synthetic
  rng = np.random.default_rng(seed)
  out = {}
  for k, v in synthetic.make_weights(opt, seed=0, model=model).items():
    shape = np.asarray(v).shape
    if k.endswith('_gamma'):
      out[k] = np.ones(shape, np.float32)
    elif k.endswith(('_beta', '_ema_mean', '_ema_var')):
      out[k] = np.zeros(shape, np.float32)
    elif k.startswith('ctrl_lstm_b_'):
      out[k] = np.full(shape, 1.0 if k.endswith('b_f') else 0.0, np.float32)
    else:
      out[k] = _trunc_normal(rng, shape)
  return out


# ----------------------------------------------------------------------------- pretrained hand-off
def _take(dst, src, keys, what):
  for k in keys:
    if k not in src:
      raise CheckpointError('pretrained file for {} lacks "{}" (the reference would raise KeyError here)'.format(what, k))
    a = np.asarray(src[k], np.float32)
    if k in dst and tuple(np.asarray(dst[k]).shape) != tuple(a.shape):
      raise CheckpointError('pretrained "{}" has shape {} but the model expects {}'.format(
          k, tuple(a.shape), tuple(np.asarray(dst[k]).shape)))
    dst[k] = a.copy()


def _wb(net, nlayers):
  return [k for ii in range(nlayers) for k in ('{}_w_{}'.format(net, ii), '{}_b_{}'.format(net, ii))]


def apply_pretrained(opt, weights, pretrain_net=None, pretrain_ctrl_net=None, pretrain_attn_net=None, pretrain_cnn=None,
                     model='full'):
  """Initial values + trainability of a new model, as the reference builds them.

  weights: the freshly initialised dict (`reference_init` or any full set).  pretrain_*: dicts or file names; when
  None they default to opt['pretrain_net'] etc.  Returns (new_weights, frozen_keys).

  model='full' (full_model.py): controller CNN (+ its per-step beta / gamma), controller LSTM, glimpse MLP and
  controller MLP come from ``pretrain_net or pretrain_ctrl_net`` (:271,315,355,387); attention CNN / DCNN (+ BN) from
  ``pretrain_net or pretrain_attn_net`` (:421,504); the score MLP from ``pretrain_net`` only (:464).  Frozen (w and b
  only — the BN `frozen` flag is computed but never reaches batch_norm, nnlib.py:231-249):
  freeze_ctrl_cnn -> ctrl_cnn, freeze_ctrl_rnn -> ctrl_lstm AND glimpse_mlp (:363), freeze_ctrl_mlp -> ctrl_mlp,
  freeze_attn_net -> attn_cnn + attn_dcnn; the flags apply whether or not a file was read.

  model='box' (box_model.py): the first `n` controller-CNN layers found in ``pretrain_net or pretrain_cnn`` under the
  prefix ``attn_cnn_`` / ``cnn_`` / ``ctrl_cnn_`` (:182-219; the prefix of the LAST layer found names all of them),
  frozen when freeze_pretrain_cnn (default True, :47-50); LSTM, glimpse MLP and controller MLP from ``pretrain_net``
  (:250-330), never frozen.  EMA shadows are never read from a file (they are not in it)."""
  T = opt['timespan']
  use_bn = bool(opt.get('use_bn', True))
  out = {k: np.asarray(v, np.float32).copy() for k, v in weights.items()}
  frozen = []
  g = lambda v, key: _as_dict(v if v is not None else opt.get(key))
  p_net = g(pretrain_net, 'pretrain_net')
  n_ccnn = len(opt['ctrl_cnn_filter_size'])
  n_cmlp, n_gmlp = opt['num_ctrl_mlp_layers'], opt['num_glimpse_mlp_layers']
  lstm = ['ctrl_lstm_' + w for w in LSTM_KEYS]
  if model == 'full':
    p_ctrl = p_net if p_net is not None else g(pretrain_ctrl_net, 'pretrain_ctrl_net')
    p_attn = p_net if p_net is not None else g(pretrain_attn_net, 'pretrain_attn_net')
    n_acnn, n_adcnn = len(opt['attn_cnn_filter_size']), len(opt['attn_dcnn_filter_size'])
    if p_ctrl is not None:
      _take(out, p_ctrl, _cnn_keys('ctrl_cnn', n_ccnn, T, use_bn), 'the controller CNN')
      _take(out, p_ctrl, lstm, 'the controller LSTM')
      _take(out, p_ctrl, _mlp_keys('glimpse_mlp', n_gmlp), 'the glimpse MLP')
      _take(out, p_ctrl, _mlp_keys('ctrl_mlp', n_cmlp), 'the controller MLP')
    if p_attn is not None:
      _take(out, p_attn, _cnn_keys('attn_cnn', n_acnn, T, use_bn), 'the attention CNN')
      _take(out, p_attn, _cnn_keys('attn_dcnn', n_adcnn, T, use_bn), 'the attention DCNN')
    if p_net is not None:
      _take(out, p_net, _mlp_keys('score_mlp', 1), 'the score MLP')
    if opt.get('freeze_ctrl_cnn', False):
      frozen += _wb('ctrl_cnn', n_ccnn)
    if opt.get('freeze_ctrl_rnn', False):
      frozen += lstm + _wb('glimpse_mlp', n_gmlp)
    if opt.get('freeze_ctrl_mlp', False):
      frozen += _wb('ctrl_mlp', n_cmlp)
    if opt.get('freeze_attn_net', False):
      frozen += _wb('attn_cnn', n_acnn) + _wb('attn_dcnn', n_adcnn)
  elif model == 'box':
    p_cnn = p_net if p_net is not None else g(pretrain_cnn, 'pretrain_cnn')
    if p_cnn is not None:
      n_pt, prefix = 0, None
      for ii in range(n_ccnn):
        for pre in ('attn_', '', 'ctrl_'):
          if '{}cnn_w_{}'.format(pre, ii) in p_cnn:
            prefix = pre
            n_pt += 1
            break
      for ii in range(n_pt):
        src_keys = ['{}cnn_w_{}'.format(prefix, ii), '{}cnn_b_{}'.format(prefix, ii)]
        dst_keys = ['ctrl_cnn_w_{}'.format(ii), 'ctrl_cnn_b_{}'.format(ii)]
        if use_bn:
          for tt in range(T):
            for w in ('beta', 'gamma'):
              src_keys.append('{}cnn_{}_{}_{}'.format(prefix, ii, tt, w))
              dst_keys.append('ctrl_cnn_{}_{}_{}'.format(ii, tt, w))
        renamed = {}
        for s, d in zip(src_keys, dst_keys):
          if s not in p_cnn:
            raise CheckpointError('pretrained CNN file lacks "{}"'.format(s))
          renamed[d] = p_cnn[s]
        _take(out, renamed, dst_keys, 'the controller CNN')
      if opt.get('freeze_pretrain_cnn', True):
        frozen += _wb('ctrl_cnn', n_pt)
    if p_net is not None:
      _take(out, p_net, lstm, 'the controller LSTM')
      _take(out, p_net, _mlp_keys('glimpse_mlp', n_gmlp), 'the glimpse MLP')
      _take(out, p_net, _mlp_keys('ctrl_mlp', n_cmlp), 'the controller MLP')
  else:
    raise CheckpointError("model must be 'full' or 'box'")
  return out, sorted(set(frozen))


# ----------------------------------------------------------------------------- run-folder checkpoints
class Saver(object):
  """utils/saver.py:12-108 on .npz files: ``model_opt.yaml`` written at construction, ``save(state, global_step)``
  -> ``model.ckpt-<step>.npz`` keeping the latest two, ``get_latest_ckpt`` / ``get_ckpt_info`` / ``restore``.
  `state` is a dict of arrays — by convention the weight dict plus ``adam/m``, ``adam/v`` (flat slots) and
  ``global_step`` (see `pack_state`)."""

  def __init__(self, folder, model_opt=None):
    os.makedirs(folder, exist_ok=True)
    self.folder = folder
    if model_opt is not None:
      self.save_opt(os.path.join(folder, kModelOptFilename), model_opt)

  def save_opt(self, fname, opt):
    import yaml
    with open(fname, 'w') as f:
      yaml.safe_dump(_plain(opt), f, default_flow_style=False)

  def _ckpts(self):
    out = []
    for fn in os.listdir(self.folder):
      if fnmatch.fnmatch(fn, 'model.ckpt-*.npz'):
        out.append((int(fn[len('model.ckpt-'):-len('.npz')]), os.path.join(self.folder, fn)))
    return sorted(out)

  def save(self, state, global_step):
    fname = os.path.join(self.folder, 'model.ckpt-{}.npz'.format(int(global_step)))
    tmp = fname + '.tmp'
    with open(tmp, 'wb') as f:
      np.savez(f, **{k: np.asarray(v) for k, v in state.items()})
    os.replace(tmp, fname)  # a crash never leaves a truncated "latest" checkpoint
    for _, old in self._ckpts()[:-kMaxToKeep]:
      os.remove(old)
    return fname

  def get_latest_ckpt(self):
    c = self._ckpts()
    if not c:
      raise CheckpointError('No checkpoint file found.')  # utils/saver.py:50
    return c[-1][1], c[-1][0]

  def get_ckpt_info(self):
    if not os.path.exists(self.folder):
      raise CheckpointError('Folder "{}" does not exist'.format(self.folder))
    opt_fname = os.path.join(self.folder, kModelOptFilename)
    model_opt = None
    if os.path.exists(opt_fname):
      import yaml
      with open(opt_fname) as f:
        model_opt = yaml.safe_load(f)
    ckpt_fname, step = self.get_latest_ckpt()
    return {'ckpt_fname': ckpt_fname, 'model_opt': model_opt, 'step': step,
            'model_id': os.path.basename(self.folder.rstrip('/'))}

  def restore(self, ckpt_fname=None):
    if ckpt_fname is None:
      ckpt_fname = self.get_latest_ckpt()[0]
    with np.load(ckpt_fname) as z:
      return {k: z[k] for k in z.files}


def _plain(v):
  """yaml-safe copy of an opt dict (numpy scalars / arrays -> Python)."""
  if isinstance(v, dict):
    return {str(k): _plain(x) for k, x in v.items()}
  if isinstance(v, (list, tuple)):
    return [_plain(x) for x in v]
  if isinstance(v, np.ndarray):
    return v.tolist()
  if isinstance(v, np.generic):
    return v.item()
  return v


def pack_state(weights, adam_m=None, adam_v=None, global_step=0):
  """Everything ``tf.train.Saver(tf.all_variables())`` would hold for this path (utils/saver.py:28-29): weights,
  BN EMA shadows, the Adam slots (flat buckets of `optim.FlatParams` order) and the step counter."""
  st = {k: np.asarray(v, np.float32) for k, v in weights.items()}
  if adam_m is not None:
    st['adam/m'] = np.asarray(adam_m, np.float32)
  if adam_v is not None:
    st['adam/v'] = np.asarray(adam_v, np.float32)
  st['global_step'] = np.asarray(global_step, np.int64)
  return st


def unpack_state(state):
  """Inverse of `pack_state`: (weights, adam_m or None, adam_v or None, global_step)."""
  w = {k: np.asarray(v, np.float32) for k, v in state.items() if k not in ('adam/m', 'adam/v', 'global_step')}
  return w, state.get('adam/m'), state.get('adam/v'), int(state['global_step']) if 'global_step' in state else 0
