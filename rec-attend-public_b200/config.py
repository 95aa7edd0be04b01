"""Model option dictionaries — the config contract of the hot path.

Same keys as the reference's ``ModelArgsParser.make_opt`` (full_model_train.py:581-658)
and ``BoxModelArgsParser.make_opt`` (box_model_train.py:412-452); the three architectures
are the ones written down in the reference's run scripts (run_cvppp.sh:37-72,
run_kitti.sh:44-111, run_cityscapes.sh:62-110).  H, W, T come from BASELINE.json's
configs (synthetic resolutions), not from cmd_args_parser.py:18-63.
"""
import copy

_COMMON = {
    'inp_depth': 3,
    'padding': 16,
    'filter_height': 48,
    'filter_width': 48,
    'ctrl_cnn_filter_size': [3] * 8,
    'ctrl_rnn_hid_dim': 256,
    'attn_cnn_filter_size': [3] * 6,
    'attn_dcnn_filter_size': [3] * 7,
    'attn_dcnn_pool': [2, 1, 2, 1, 2, 1, 1],
    'attn_cnn_pool': [1, 2, 1, 2, 1, 2],
    'num_ctrl_mlp_layers': 1,
    'ctrl_mlp_dim': 256,
    'mlp_dropout': None,
    'weight_decay': 5e-5,
    'base_learn_rate': 0.001,
    'learn_rate_decay': 0.85,
    'steps_per_learn_rate_decay': 5000,
    'loss_mix_ratio': 1.0,
    'segm_loss_fn': 'iou',
    'box_loss_fn': 'iou',
    'use_bn': True,
    'attn_box_padding_ratio': 0.2,
    'use_knob': False,  # full_model_eval.py:172-174 (eval forward)
    'knob_decay': 0.5,
    'knob_base': 1.0,
    'steps_per_knob_decay': 1500,
    'knob_box_offset': 100,
    'knob_segm_offset': 8000,
    'knob_use_timescale': True,
    'gt_box_ctr_noise': 0.05,
    'gt_box_pad_noise': 0.1,
    'gt_segm_noise': 0.3,
    'squash_ctrl_params': False,
    'clip_gradient': 1.0,
    'fixed_order': False,
    'fixed_var': False,
    'num_ctrl_rnn_iter': 5,
    'num_glimpse_mlp_layers': 2,
    'pretrain_ctrl_net': None,
    'pretrain_attn_net': None,
    'pretrain_net': None,
    'freeze_ctrl_cnn': False,
    'freeze_ctrl_rnn': False,
    'freeze_ctrl_mlp': False,
    'freeze_attn_net': False,
    'stop_canvas_grad': True,
    'use_iou_box': False,
    'disable_overwrite': False,  # never passed by the run scripts
    'ctrl_add_inp': True,
    'ctrl_add_canvas': True,
    'attn_add_inp': True,
    'attn_add_canvas': True,
    'rnd_hflip': False,
    'rnd_vflip': False,
    'rnd_transpose': False,
    'rnd_colour': False,
    'finetune': False,
}

_ARCH = {
    # run_cvppp.sh:37-72
    'cvppp': {
        'ctrl_cnn_depth': [8, 8, 16, 16, 32, 32, 64, 64],
        'ctrl_cnn_pool': [1, 2, 1, 2, 1, 2, 2, 2],
        'attn_cnn_depth': [8, 8, 16, 16, 32, 32],
        'attn_dcnn_depth': [32, 32, 16, 16, 8, 8, 1],
        'attn_cnn_skip': '1,1,1',
        'add_skip_conn': False,
        'fixed_gamma': True,
        'dynamic_var': False,
        'add_d_out': False, 'add_y_out': False,
        'attn_add_d_out': False, 'attn_add_y_out': False,
        'ctrl_add_d_out': False, 'ctrl_add_y_out': False,
        'num_semantic_classes': 1,
    },
    # run_kitti.sh:68-111
    'kitti': {
        'ctrl_cnn_depth': [16, 16, 32, 32, 64, 64, 64, 64],
        'ctrl_cnn_pool': [2, 2, 1, 2, 1, 2, 1, 2],
        'attn_cnn_depth': [16, 32, 32, 64, 64, 96],
        'attn_dcnn_depth': [64, 64, 32, 32, 16, 16, 1],
        'attn_cnn_skip': '1,0,1,0,1,0,1,0',
        'add_skip_conn': True,
        'fixed_gamma': False,
        'dynamic_var': True,
        'add_d_out': True, 'add_y_out': True,
        'attn_add_d_out': True, 'attn_add_y_out': True,
        'ctrl_add_d_out': True, 'ctrl_add_y_out': True,
        'num_semantic_classes': 1,
    },
    # run_cityscapes.sh:62-110
    'cityscapes': {
        'ctrl_cnn_depth': [16, 16, 32, 32, 64, 64, 64, 64],
        'ctrl_cnn_pool': [2, 2, 1, 2, 1, 2, 1, 2],
        'attn_cnn_depth': [16, 32, 32, 64, 64, 96],
        'attn_dcnn_depth': [64, 64, 32, 32, 16, 16, 1],
        'attn_cnn_skip': '1,0,1,0,1,0,1,0',
        'add_skip_conn': True,
        'fixed_gamma': True,
        'dynamic_var': True,
        'use_iou_box': True,
        'add_d_out': True, 'add_y_out': True,
        'attn_add_d_out': True, 'attn_add_y_out': True,
        'ctrl_add_d_out': True, 'ctrl_add_y_out': True,
        'num_semantic_classes': 9,
    },
}


def full_model_opt(arch, inp_height, inp_width, timespan, **overrides):
  """opt dict for ``full_model.get_model`` (full_model_train.py:581-658)."""
  opt = copy.deepcopy(_COMMON)
  opt.update(copy.deepcopy(_ARCH[arch]))
  opt.update({'inp_height': inp_height, 'inp_width': inp_width, 'timespan': timespan, 'arch': arch})
  opt.update(overrides)
  return opt


def box_model_opt(inp_height, inp_width, timespan, **overrides):
  """opt dict for ``box_model.get_model`` with the KITTI box flags (run_kitti.sh:44-60,
  box_model_train.py:412-452)."""
  opt = {k: copy.deepcopy(v) for k, v in _COMMON.items() if not k.startswith(('attn_cnn', 'attn_dcnn', 'knob'))}
  opt.update({
      'ctrl_cnn_depth': [16, 16, 32, 32, 64, 64, 64, 64],
      'ctrl_cnn_pool': [1, 2, 1, 2, 1, 2, 2, 2],
      'dynamic_var': True,
      'fixed_var': False,
      'add_d_out': True,
      'add_y_out': True,
      'num_semantic_classes': 1,
      'pretrain_cnn': None,
      'learn_rate_decay': 0.9,
      'inp_height': inp_height, 'inp_width': inp_width, 'timespan': timespan, 'arch': 'kitti_box',
  })
  opt.update(overrides)
  return opt


_FG_ARCH = {
    # fg_model_train.py:425-441 (parser defaults; skip connections only with --add_skip_conn)
    'default': {
        'cnn_depth': [8, 8, 16, 16, 32, 32, 64, 64, 128, 128],
        'cnn_pool': [1, 2, 1, 2, 1, 2, 1, 2, 1, 2],
        'cnn_skip_mask': [1, 0, 0, 0, 0, 0, 1, 0, 1, 0],
        'dcnn_depth': [128, 128, 64, 64, 32, 32, 16, 16, 8, 8, 1],
        'dcnn_pool': [2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 1],
        'dcnn_skip_mask': [0, 1, 0, 1, 0, 0, 0, 0, 0, 1],
        'add_skip_conn': False, 'add_orientation': False, 'num_semantic_classes': 1, 'segm_loss_fn': 'iou',
    },
    # run_kitti.sh:13-27 (`--cnn_skip` / `--dcnn_skip` are argparse prefixes of --cnn_skip_mask / --dcnn_skip_mask)
    'kitti': {
        'cnn_depth': [32, 64, 64, 96, 96, 128, 128, 128, 128, 128, 128, 128, 128, 256, 256, 256, 256, 512],
        'cnn_pool': [1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 2],
        'cnn_skip_mask': [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1],
        'dcnn_depth': [256, 256, 128, 128, 96, 96, 64, 64, 32, 32, 9],
        'dcnn_pool': [2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 1],
        'dcnn_skip_mask': [1, 0, 1, 0, 1, 0, 0, 0, 0, 1],
        'add_skip_conn': True, 'add_orientation': True, 'num_semantic_classes': 1, 'segm_loss_fn': 'bce',
    },
    # run_cityscapes.sh:8-30
    'cityscapes': {
        'cnn_depth': [64, 96, 96, 128, 128, 192, 192, 256, 256, 256, 256, 256, 256, 256, 256, 512, 512, 512, 512, 512],
        'cnn_pool': [1, 2, 1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 2],
        'cnn_skip_mask': [1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0],
        'dcnn_depth': [512, 512, 256, 256, 192, 192, 128, 128, 96, 96, 64, 64, 17],
        'dcnn_pool': [2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 1],
        'dcnn_skip_mask': [1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 0],
        'add_skip_conn': True, 'add_orientation': True, 'num_semantic_classes': 9, 'segm_loss_fn': 'bce',
    },
}


def fg_model_opt(arch, inp_height, inp_width, **overrides):
  """opt dict for ``fg_model.get_model`` (fg_model_train.py:470-499), the FCN that produces d_in / y_in."""
  a = copy.deepcopy(_FG_ARCH[arch])
  opt = {
      'inp_height': inp_height, 'inp_width': inp_width, 'inp_depth': 3, 'padding': 16,
      'cnn_filter_size': [3] * len(a['cnn_depth']), 'dcnn_filter_size': [3] * len(a['dcnn_depth']),
      'weight_decay': 5e-5, 'use_bn': True, 'rnd_hflip': False, 'rnd_vflip': False, 'rnd_transpose': False,
      'rnd_colour': False, 'base_learn_rate': 1e-3, 'learn_rate_decay': 0.96, 'steps_per_learn_rate_decay': 5000,
      'num_orientation_classes': 8, 'optimizer': 'adam', 'arch': 'fg_' + arch,
  }
  opt.update(a)
  opt['cnn_skip_mask'] = [bool(v) for v in opt['cnn_skip_mask']]
  opt['dcnn_skip_mask'] = [bool(v) for v in opt['dcnn_skip_mask']]
  opt.update(overrides)
  return opt


def fg_skip_wiring(opt):
  """fg_model.py:122-146: which activation feeds each DCNN layer as its skip input.  Returns a list, one entry
  per DCNN layer: None or the index j into [x] + h_cnn[:-1] (j = 0 is the input image, j >= 1 is h_cnn[j-1]),
  plus the list of skip channel counts.  The CNN-side mask picks layers in order, the DCNN-side mask consumes them
  from the LAST one backwards; DCNN layer 0 never has a skip."""
  n_d = len(opt['dcnn_depth'])
  if not opt.get('add_skip_conn', False):
    return [None] * n_d, [0] * n_d
  cnn_channels = [opt['inp_depth']] + list(opt['cnn_depth'])
  n_all = len(opt['cnn_depth'])  # len([x] + h_cnn[:-1])
  cnn_mask = opt.get('cnn_skip_mask', opt.get('cnn_skip', [True] * n_all))
  picked = [j for j, sk in zip(range(n_all), cnn_mask) if sk]
  dcnn_mask = opt.get('dcnn_skip_mask', list(cnn_mask)[::-1])
  src, ch = [None], [0]
  counter = len(picked) - 1
  for sk in dcnn_mask:
    if sk:
      if counter < 0:
        raise ValueError('dcnn_skip_mask asks for more skips than cnn_skip_mask provides (IndexError in the reference)')
      src.append(picked[counter])
      ch.append(cnn_channels[picked[counter]])
      counter -= 1
    else:
      src.append(None)
      ch.append(0)
  if len(src) < n_d:
    raise ValueError('dcnn_skip_mask needs at least one entry per DCNN layer after the first ({} < {}; IndexError in '
                     'nnlib.run_dcnn)'.format(len(src) - 1, n_d - 1))
  return src[:n_d], ch[:n_d]  # nnlib.dcnn only reads the first nlayers entries (run_cityscapes.sh passes one extra)


# BASELINE.json configs (index = position in BASELINE.json "configs")
BASELINE_CONFIGS = [
    {'name': 'cvppp_128x128_T8_B1', 'model': 'full', 'arch': 'cvppp', 'H': 128, 'W': 128, 'T': 8, 'B': 1},
    {'name': 'cvppp_256x256_T20_B16', 'model': 'full', 'arch': 'cvppp', 'H': 256, 'W': 256, 'T': 20, 'B': 16},
    {'name': 'kitti_256x512_T20_B32', 'model': 'full', 'arch': 'kitti', 'H': 256, 'W': 512, 'T': 20, 'B': 32},
    {'name': 'cityscapes_512x1024_T32_B8', 'model': 'full', 'arch': 'cityscapes', 'H': 512, 'W': 1024, 'T': 32,
     'B': 8},
    {'name': 'kitti_box_256x512_T20_B64', 'model': 'box', 'arch': 'kitti_box', 'H': 256, 'W': 512, 'T': 20, 'B': 64},
]


def baseline_opt(idx, **overrides):
  c = BASELINE_CONFIGS[idx]
  if c['model'] == 'box':
    return box_model_opt(c['H'], c['W'], c['T'], **overrides)
  return full_model_opt(c['arch'], c['H'], c['W'], c['T'], **overrides)


def input_depths(opt):
  """(ctrl CNN input depth, attn CNN input depth, semantic classes) — full_model.py:240-258."""
  nsc = opt.get('num_semantic_classes', 1)
  d = opt['inp_depth'] + 1
  if opt.get('add_d_out', False):
    d += 8 + nsc
  return d, d, nsc
