"""ctypes binding of librecattend_b200.so (the C ABI declared in include/rec_attend_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is
raised.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C rec-attend-public_b200/csrc``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'librecattend_b200.so')

RA_OK = 0
_ERRORS = {-1: 'RA_ERR_INVALID_ARG', -2: 'RA_ERR_CUDA', -3: 'RA_ERR_UNSUPPORTED', -4: 'RA_ERR_NO_DEVICE'}

# controller flags / box slots (include/rec_attend_b200.h)
CTRL_SQUASH, CTRL_FIXED_VAR, CTRL_DYNAMIC_VAR, CTRL_FIXED_GAMMA = 1, 2, 4, 8
BOX_STRIDE = 16
BOX_CTR_Y, BOX_CTR_X, BOX_SIZE_Y, BOX_SIZE_X, BOX_LGVAR_Y, BOX_LGVAR_X = 0, 1, 2, 3, 4, 5
BOX_GAMMA_ATTN, BOX_GAMMA_BOX, BOX_GAMMA_Y = 6, 7, 8
BOX_TL_Y, BOX_TL_X, BOX_BR_Y, BOX_BR_X = 9, 10, 11, 12
LOSS_NAMES = ['box_loss', 'segm_loss', 'conf_loss', 'iou_soft', 'iou_hard', 'wt_cov_soft', 'unwt_cov_soft',
              'wt_cov_hard', 'unwt_cov_hard', 'dice', 'count_acc', 'dic', 'dic_abs', 'loss']
LOSS_COUNT = 16
HUNG_MAX_N = 64

_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_Z = ctypes.c_size_t

# name -> argtypes, in header order
_SIGNATURES = {
    'ra_hungarian_f32': [_P, _I, _I, _I, _P, _P, _P, _P, _P],
    'ra_hungarian_f32_host': [_P, _I, _I, _I, _P, _P, _P, _P],
    'ra_segm_match_f32': [_P, _P, _I, _I, _P, _P, _P, _P],
    'ra_conv3x3_f32': [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    'ra_canvas_conv_f32': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    'ra_conv3x3_umma_plan': [_I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    'ra_conv3x3_umma_plan_split': [_I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    'ra_conv3x3_umma_plan_info': [_I, _I, _I, _I, _I, _I, _P],
    'ra_umma_pack_f16': [_P, ctypes.c_longlong, _I, _I, _P, _P],
    'ra_debug_conv_timeline': [_P],
    'ra_conv3x3_umma_chain_prepare': [_P, _I, _P, _P, _P],
    'ra_conv3x3_umma_chain_run': [_P, _I, _I, _Z, _P, _P],
    'ra_conv3x3_umma_f32': [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    'ra_controller_step_f32': [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P,
                               _P, _P, _P, _P],
    'ra_gaussian_filters_f32': [_P, _I, _I, _I, _I, _P, _P, _P, _P],
    'ra_gaussian_extract_f32': [_P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _I, _P],
    'ra_paste_back_f32': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _Z, _P, _P],
    'ra_score_f32': [_P, _I, _P, _I, _P, _P, _I, _P, _I, _P],
    'ra_gt_box_f32': [_P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P],
    'ra_pairwise_iou_f32': [_P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P],
    'ra_loss_block_f32': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _F, _P, _P],
    'ra_box_gt_step_f32': [_P, _Z, _P, _P, _P, _Z, _I, _I, _I, _I, _P, _I, _P, _P, _P],
    'ra_concat_channels_f32': [_P, _I, _P, _I, _P, _I, _Z, _P, _P],
    'ra_bn_train_block_f32': [_P, _I, _I, _I, _I, _P, _P, _F, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    'ra_random_transformation_f32': [_P, _Z, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    'ra_gt_attn_noise_f32': [_P, _P, _P, _P, _F, _I, _I, _P, _P, _P],
    'ra_knob_greedy_box_f32': [_P, _Z, _P, _I, _I, _I, _I, _P, _I, _P, _P],
    'ra_greedy_iou_box_f32': [_P, _P, _P, _I, _I, _P, _I, _P, _P],
    'ra_box_gt_canvas_f32': [_P, _P, _P, _Z, _I, _I, _I, _I, _P, _P],
    'ra_knob_mix_box_f32': [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    'ra_knob_canvas_f32': [_P, _P, _P, _Z, _P, _I, _P, _Z, _I, _I, _I, _I, _P, _P],
    'ra_bn_train_block_bwd_f32': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P],
    'ra_conv3x3_bwd_weight_f32': [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    'ra_filter_flip_transpose_f32': [_P, _I, _I, _P, _P],
    'ra_subsample2_f32': [_P, _I, _I, _I, _I, _I, _P, _P],
    'ra_iou_loss_bwd_f32': [_P, _Z, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P],
    'ra_conf_loss_bwd_f32': [_P, _P, _I, _I, _I, _F, _P, _P],
    'ra_paste_back_bwd_f32': [_P, _P, _Z, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    'ra_gaussian_extract_bwd_f32': [_P, _I, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    'ra_gaussian_filters_bwd_f32': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P],
    'ra_controller_tape_f32': [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    'ra_controller_tape_layout': [_I, _I, _I, _P],
    'ra_controller_head_bwd_f32': [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P],
    'ra_controller_bwd_f32': [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    'ra_outer_sum_f32': [_P, _Z, _I, _P, _Z, _I, _I, _P, _P, _P],
    'ra_outer_sum_ex_f32': [_P, _Z, _I, _P, _Z, _I, _I, _P, _P, _P, _P],
    'ra_bn_train_block_bwd_grouped_f32': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P],
    'ra_conv3x3_bwd_weight_ex_f32': [_P, _I, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    'ra_paste_back_bwd_ex_f32': [_P, _P, _Z, _I, _Z, _P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P,
                                 _P],
    'ra_gaussian_extract_bwd_ex_f32': [_P, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P,
                                       _P, _P, _P],
    'ra_param_gather_f32': [_P, _P, _P, _P, _I, ctypes.c_longlong, _P],
    'ra_param_scatter_f32': [_P, _P, _P, _P, _I, ctypes.c_longlong, _P],
    'ra_bn_fold_f32': [_P, _P, _P, _P, _P, _I, _I, _F, _P, _P, _P],
    'ra_weight_decay_f32': [_P, _P, _Z, _P, _P, _P],
    'ra_sum_groups_f32': [_P, _I, _Z, _P, _P],
    'ra_add_f32': [_P, _P, _Z, _P],
    'ra_split_channels_f32': [_P, _Z, _I, _I, _P, _P, _I, _P],
    'ra_score_bwd_f32': [_P, _P, _I, _I, _P, _I, _I, _P, _P, _P, _I, _P],
    'ra_knob_box_bwd_f32': [_P, _P, _P, _I, _I, _P, _P],
    'ra_iou_box_coord_bwd_f32': [_P, _P, _P, _P, _I, _I, _F, _P, _P],
    'ra_pairwise_iou_umma_f32': [_P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P],
    'ra_wt_cov_f32': [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    'ra_loss_select_f32': [_P, _P, _P, _F, _F, _P],
    'ra_u8_to_f32': [_P, _Z, _P, _P],
    'ra_fg_head_f32': [_P, _Z, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P],
    'ra_fg_head_bwd_f32': [_P, _P, _Z, _I, _I, _P, _P, _I, _P, _P, _P],
    'ra_adam_step_f32': [_P, _P, _P, _P, _P, _Z, _F, _F, _F, _F, _F, _F, _I, _P],
    'ra_postprocess_f32': [_P, _P, _P, _I, _I, _I, _I, ctypes.c_double, _F, _P, _P, _P, _P, _P, _P],
}
EXPORTED = sorted(list(_SIGNATURES) + ['ra_version', 'ra_device_count', 'ra_last_error', 'ra_launch_count',
                                         'ra_pairwise_iou_workspace', 'ra_postprocess_workspace',
                                         'ra_bn_train_workspace', 'ra_fg_head_workspace',
                                         'ra_bn_train_block_bwd_workspace', 'ra_conv3x3_bwd_weight_workspace',
                                         'ra_iou_loss_bwd_workspace', 'ra_paste_back_bwd_workspace',
                                         'ra_gaussian_extract_bwd_workspace', 'ra_controller_tape_floats',
                                         'ra_bn_train_block_bwd_grouped_workspace', 'ra_weight_decay_workspace',
                                         'ra_pairwise_iou_umma_workspace', 'ra_conv3x3_umma_chain_desc_bytes',
                                         'ra_outer_sum_workspace', 'ra_conv3x3_umma_set_f16'])

_lib = None
TAG = ''  # set by the model code so that bench.py can attribute kernel time to a sub-network


class RecAttendError(RuntimeError):
  pass


def lib():
  """The loaded library (loads on first use; raises if the .so has not been built)."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise RecAttendError('librecattend_b200.so is not built ({}); there is no CPU fallback — run '
                           '`make -C rec-attend-public_b200/csrc` or __graft_entry__.build()'.format(LIB_PATH))
    l = ctypes.CDLL(LIB_PATH)
    for name, args in _SIGNATURES.items():
      fn = getattr(l, name)
      fn.argtypes = args
      fn.restype = _I
    l.ra_version.restype = _I
    l.ra_conv3x3_umma_set_f16.argtypes = [_I]
    l.ra_conv3x3_umma_set_f16.restype = _I
    l.ra_device_count.restype = _I
    l.ra_last_error.restype = ctypes.c_char_p
    l.ra_launch_count.restype = ctypes.c_ulonglong
    l.ra_pairwise_iou_workspace.argtypes = [_I, _I, _I, _I]
    l.ra_pairwise_iou_workspace.restype = _Z
    l.ra_postprocess_workspace.argtypes = [_I, _I]
    l.ra_postprocess_workspace.restype = _Z
    l.ra_bn_train_workspace.argtypes = [_I, _I, _I, _I]
    l.ra_bn_train_workspace.restype = _Z
    l.ra_fg_head_workspace.argtypes = []
    l.ra_fg_head_workspace.restype = _Z
    l.ra_bn_train_block_bwd_workspace.argtypes = [_I, _I, _I, _I, _I]
    l.ra_bn_train_block_bwd_workspace.restype = _Z
    l.ra_conv3x3_bwd_weight_workspace.argtypes = [_I, _I, _I, _I, _I, _I]
    l.ra_conv3x3_bwd_weight_workspace.restype = _Z
    l.ra_iou_loss_bwd_workspace.argtypes = [_I, _I, _I]
    l.ra_iou_loss_bwd_workspace.restype = _Z
    l.ra_paste_back_bwd_workspace.argtypes = [_I, _I, _I, _I]
    l.ra_paste_back_bwd_workspace.restype = _Z
    l.ra_gaussian_extract_bwd_workspace.argtypes = [_I, _I, _I, _I]
    l.ra_gaussian_extract_bwd_workspace.restype = _Z
    l.ra_bn_train_block_bwd_grouped_workspace.argtypes = [_I, _I, _I, _I, _I, _I]
    l.ra_bn_train_block_bwd_grouped_workspace.restype = _Z
    l.ra_pairwise_iou_umma_workspace.argtypes = [_I, _I, _I, _I]
    l.ra_pairwise_iou_umma_workspace.restype = _Z
    l.ra_outer_sum_workspace.argtypes = [_I, _I, _I]
    l.ra_outer_sum_workspace.restype = _Z
    l.ra_conv3x3_umma_chain_desc_bytes.argtypes = [_I]
    l.ra_conv3x3_umma_chain_desc_bytes.restype = _Z
    l.ra_weight_decay_workspace.argtypes = []
    l.ra_weight_decay_workspace.restype = _Z
    l.ra_controller_tape_floats.argtypes = [_I, _I, _I, _I]
    l.ra_controller_tape_floats.restype = _Z
    _lib = l
  return _lib


def check(rc, name):
  if rc != RA_OK:
    msg = lib().ra_last_error().decode('utf-8', 'replace')
    raise RecAttendError('{} failed: {} ({}) {}'.format(name, _ERRORS.get(rc, rc), rc, msg))


def call(name, *args):
  check(getattr(lib(), name)(*args), name)
