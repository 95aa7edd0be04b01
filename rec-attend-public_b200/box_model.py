"""Host side of the controller-only model — the drop-in for ``box_model.get_model`` (box_model.py:11),
eval mode, training-mode forward (batch-statistics BN, nnlib.py:96-119) and the training step
(box_model.py:635-652): the same controller CNN + glimpse LSTM + box head as the full model, no patch CNN / mask
head; the canvas is driven by the greedily matched ground-truth masks in eval too (box_model.py:484-504).
BASELINE config 5.  Same conventions as full_model.py (opt dict, named inputs/outputs, weight keys of
box_model_read.py:31-52); every FLOP runs in librecattend_b200.so.
"""
import numpy as np
import torch

from . import _lib, ops
from .full_model import _ModelBase, capture_graph


class BoxModel(_ModelBase):

  def __init__(self, opt, device=None):
    super(BoxModel, self).__init__(opt, device)
    if self.opt.get('num_semantic_classes', 1) != 1 and not self.add_d:
      raise _lib.RecAttendError('multi-class score head is not used by the shipped box configs')
    if self.opt['box_loss_fn'] != 'iou':
      raise _lib.RecAttendError("only box_loss_fn='iou' works in the reference (SURVEY §9.13)")
    self.min_padding = 10.0  # modellib.get_gt_box default (box_model.py:386-387)

  def _fixed_var(self):
    return bool(self.opt.get('fixed_var', True))  # box_model.py:58-61

  def _fixed_gamma(self):
    return True  # the box model has no attention / mask gains

  def load_weights(self, weights):
    self._begin_load(weights)
    self.w = self._load_controller(weights)
    if self.w['score_mlp_w_0'].numel() != self.Hd:
      raise _lib.RecAttendError('box model score MLP takes the controller state only (box_model.py:359-363)')
    self._set_wd_term(weights)
    return self

  def _alloc(self, B):
    bufs = {}
    self._alloc_controller(B, bufs)
    dev, T = self.device, self.T
    bufs['iou_box'] = torch.empty((B, T, T), device=dev, dtype=torch.float32)
    bufs['grd'] = torch.empty((B, T), device=dev, dtype=torch.float32)
    return bufs

  def _run(self, bufs, B, noise, train=False, tape=False):
    self._tape_on = bufs['tape'] if tape else None
    try:
      return self._run_inner(bufs, B, noise, train, tape)
    finally:
      self._tape_on = None

  def _run_inner(self, bufs, B, noise, train, tape):
    w, o = self.w, self.opt
    T, H, W = self.T, self.H, self.W
    st = bufs['static_in']
    y_gt, s_gt = st['y_gt'], st['s_gt']
    self._prepare(bufs, st['x'], st.get('d_in'), st.get('y_in'))
    _lib.TAG = 'loss'
    tl, br, box_gt, rect, area = ops.get_gt_box(y_gt, padding_ratio=o['attn_box_padding_ratio'],
                                                min_padding=self.min_padding, want_box=False)
    thw = T * H * W
    for t in range(T):
      self._controller(bufs, t, train)
      _lib.TAG = 'paste_back'
      fy, fx = self._filt(bufs, t)
      ops.paste_back(None, bufs['box_all'][t], fy, fx, None, attn_box=bufs['attn_box'][:, t],
                     y_out=None, out_bstride=thw, band=bufs['band'])
      _lib.TAG = 'box_gt'
      noise_t = None if noise is None else noise[:, t]
      if o.get('use_iou_box', False):  # box_model.py:487-491: coordinate IoU instead of the soft box IoU
        ops.greedy_iou_box(bufs['box_all'][t], tl, br, bufs['iou_box'][:, t], T * T, bufs['grd'])
        ops.box_gt_canvas(bufs['grd'], y_gt, noise_t, thw, bufs['canvas'])
      else:
        ops.box_gt_step(bufs['attn_box'][:, t], thw, rect, y_gt, noise_t, thw, bufs['iou_box'][:, t], T * T,
                        bufs['grd'], bufs['canvas'])
      ops.score(bufs['h_all'][t], None, w['score_mlp_w_0'], w['score_mlp_b_0'], bufs['s_out'][:, t], T)
    out = {}
    self._controller_outputs(bufs, out)
    _lib.TAG = 'loss'
    match_box = ops.f_segm_match(bufs['iou_box'], s_gt)
    scal = ops.loss_block(bufs['iou_box'], match_box, bufs['iou_box'], match_box, None, None, bufs['s_out'], s_gt,
                          area, 1.0, 0.0, segm_coeff=0.0)
    self._add_wd(scal)
    out.update({'iou_soft_box_pairwise': bufs['iou_box'], 'match_box': match_box, 'attn_top_left_gt': tl,
                'attn_bot_right_gt': br, '_gt_rect': rect, 'box_loss': scal[0], 'conf_loss': scal[2],
                'loss': scal[13]})
    if tape:
      from . import train as TR
      gs = bufs['grads']
      TR.box_model_backward(self, gs, bufs, B, out)
      self._trainer.scatter(B, gs.entries)
    return out

  def train_step(self, batch, frozen=(), use_graph=True, grad_scale=None):
    """``sess.run([loss, train_step])`` (runner.py:98-105) for the graph of box_model.py:635-652: training-mode
    forward, backward (box loss on the per-step IoUs + confidence loss), gradient all-reduce, clip + Adam, device-side
    refresh of the weight images.  batch may carry the explicit canvas-noise draw ``canvas_noise``."""
    if self.w is None:
      raise _lib.RecAttendError('load_weights() first')
    if self._trainer is None:
      from . import train as TR
      self._trainer = TR.Trainer(self, frozen=frozen)
    out = self.forward(batch, phase_train=True, use_graph=use_graph, _tape=True)
    lr = self._trainer.apply(grad_scale=grad_scale)
    return {'loss': out['loss'], 'box_loss': out['box_loss'], 'conf_loss': out['conf_loss'], 'learn_rate': lr,
            'global_step': self._trainer.optim.global_step}

  @property
  def optimizer(self):
    return None if self._trainer is None else self._trainer.optim

  def forward(self, batch, outputs=None, phase_train=False, use_graph=True, _tape=False):
    """``sess.run`` replacement for the box model.  batch: x, y_gt, s_gt[, d_in, y_in] and the
    optional explicit random draw ``canvas_noise`` [B,T,H,W] of box_model.py:501-502 (zeros when absent).
    phase_train=True: every BN layer of the controller CNN normalises with the statistics of this batch and moves
    the EMA shadows of its (layer, step) copy (nnlib.py:96-119).
    With ``use_graph`` the launches of the T-step loop are captured once per (batch size, noise present, mode) in a
    CUDA graph and replayed; the returned tensors are then the graph's static outputs."""
    if self.w is None:
      raise _lib.RecAttendError('load_weights() first')
    train = bool(phase_train)
    if train:
      self._bn_dirty = True
    elif self._bn_dirty:
      self._refold_bn()
    x, d_in, y_in, y_gt, s_gt = self._inputs(batch)
    if y_gt is None or s_gt is None:
      raise _lib.RecAttendError('the box model needs y_gt / s_gt: its canvas is driven by the ground truth')
    B = x.shape[0]
    bufs = self._buffers(B)
    cur = torch.cuda.current_stream()
    st = self._stage_inputs(bufs, 0, {'x': x, 'd_in': d_in, 'y_in': y_in, 'y_gt': y_gt, 's_gt': s_gt}, cur)
    bufs['static_in'] = st
    noise = None
    if batch.get('canvas_noise') is not None:
      src = torch.as_tensor(batch['canvas_noise'], dtype=torch.float32)
      if 'static_noise' not in bufs:
        bufs['static_noise'] = torch.empty(tuple(src.shape), device=self.device, dtype=torch.float32)
      bufs['static_noise'].copy_(src, non_blocking=True)
      noise = bufs['static_noise']
    if _tape and 'tape' not in bufs:
      from . import train as TR
      bufs['tape'] = TR.alloc_tape(self, B, full=False)
      bufs['grads'] = TR._alloc_grads(self, full=False)
    if not use_graph:
      out = self._run(bufs, B, noise, train, _tape)
    else:
      graphs = bufs.setdefault('graphs', {})
      key = (noise is not None, train, bool(_tape))
      if key not in graphs:
        # warm-up on a side stream (lazy weight packing, attribute calls, workspace allocation), then capture
        side = torch.cuda.Stream()
        ema_keep = None
        if train:  # the warm-up and the capture run must not move the EMA shadows: only replays count
          ema_keep = {k: v.clone() for k, v in self.w.items() if '_ema_' in k and isinstance(v, torch.Tensor)}
        side.wait_stream(cur)
        with torch.cuda.stream(side):
          self._run(bufs, B, noise, train, _tape)
        cur.wait_stream(side)
        g, static_out = capture_graph(lambda: self._run(bufs, B, noise, train, _tape))
        if ema_keep is not None:
          for k, v in ema_keep.items():
            self.w[k].copy_(v)
        graphs[key] = (g, static_out)
      g, out = graphs[key]
      g.replay()
    if outputs is not None:
      out = {k: out[k] for k in outputs}
    return out


def get_model(opt, device=None):
  """Same call as the reference's ``box_model.get_model(opt)``."""
  return BoxModel(opt, device=device)
