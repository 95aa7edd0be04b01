"""Index bookkeeping between the reference's weight dict (weights.h5 key schema, full_model_read.py:32-70) and the
device-side images the kernels consume.

The kernels never see the reference layouts directly: filters are channel-permuted (the canvas channel is split off the
first controller layer), zero-padded (glimpse channels to a multiple of 4), flipped / transposed (conv2d_transpose as a
convolution), stacked (LSTM gates) and, for the tcgen05 convolution, packed into hi / lo tf32 shared-memory images.
Every one of these is a pure PERMUTATION with zero padding, so each device tensor is described by an integer array of the
same shape whose entries name the source element (0 = constant zero, k > 0 = element k - 1 of the flat vector of ALL
weights in sorted key order).  The same arrays drive

  * `ra_param_gather_f32`: trainable flat bucket -> all device images, one launch, after `apply_gradients`
    (full_model.py:1056) - no host round trip, no re-packing on the host;
  * `ra_param_scatter_f32`: per-tensor gradients in device layout -> the flat gradient bucket that is all-reduced.

Pure numpy: importable (and tested) without a GPU.
"""
import numpy as np

KIND_VALUE, KIND_HI, KIND_LO = 0, 1, 2


class AllLayout(object):
  """key -> (offset, shape) over every tensor of the weight dict, sorted by key."""

  def __init__(self, weights):
    self.keys = sorted(weights)
    self.layout = {}
    off = 0
    for k in self.keys:
      shape = tuple(np.asarray(weights[k]).shape)
      n = int(np.prod(shape)) if shape else 1
      self.layout[k] = (off, shape)
      off += n
    self.numel = off

  def index(self, key):
    """int64 array shaped like weights[key]: 1 + the position of every element in the all-weights vector."""
    off, shape = self.layout[key]
    n = int(np.prod(shape)) if shape else 1
    return (np.arange(off + 1, off + 1 + n, dtype=np.int64)).reshape(shape)


class WI(object):
  """A weight array travelling together with the index array that says where each element came from."""

  def __init__(self, val, idx, kind=None):
    self.val = np.asarray(val, np.float32)
    self.idx = np.asarray(idx, np.int64)
    self.kind = None if kind is None else np.asarray(kind, np.int64)
    assert self.val.shape == self.idx.shape, (self.val.shape, self.idx.shape)

  def map(self, f):
    """Apply a pure indexing / transposing / zero-padding function to values and indices alike."""
    return WI(f(self.val), f(self.idx))

  @staticmethod
  def stack(items, axis=0):
    return WI(np.stack([i.val for i in items], axis), np.stack([i.idx for i in items], axis))


def wi_of(weights, layout, key):
  return WI(np.asarray(weights[key], np.float32), layout.index(key))


def pad_cin(a, cin):
  """Zero rows for padded input channels of an HWIO filter (appended after the real ones); dtype-preserving."""
  if a.shape[2] == cin:
    return a
  out = np.zeros(a.shape[:2] + (cin, a.shape[3]), a.dtype)
  out[:, :, :a.shape[2]] = a
  return out


def deconv_to_conv(a):
  """conv2d_transpose filter [kh,kw,Cout,Cin] (nnlib.py:320-325,372-376) -> the HWIO filter of the equivalent
  stride-1 convolution over the zero-inserted input: spatial flip + swap of the channel axes."""
  return np.ascontiguousarray(a[::-1, ::-1].transpose(0, 1, 3, 2))


def flip_transpose(a):
  """HWIO conv filter -> the filter of its data-gradient convolution: out[ky,kx,co,ci] = w[2-ky,2-kx,ci,co]."""
  return np.ascontiguousarray(a[::-1, ::-1].transpose(0, 1, 3, 2))


def umma_layout(a, KC, NPc, n_split, fill=0):
  """[3,3,Cin,Cout] -> [n_split, n_chunks, 9, KC/4, NPc, 4] (zero / `fill` padded): the operand order of one half
  (hi or lo) of the tcgen05 kernel's shared-memory filter image (csrc/conv_umma.cu)."""
  _, _, Cin, Cout = a.shape
  n_chunks = (Cin + KC - 1) // KC
  NP = NPc * n_split
  wp = np.full((9, n_chunks * KC, NP), fill, a.dtype)
  wp[:, :Cin, :Cout] = a.reshape(9, Cin, Cout)
  return wp.reshape(9, n_chunks, KC // 4, 4, n_split, NPc).transpose(4, 1, 0, 2, 5, 3)


def tf32_hi(a):
  """Round fp32 to the nearest tf32 (10-bit mantissa), ties away from zero in magnitude like the device kernel."""
  a = np.ascontiguousarray(a, np.float32)
  return ((a.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def pack_umma(wi, KC, NPc, n_split, rowstack=0):
  """WI of an HWIO filter -> WI of the packed image [n_split][n_chunks][9][KC/4][2*NPc][4] with per-element kinds:
  rows 0..NPc-1 = hi (w rounded to the nearest tf32), rows NPc..2NPc-1 = lo = w - hi (exact).
  rowstack (narrow layers, csrc/conv_umma.cu issue_chunk_rs): [n_split][n_chunks][3 ky][KC/4][6*NPc][4] with rows
  [hi kx0 | hi kx1 | hi kx2 | lo kx0 | lo kx1 | lo kx2]."""
  v = umma_layout(wi.val, KC, NPc, n_split)
  i = umma_layout(wi.idx, KC, NPc, n_split)
  if rowstack:
    def stack(a):  # [ns, nch, 9 = (ky, kx), pl, NPc, 4] -> [ns, nch, ky, pl, kx * NPc, 4]
      ns, nch, _, pl, npc, four = a.shape
      return a.reshape(ns, nch, 3, 3, pl, npc, four).transpose(0, 1, 2, 4, 3, 5, 6).reshape(ns, nch, 3, pl, 3 * npc, four)
    v, i = stack(v), stack(i)
  hi = tf32_hi(v)
  lo = (v - hi).astype(np.float32)
  val = np.ascontiguousarray(np.concatenate([hi, lo], axis=4))
  idx = np.ascontiguousarray(np.concatenate([i, i], axis=4))
  kind = np.ascontiguousarray(np.concatenate([np.full(i.shape, KIND_HI, np.int64), np.full(i.shape, KIND_LO, np.int64)],
                                             axis=4))
  return WI(val, idx, kind)


def umma_f16_source(wi, KC, NPc, n_split):
  """WI of an HWIO filter -> WI of the fp32 SOURCE of the fp16 hi / lo filter image (plans whose layout flags have
  bit 1 set, RA_UMMA_F16): [n_split][n_chunks][9][KC/4][NPc][4], plain values (no kinds).  The image itself is made from
  it on the device by ra_umma_pack_f16 (ops.umma_pack_f16), at load time and after every optimiser step."""
  return WI(np.ascontiguousarray(umma_layout(wi.val, KC, NPc, n_split)),
            np.ascontiguousarray(umma_layout(wi.idx, KC, NPc, n_split)))


def pack_umma_f16_reference(v):
  """numpy model of ra_umma_pack_f16 (tests): v [..., KC/4, NPc, 4] fp32 -> [..., KC/8, 2 NPc, 8] float16."""
  v = np.asarray(v, np.float32)
  lead, (pl, npc, _) = v.shape[:-3], v.shape[-3:]
  x = v.reshape(lead + (pl // 2, 2, npc, 4)).swapaxes(-3, -2).reshape(lead + (pl // 2, npc, 8))
  hi = x.astype(np.float16)
  lo = ((x - hi.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
  return np.concatenate([hi, lo], axis=-2)


def to_train_map(all_layout, flat_params):
  """Vector m of length numel+1: m[0] = 0 (the zero constant), m[1 + all position] = 1 + position in the TRAINABLE
  flat bucket (`optim.FlatParams`), or -1 for tensors that are not trained (EMA shadows, frozen nets)."""
  m = np.full(all_layout.numel + 1, -1, np.int64)
  m[0] = 0
  for k, (off_t, shape) in flat_params.layout.items():
    off_a, shape_a = all_layout.layout[k]
    assert tuple(shape) == tuple(shape_a), k
    n = int(np.prod(shape)) if shape else 1
    m[1 + off_a:1 + off_a + n] = np.arange(off_t + 1, off_t + 1 + n, dtype=np.int64)
  return m


def encode(idx, kind, tmap):
  """Codes of `ra_param_gather_f32` / `ra_param_scatter_f32` for one device tensor, or None when the tensor draws on
  a non-trainable source (it never changes: leave it alone)."""
  t = tmap[np.asarray(idx, np.int64).reshape(-1)]
  if (t < 0).any():
    return None
  if t.max(initial=0) >= (1 << 29):
    raise ValueError('parameter bucket too large for 32-bit codes')
  k = np.zeros_like(t) if kind is None else np.asarray(kind, np.int64).reshape(-1)
  code = (t << 2) | k
  code[t == 0] = 0
  return code.astype(np.int32)


def gather_reference(flat, code):
  """numpy model of ra_param_gather_f32 (tests)."""
  flat = np.asarray(flat, np.float32)
  code = np.asarray(code, np.int64)
  src = flat[np.maximum((code >> 2) - 1, 0)]
  hi = tf32_hi(src)
  kind = code & 3
  out = np.where(kind == KIND_VALUE, src, np.where(kind == KIND_HI, hi, (src - hi).astype(np.float32)))
  out[code == 0] = 0.0
  return out.astype(np.float32)
