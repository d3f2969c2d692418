"""Kernel wrappers and autograd Functions over the C-ABI (include/sg2im_b200.h).

Internal tensor convention: activations are 4-D tensors of SHAPE (N, H, W, C)
("NHWC"), fp32, normally contiguous; matrices are (rows, features).  The
nn.Module mirror (layers.py etc.) presents the reference's NCHW shapes to
callers by permuting views — no data movement.

PyTorch here is plumbing only: allocation (caching allocator), stream choice,
autograd graph bookkeeping.  Every arithmetic op of the hot path is a launch
into libsg2im_b200.so; a non-CUDA tensor raises (no CPU fallback).
"""
import os

import torch

from . import _lib

_call = _lib.call


def _stream():
  return torch.cuda.current_stream().cuda_stream


def _p(t):
  return None if t is None else t.data_ptr()


def _chk(t, dtype=torch.float32, name='tensor'):
  if not t.is_cuda:
    raise RuntimeError('sg2im_b200: %s must be a CUDA tensor (the hot path has no CPU '
                       'fallback; the CPU oracle lives in oracle/ for tests only)' % name)
  if t.dtype != dtype:
    raise RuntimeError('sg2im_b200: %s must be %s, got %s' % (name, dtype, t.dtype))
  return t


def _count(n=1):
  _lib.launches += n


# bench.py sets PROFILE to a list to time the conv kernels with CUDA events on
# the launching stream: entries (kernel family, algorithmic flops, start, end)
PROFILE = None
# arithmetic of the convolution / Linear path (set_conv_math): 'bf16x3' (default) = tcgen05 tensor cores
# on in-kernel bf16 hi / mid operand pairs, three products per fp32 multiply — fp32-grade results
# (1e-3 parity bar met with ~2 orders of magnitude to spare); 'tf32' / 'bf16' = faster, less exact
# tensor-core modes; 'fp32' = the exact FFMA kernels
CONV_MATH = 'bf16x3'


class _prof(object):
  def __init__(self, name, flops, shape=None):
    self.name, self.flops, self.shape = name, flops, shape

  def __enter__(self):
    if PROFILE is not None:
      self.a = torch.cuda.Event(enable_timing=True)
      self.b = torch.cuda.Event(enable_timing=True)
      self.a.record()

  def __exit__(self, *exc):
    if PROFILE is not None:
      self.b.record()
      PROFILE.append((self.name, self.flops, self.a, self.b, self.shape))


# bench.py sets PROFILE_HBM to a list to time the HBM-bound kernels (graph gather / pooling,
# layout warp, crops, normalise / activate passes, layout conversions) the same way: entries
# (entry point, ALGORITHMIC bytes of the launch — every operand read or written once —, start, end)
PROFILE_HBM = None


def _event():
  return torch.cuda.Event(enable_timing=True)


def _call_b(nbytes, name, *args):
  """_call, bracketed by CUDA events on the launching stream while bench.py profiles."""
  if PROFILE_HBM is None:
    return _call(name, *args)
  a, b = _event(), _event()
  a.record()
  _call(name, *args)
  b.record()
  PROFILE_HBM.append((name, float(nbytes), a, b))


# --------------------------------------------------------------------------
# raw kernel wrappers (no autograd)
# --------------------------------------------------------------------------

class ZeroArena(object):
  """Small zero-initialised scratch buffers (BatchNorm statistics, per-channel sums, tiny gradient
  accumulators) carved out of ONE persistent buffer that is cleared by ONE memset at the start of a
  training step, instead of one torch.zeros fill kernel per buffer (~85 launches per step).  Only
  active inside TrainStep.step (ops.ZERO_ARENA): buffers are valid until the next step begins, and
  nothing handed out is kept across steps (parameter gradients live in the flat buckets)."""

  def __init__(self, nbytes, device):
    self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    self.off = 0
    self.high = 0

  def reset(self):
    self.high = max(self.off, self.high)            # monotone: a captured memset covers every later step
    if self.high:
      self.buf[:self.high].zero_()
    self.off = 0

  def take(self, shape, dtype):
    n = 1
    for d in shape:
      n *= int(d)
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    start = (self.off + 15) // 16 * 16
    if start + nbytes > self.buf.numel():
      return None
    self.off = start + nbytes
    self.high = max(self.high, self.off)
    return self.buf[start:start + nbytes].view(dtype).view(shape)


ZERO_ARENA = None           # set by TrainStep.step
# Data parallel: called (no arguments) when the backward pass of the cascaded-refinement network has
# finished, i.e. when the layout's backward starts — 95 % of the generator's gradient bytes are then
# final, and TrainStep starts their all-reduce while the layout / mask-head / graph-convolution
# backward still runs.
GRAD_READY_HOOK = None


def _zeros(shape, dtype, device):
  """torch.zeros, from the step's zero arena when one is active and the buffer is small."""
  shape = tuple(int(d) for d in (shape if isinstance(shape, (tuple, list)) else (shape,)))
  if ZERO_ARENA is not None:
    n = 1
    for d in shape:
      n *= d
    if n <= (1 << 18):
      t = ZERO_ARENA.take(shape, dtype)
      if t is not None:
        return t
  return torch.zeros(shape, dtype=dtype, device=device)


def csr_build(idx, nroles, num_rows):
  """idx: int64 (T,) for nroles=1 or (T,2) for nroles=2.  Returns
  (row_ptr int32[R+1], entries int32[nroles*T])."""
  _chk(idx, torch.int64, 'idx')
  idx = idx.contiguous()
  T = idx.size(0)
  stride = 1 if idx.dim() == 1 else idx.size(1)
  row_ptr = torch.empty(num_rows + 1, dtype=torch.int32, device=idx.device)
  entries = torch.empty(max(nroles * T, 1), dtype=torch.int32, device=idx.device)
  _call('sg2im_csr_build', _p(idx), T, stride, nroles, num_rows, _p(row_ptr), _p(entries),
        _stream())
  _count(3)
  return row_ptr, entries


def triple_gather(rows, mid, edges, Wm, row_ptr=None):
  rows = _chk(rows).contiguous()
  T, Wr = edges.size(0), rows.size(1)
  if mid is not None:
    mid = _chk(mid).contiguous()
  out = torch.empty(T, 2 * Wr + Wm, dtype=torch.float32, device=rows.device)
  _call_b(8 * T * (2 * Wr + Wm) + 16 * T,
          'sg2im_triple_gather', _p(rows), _p(mid), _p(edges), T, Wr, Wm, _p(row_ptr), _p(out),
          _stream())
  _count()
  return out


def segment_sum(src, off0, off1, W, row_ptr, entries, num_rows, avg):
  src = _chk(src).contiguous()
  out = torch.empty(num_rows, W, dtype=torch.float32, device=src.device)
  _call_b(4 * W * (entries.numel() + num_rows) + 4 * entries.numel(),
          'sg2im_segment_sum', _p(src), src.size(1), off0, off1, W, _p(row_ptr), _p(entries),
          num_rows, int(avg), _p(out), _stream())
  _count()
  return out


def conv_out_size(h, k, s, p):
  return (h + 2 * p - k) // s + 1


def conv_igemm(mode, x, w_packed, bias, KH, KW, S, P, out_hw, Cout, act=0, slope=0.0,
               out=None, out_coff=0):
  """x: (N,Hin,Win,Cin) any strides.  Returns/writes (N,Hout,Wout,Cout[slice])."""
  _chk(x)
  N, Hin, Win, Cin = x.shape
  Hout, Wout = out_hw
  sn, sh, sw, sc = x.stride()
  if out is None:
    out = torch.empty(N, Hout, Wout, Cout, dtype=torch.float32, device=x.device)
  cstride = out.size(3)
  pix = N * Hout * Wout if mode == 0 else N * Hin * Win
  with _prof('conv_fwd' if mode == 0 else 'conv_dgrad', 2.0 * pix * Cin * Cout * KH * KW,
             (N, Hin, Win, Cin, Cout, KH, S)):
    _call('sg2im_conv_igemm', mode, _p(x), sn, sh, sw, sc, N, Hin, Win, Cin, _p(w_packed),
          _p(bias), KH, KW, S, P, Hout, Wout, Cout, int(act), float(slope), _p(out), cstride,
          out_coff, _stream())
  _count()
  return out


TC_MODES = ('tf32', 'bf16x3', 'bf16')
_MATH_ID = {'tf32': 0, 'bf16x3': 1, 'bf16': 2}     # SG2IM_MATH_* of include/sg2im_b200.h


def set_conv_math(mode):
  """'fp32': every convolution / Linear on the exact-fp32 FFMA kernels.
  The other modes run stride-1 convolutions and Linears whose shape tiles on the
  tcgen05 tensor-core kernels (fp32 accumulate in TMEM; activations and weights stay
  fp32 in HBM in all of them, only what the tensor core multiplies differs):
  'bf16x3': converter warps split every shared-memory operand tile into bf16 hi / mid
  halves and each fp32 product is issued as hi*hi + mid*hi + hi*mid (kind::f16) — 2^-17
  operand precision, the mode that meets the 1e-3 parity bar against the fp32 reference
  (scripts/train.py:423) on the tensor core, and what bench.py measures by default.
  'tf32': kind::tf32 on the fp32 words (2^-11 operand precision; producers hand over
  round-to-nearest TF32 values) — 2/3 of the tensor-pipe time of 'bf16x3', ~1e-2 off the
  reference end to end at benchmark size (tools/tf32_attribution.py): the labelled fast line.
  'bf16': the bf16x3 kernels issuing hi*hi only (plain bf16 operands, BASELINE.json configs[3])."""
  global CONV_MATH
  if mode not in ('fp32',) + TC_MODES:
    raise ValueError("conv math must be one of 'fp32', 'tf32', 'bf16x3', 'bf16'")
  CONV_MATH = mode


def _tc_math():
  return CONV_MATH in TC_MODES


def _math_id():
  return _MATH_ID[CONV_MATH]


def _tc_shape_ok(N, H, W, C, KH, KW, P, Cout, out_hw):
  """conv_tc_ok for a contiguous, 16-byte aligned NHWC tensor that does not exist yet."""
  return bool(_lib.load().sg2im_conv_tc_supported(N, H, W, C, C, KH, KW, 1, P, out_hw[0],
                                                  out_hw[1], Cout, Cout, 0))


def _pixel_stride(x):
  """Pixel stride (floats) if x is an NHWC channel-prefix view of a dense
  buffer, else None."""
  N, H, W, C = x.shape
  if x.is_contiguous():
    return C
  if x.stride(3) != 1:
    return None
  cs = x.stride(2) if W > 1 else (x.stride(1) if H > 1 else x.stride(0))
  ok = ((W == 1 or x.stride(2) == cs) and (H == 1 or x.stride(1) == W * cs)
        and (N == 1 or x.stride(0) == H * W * cs) and cs >= C)
  return cs if ok else None


def conv_tc_ok(x, KH, KW, S, P, Cout, out_hw=None, y_cstride=None, y_coff=0):
  if not _tc_math() or S != 1:
    return False
  cs = _pixel_stride(x)
  if cs is None or x.data_ptr() % 16:
    return False
  N, H, W, C = x.shape
  Hout, Wout = out_hw if out_hw is not None else (H + 2 * P - KH + 1, W + 2 * P - KW + 1)
  return bool(_lib.load().sg2im_conv_tc_supported(
      N, H, W, C, cs, KH, KW, S, P, Hout, Wout, Cout, Cout if y_cstride is None else y_cstride,
      y_coff))


def conv_tc(x, w_tc, bias, KH, KW, P, Cout, act=0, slope=0.0, out=None, out_coff=0,
            tag='conv_fwd_tc', out_hw=None, stats=None, round_out=False):
  """Tensor-core stride-1 convolution; x NHWC (channel-prefix view allowed),
  w_tc packed [KH*KW][Cout][Cin]; out_hw: explicit output size (reads outside
  the input are zero)."""
  N, H, W, C = x.shape
  cs = _pixel_stride(x)
  Hout, Wout = out_hw if out_hw is not None else (H + 2 * P - KH + 1, W + 2 * P - KW + 1)
  if out is None:
    out = torch.empty(N, Hout, Wout, Cout, dtype=torch.float32, device=x.device)
  with _prof(tag, 2.0 * N * Hout * Wout * C * Cout * KH * KW, (N, H, W, C, Cout, KH, 1)):
    _call('sg2im_conv_tc', _p(x), cs, N, H, W, C, _p(w_tc), _p(bias), KH, KW, P, Hout, Wout,
          Cout, int(act), float(slope), _p(out), out.size(3), out_coff, _p(stats),
          int(round_out), _math_id(), _stream())
  _count()
  return out


def _pack(weight, cin_use, want_fwd, want_dgrad=None):
  """One launch of sg2im_pack_weights; returns the requested layout, or the pair
  (fwd, dgrad) when both are requested (weights read once)."""
  if want_dgrad is None:
    want_dgrad = not want_fwd
  weight = weight.contiguous()
  Co, Ci, KH, KW = weight.shape
  T = KH * KW
  cu = Ci if cin_use is None else cin_use
  dev = weight.device
  f = torch.empty((T, Co, cu), dtype=torch.float32, device=dev) if want_fwd else None
  d = torch.empty((T, cu, Co), dtype=torch.float32, device=dev) if want_dgrad else None
  _call_b(4 * T * Co * cu * (1 + int(want_fwd) + int(want_dgrad)),
          'sg2im_pack_weights', _p(weight), Co, Ci, cu, T, _p(f), _p(d), int(CONV_MATH == 'tf32'),
          _stream())                                        # tf32: RN-TF32 (the tensor core would truncate)
  _count()
  if want_fwd and want_dgrad:
    return f, d
  return f if want_fwd else d


def pack_tc_fwd(weight, cin_use=None):
  """OIHW (first cin_use input channels) -> [KH*KW][Cout][Cin]."""
  return _pack(weight, cin_use, True)


def pack_tc_dgrad(weight, cin_use=None):
  """OIHW -> [KH*KW (flipped)][Cin][Cout]: the data gradient of a stride-1 conv
  is a conv of dY with the spatially flipped, channel-transposed filter."""
  return _pack(weight, cin_use, False)


def unpack_wgrad_oihw(dw, wshape, cin_use):
  """dw [T][cin_use][Cout] -> OIHW gradient of the full weight (channels beyond
  cin_use get zero)."""
  Co, Ci, KH, KW = wshape
  grad = (torch.zeros if cin_use != Ci else torch.empty)(wshape, dtype=torch.float32, device=dw.device)
  _call_b(4 * (dw.numel() + grad.numel()),
          'sg2im_unpack_wgrad', _p(dw), Co, Ci, cin_use, KH * KW, _p(grad), 0, _stream())
  _count()
  return grad


def conv_wgrad(x, dy, KH, KW, S, P, accumulate_into=None, s2d_c=0):
  """Returns dw packed (KH*KW*Cin, Cout).  accumulate_into: a contiguous buffer of that size
  the kernels ADD into (they combine partial tiles with atomics anyway) instead of a fresh
  zeroed one — e.g. the parameter's slice of the flat gradient bucket.
  s2d_c = C > 0: x is the space-to-depth form of a 4x4 stride-2 convolution's input (2x2 taps,
  4C channels) and the rows are written in THAT filter's (16, C, Cout) order (tensor-core
  kernel only)."""
  _chk(x)
  dy = _chk(dy).contiguous()
  N, Hin, Win, Cin = x.shape
  _, Hout, Wout, Cout = dy.shape
  sn, sh, sw, sc = x.stride()
  if accumulate_into is not None:
    assert accumulate_into.is_contiguous() and accumulate_into.numel() == KH * KW * Cin * Cout
    dw = accumulate_into.view(KH * KW * Cin, Cout)
  else:
    dw = torch.zeros(KH * KW * Cin, Cout, dtype=torch.float32, device=x.device)
  if _tc_math() and S == 1:
    cs = _pixel_stride(x)
    if (cs is not None and x.data_ptr() % 16 == 0 and _lib.load().sg2im_conv_wgrad_tc_supported(
        N, Hin, Win, Cin, cs, KH, KW, S, P, Hout, Wout, Cout)):
      with _prof('conv_wgrad_tc', 2.0 * N * Hout * Wout * Cin * Cout * KH * KW,
                 (N, Hin, Win, Cin, Cout, KH, S)):
        _call('sg2im_conv_wgrad_tc', _p(x), cs, N, Hin, Win, Cin, _p(dy), KH, KW, P, Hout, Wout,
              Cout, _p(dw), _math_id(), int(s2d_c), _stream())
      _count()
      return dw
  if s2d_c:
    raise RuntimeError('sg2im_b200: the re-tiled 4x4 stride-2 weight gradient needs the tensor-core kernel')
  with _prof('conv_wgrad', 2.0 * N * Hout * Wout * Cin * Cout * KH * KW,
             (N, Hin, Win, Cin, Cout, KH, S)):
    _call('sg2im_conv_wgrad', _p(x), sn, sh, sw, sc, N, Hin, Win, Cin, _p(dy), KH, KW, S, P,
          Hout, Wout, Cout, _p(dw), _stream())
  _count()
  return dw


def colsum(x2d):
  x2d = _chk(x2d).contiguous()
  M, C = x2d.shape
  out = torch.empty(C, dtype=torch.float32, device=x2d.device)
  scratch = torch.empty(C, dtype=torch.float64, device=x2d.device)
  _call('sg2im_colsum', _p(x2d), M, C, _p(out), _p(scratch), _stream())
  _count(3)
  return out


def act_bwd(dy, y, slope):
  dy = _chk(dy).contiguous()
  dx = torch.empty_like(dy)
  _call_b(12 * dy.numel(),
          'sg2im_act_bwd', _p(dy), _p(y), float(slope), dy.numel(), _p(dx), _stream())
  _count()
  return dx


# One pass for the activation backward and the bias gradient of a conv+bias+LeakyReLU epilogue
# (instead of act_bwd + the three-launch colsum), accumulating straight into the bias' slot of the
# flat gradient bucket inside TrainStep (DIRECT_WGRAD).  Default since round 2 (validated and timed
# on the B200); SG2IM_ACTBWD_FUSED=0 restores the separate passes.
FUSE_ACT_BWD = os.environ.get('SG2IM_ACTBWD_FUSED') != '0'


def act_bwd_bias(dy, y, slope, bias):
  """Returns (dx, db) like act_bwd + colsum.  db is None when it was accumulated directly into
  bias.grad (the caller then returns no bias gradient to autograd)."""
  C = dy.size(-1)
  if not (FUSE_ACT_BWD and C % 4 == 0 and dy.numel() < (1 << 31)):
    dx = act_bwd(dy, y, slope)
    return dx, colsum(dx.view(-1, C))
  dy = _chk(dy).contiguous()
  dx = torch.empty_like(dy)
  g = bias.grad if bias is not None else None
  direct = (DIRECT_WGRAD and g is not None and g.is_contiguous() and g.numel() == C
            and g.data_ptr() % 4 == 0)
  db = g if direct else _zeros(C, torch.float32, dy.device)
  _call_b(12 * dy.numel() + 4 * C,
          'sg2im_act_bwd_colsum', _p(dy), _p(y), float(slope), dy.numel() // C, C, _p(dx), _p(db),
          _stream())
  _count()
  return dx, (None if direct else db)


def bn_scale_shift(x, gamma, beta, running_mean, running_var, training, momentum, eps,
                   unbias_mult=1, sums=None, num_batches_tracked=None):
  """Batch statistics of x (rows = all dims but the last) -> (scale, shift, save)."""
  C = x.size(-1)
  M = x.numel() // C
  dev = x.device
  scale = torch.empty(C, dtype=torch.float32, device=dev)
  shift = torch.empty(C, dtype=torch.float32, device=dev)
  save = torch.empty(2 * C, dtype=torch.float32, device=dev)
  if training and sums is None:
    sums = _zeros(2 * C, torch.float64, dev)
    _call_b(4 * M * C, 'sg2im_bn_stats', _p(x), M, C, _p(sums), _stream())
    _count()
  _call('sg2im_bn_finalize', _p(sums), M, unbias_mult, C, _p(gamma), _p(beta), float(eps),
        float(momentum), int(training), _p(running_mean), _p(running_var), _p(scale), _p(shift),
        _p(save), _p(num_batches_tracked), _stream())
  _count()
  return scale, shift, save


def scale_act_fwd(x, scale, shift, slope, up, out=None, out_coff=0):
  N, H, W, C = x.shape
  if out is None:
    out = torch.empty(N, H * up, W * up, C, dtype=torch.float32, device=x.device)
  _call_b(4 * x.numel() * (1 + up * up),
          'sg2im_scale_act_fwd', _p(x), N, H, W, C, _p(scale), _p(shift), float(slope), up,
          _p(out), out.size(3), out_coff, int(CONV_MATH == 'tf32'), _stream())
  _count()
  return out


def scale_act_bwd(dy, dy_coff, x, scale, shift, save, slope, up, training, want_param_grads,
                  grad_into=None):
  """dy: (N,H*up,W*up,Ctot) contiguous, slice [dy_coff, dy_coff+C).  Returns
  (dx, dgamma, dbeta).  grad_into = (gamma.grad, beta.grad): ADD the parameter gradients straight
  into those slots of the flat gradient bucket (returns None for them)."""
  N, H, W, C = x.shape
  dev = x.device
  sums = None
  need_sums = (training and save is not None) or want_param_grads
  if need_sums:
    sums = _zeros(2 * C, torch.float64, dev)
    _call_b(4 * x.numel() * (1 + up * up),
            'sg2im_scale_act_bwd_reduce', _p(dy), dy.size(3), dy_coff, _p(x), N, H, W, C,
            _p(scale), _p(shift), _p(save), float(slope), up, _p(sums), _stream())
    _count()
  dx = torch.empty_like(x)
  dgamma = dbeta = None
  flag = int(training)
  if want_param_grads and grad_into is not None:
    dgamma, dbeta = grad_into
    flag |= 2                                       # accumulate into the bucket slots
  elif want_param_grads:
    dgamma = torch.empty(C, dtype=torch.float32, device=dev)
    dbeta = torch.empty(C, dtype=torch.float32, device=dev)
  _call_b(4 * x.numel() * (2 + up * up),
          'sg2im_scale_act_bwd_apply', _p(dy), dy.size(3), dy_coff, _p(x), N, H, W, C, _p(scale),
          _p(shift), _p(save), float(slope), up, flag, _p(sums), _p(dx), _p(dgamma),
          _p(dbeta), _stream())
  _count(2 if want_param_grads else 1)
  if grad_into is not None and want_param_grads:
    return dx, None, None
  return dx, dgamma, dbeta


def avgpool2_fwd(x, x_coff, C, out, out_coff):
  N, H, W, _ = x.shape
  _call_b(5 * N * H * W * C,
          'sg2im_avgpool2_fwd', _p(x), x.size(3), x_coff, N, H, W, C, _p(out), out.size(3),
          out_coff, _stream())
  _count()


def avgpool2_bwd(dcoarse, dc_coff, C, dfine, df_coff, accumulate):
  N, H, W, _ = dfine.shape
  _call_b((9 if accumulate else 5) * N * H * W * C,
          'sg2im_avgpool2_bwd', _p(dcoarse), dcoarse.size(3), dc_coff, N, H, W, C, _p(dfine),
          dfine.size(3), df_coff, int(accumulate), _stream())
  _count()


def round_tf32(src, dst):
  """dst[i] = src[i] rounded to nearest TF32 (flat contiguous buffers)."""
  _chk(src); _chk(dst)
  _call('sg2im_round_tf32', _p(src), src.numel(), _p(dst), _stream())
  _count()


def adam_flat(params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, weight_decay=0.0,
              found_inf=None, shadow=None, grad_scale=1.0):
  """One Adam update of a flat fp32 bucket in place (torch.optim.Adam arithmetic);
  `step` is a 0-dim device float incremented by the call, `found_inf` (0-dim
  device float or None) nonzero skips update and increment; `shadow` (or None)
  receives the updated parameters rounded to nearest TF32; `grad_scale` multiplies every gradient
  first (1 / world after a SUM all-reduce: the mean without a separate pass over the bucket)."""
  for t in (params, grads, exp_avg, exp_avg_sq):
    _chk(t)
    if not t.is_contiguous() or t.numel() != params.numel():
      raise RuntimeError('sg2im_b200: adam_flat needs four contiguous buffers of equal length')
  _call('sg2im_adam_flat', _p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), params.numel(),
        float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), _p(step),
        _p(found_inf), _p(shadow), float(grad_scale), _stream())
  _count(2)


# --------------------------------------------------------------------------
# weight packing (OIHW master parameters -> kernel layouts)
# --------------------------------------------------------------------------

def pack_conv_fwd(weight):
  """OIHW -> (KH*KW*Cin, Cout) for mode 0."""
  Co, Ci, KH, KW = weight.shape
  return weight.permute(2, 3, 1, 0).reshape(KH * KW * Ci, Co).contiguous()


def pack_conv_dgrad(weight):
  """OIHW -> (KH*KW*Cout, Cin) for mode 1."""
  Co, Ci, KH, KW = weight.shape
  return weight.permute(2, 3, 0, 1).reshape(KH * KW * Co, Ci).contiguous()


# --------------------------------------------------------------------------
# autograd Functions
# --------------------------------------------------------------------------

class Conv(torch.autograd.Function):
  """y = act(conv(x, W) + b), NHWC.  weight is the OIHW master parameter.
  ``in_ch``: use only the first in_ch input channels of the weight (the CRN's
  first stage, whose extra input channel is identically zero)."""

  @staticmethod
  def forward(ctx, x, weight, bias, stride, pad, act, slope, in_ch, out_hw=None,
              zero_bias_grad=False, stats_out=None, round_out=False):
    _chk(weight, name='weight')
    Co, Ci_w, KH, KW = weight.shape
    Ci = Ci_w if in_ch is None else in_ch
    assert x.size(3) == Ci, 'conv: input has %d channels, weight expects %d' % (x.size(3), Ci)
    w_used = weight if Ci == Ci_w else weight[:, :Ci]
    Hout = conv_out_size(x.size(1), KH, stride, pad)
    Wout = conv_out_size(x.size(2), KW, stride, pad)
    if out_hw is not None:
      # cropped output (space-to-depth route of the stride-2 convs): tensor-core only
      assert out_hw[0] <= Hout and out_hw[1] <= Wout and conv_tc_ok(x, KH, KW, stride, pad, Co, out_hw)
      Hout, Wout = out_hw
    fused_stats = stats_out is not None and act == 0 and Co <= 1024
    if _tc_math() and conv_tc_ok(x, KH, KW, stride, pad, Co, (Hout, Wout)):
      y = conv_tc(x, pack_tc_fwd(weight, Ci), bias, KH, KW, pad, Co, act, slope,
                  out_hw=(Hout, Wout), stats=stats_out if fused_stats else None,
                  round_out=round_out)
    else:
      fused_stats = False
      y = conv_igemm(0, x, pack_conv_fwd(w_used), bias, KH, KW, stride, pad, (Hout, Wout), Co,
                     act, slope)
    if stats_out is not None and not fused_stats:
      # per-channel sum / sum of squares of the output for the BatchNorm that follows
      _call_b(4 * y.numel(), 'sg2im_bn_stats', _p(y), y.numel() // Co, Co, _p(stats_out), _stream())
      _count()
    ctx.cfg = (stride, pad, act, slope, Ci, tuple(weight.shape))
    ctx.save_for_backward(x, weight, y if act else None)
    ctx.bias_ref = bias                         # the Parameter itself (its .grad slot), not a saved value
    ctx.has_bias = bias is not None
    ctx.zero_bias_grad = bool(zero_bias_grad)
    return y

  @staticmethod
  def backward(ctx, dy):
    stride, pad, act, slope, Ci, wshape = ctx.cfg
    x, weight, y = ctx.saved_tensors
    Co, Ci_w, KH, KW = wshape
    dy = dy.contiguous()
    dx = dw = db = None
    db_done = False
    if act:
      if FUSE_ACT_BWD and ctx.has_bias and ctx.needs_input_grad[2] and not ctx.zero_bias_grad:
        dy, db = act_bwd_bias(dy, y, slope, ctx.bias_ref)
        db_done = True
      else:
        dy = act_bwd(dy, y, slope)
    if ctx.needs_input_grad[0]:
      w_used = weight if Ci == Ci_w else weight[:, :Ci]
      pad_t = KH - 1 - pad
      if (KH == KW and pad_t >= 0 and stride == 1
          and conv_tc_ok(dy, KH, KW, stride, pad_t, Ci, (x.size(1), x.size(2)))):
        dx = conv_tc(dy, pack_tc_dgrad(weight, Ci), None, KH, KW, pad_t, Ci, tag='conv_dgrad_tc',
                     out_hw=(x.size(1), x.size(2)))
      else:
        dx = conv_igemm(1, dy, pack_conv_dgrad(w_used), None, KH, KW, stride, pad,
                        (x.size(1), x.size(2)), Ci)
    if ctx.needs_input_grad[1]:
      dwp = conv_wgrad(x, dy, KH, KW, stride, pad)
      dw = unpack_wgrad_oihw(dwp, wshape, Ci)
    if ctx.has_bias and ctx.needs_input_grad[2] and not db_done:
      if ctx.zero_bias_grad:
        # this conv feeds a train-mode BatchNorm: d(loss)/d(bias) is identically
        # zero (BN subtracts the batch mean); the reference accumulates rounding
        # noise there.  Skip the full pass over dy.
        db = torch.zeros(Co, dtype=torch.float32, device=dy.device)
      else:
        db = colsum(dy.view(-1, Co))
    return dx, dw, db, None, None, None, None, None, None, None, None, None


def is_kcc(weight):
  """True if an OIHW-shaped (or (out, in)) weight is stored tap-major / input-channel /
  output-channel fastest — the layout the weight-gradient kernel writes and
  sg2im_conv_tc_kcc reads (layers.to_kcc_ puts the parameters there)."""
  if weight.dim() == 2:
    return weight.size(0) > 1 and weight.stride(0) == 1 and weight.t().is_contiguous()
  return (weight.dim() == 4 and weight.stride(0) == 1
          and weight.permute(2, 3, 1, 0).is_contiguous())


def conv_tc_kcc(x, w_kcc, rows_full, dgrad, bias, KH, KW, P, Cout, act=0, slope=0.0, out_hw=None,
                stats=None, round_out=False, tag='conv_fwd_tc'):
  """Tensor-core stride-1 convolution with the weights read in place from the
  weight-gradient layout w_kcc [KH*KW][rows_full][cols] (sg2im_conv_tc_kcc)."""
  N, H, W, C = x.shape
  cs = _pixel_stride(x)
  Hout, Wout = out_hw if out_hw is not None else (H + 2 * P - KH + 1, W + 2 * P - KW + 1)
  out = torch.empty(N, Hout, Wout, Cout, dtype=torch.float32, device=x.device)
  with _prof(tag, 2.0 * N * Hout * Wout * C * Cout * KH * KW, (N, H, W, C, Cout, KH, 1)):
    _call('sg2im_conv_tc_kcc', _p(x), cs, N, H, W, C, _p(w_kcc), rows_full, int(dgrad), _p(bias), KH,
          KW, P, Hout, Wout, Cout, int(act), float(slope), _p(out), Cout, 0, _p(stats),
          int(round_out), _math_id(), _stream())
  _count()
  return out


def conv_tc_presplit(x, w_split, rows_per_tap, bias, KH, KW, P, Cout, act=0, slope=0.0, out_hw=None,
                     stats=None, tag='conv_fwd_tc'):
  """Tensor-core stride-1 convolution (bf16 arithmetic) with a PRE-SPLIT B operand: w_split
  (taps, rows_per_tap, pitch) fp32-typed storage whose rows are 32-channel blocks of
  [32 x bf16 hi | 32 x bf16 mid] (SplitShadows); the kernels then split only the activation tiles."""
  N, H, W, C = x.shape
  cs = _pixel_stride(x)
  Hout, Wout = out_hw if out_hw is not None else (H + 2 * P - KH + 1, W + 2 * P - KW + 1)
  out = torch.empty(N, Hout, Wout, Cout, dtype=torch.float32, device=x.device)
  with _prof(tag, 2.0 * N * Hout * Wout * C * Cout * KH * KW, (N, H, W, C, Cout, KH, 1)):
    _call('sg2im_conv_tc_presplit', _p(x), cs, N, H, W, C, _p(w_split), w_split.size(2), rows_per_tap,
          _p(bias), KH, KW, P, Hout, Wout, Cout, int(act), float(slope), _p(out), Cout, 0, _p(stats),
          _math_id(), _stream())
  _count()
  return out


# Inside TrainStep.step (bf16 arithmetic, weights='kcc'): read the weights' pre-split operand copies
# (SplitShadows, refreshed at the start of every step) instead of splitting weight tiles in-kernel.
USE_SPLIT_SHADOWS = False


class SplitShadows(object):
  """bf16 hi / mid operand copies of the convolution / Linear weights of one network, for the
  'bf16x3' / 'bf16' arithmetic: weights are constant within a training step and re-read by every
  CTA of every launch, so they are split ONCE per step (one launch for the whole network,
  sg2im_split_weights) instead of by the converter warps of every tile.  Each weight stored in the
  weight-gradient layout gets `_split_fwd` (T, Co, cin_pad) and `_split_dgrad` (T, Ci, cout_pad)
  attributes; ops.conv2d / ops.linear pick them up while USE_SPLIT_SHADOWS is set."""

  def __init__(self, params):
    self.weights, rows, tiles = [], [], 0
    for w in params:
      if not (w.dim() in (2, 4) and is_kcc(w) and w.size(0) % 4 == 0):
        continue
      v = _kcc_view(w.detach())
      T, Ci, Co = v.shape
      s2d_c = 0
      if w.dim() == 4 and w.size(2) == 4 and w.size(3) == 4:
        # the discriminators' 4x4 stride-2 filters run as 2x2 stride-1 on the space-to-depth
        # input: the kernel writes the copies in that re-tiled order straight from the master
        if Co % 32:
          continue
        s2d_c, T, Ci = Ci, 4, 4 * Ci
      cip, cop = (Ci + 31) // 32 * 32, (Co + 31) // 32 * 32
      w._split_fwd = torch.zeros(T, Co, cip, dtype=torch.float32, device=w.device)
      w._split_dgrad = torch.zeros(T, Ci, cop, dtype=torch.float32, device=w.device)
      rows.append([v.data_ptr(), w._split_fwd.data_ptr(), w._split_dgrad.data_ptr(), T, Ci, Co, tiles,
                   s2d_c])
      tiles += T * (cip // 32) * (cop // 32)
      self.weights.append(w)
    self.tiles = tiles
    self.table = (torch.tensor(rows, dtype=torch.int64).to(self.weights[0].device)
                  if rows else None)
    self._rows = rows

  def refresh(self):
    if self.table is None:
      return
    for w, r in zip(self.weights, self._rows):       # the masters must not have been re-allocated
      if w.data_ptr() != r[0]:
        raise RuntimeError('sg2im_b200: a parameter moved after SplitShadows was built')
    _call_b(12 * sum(w.numel() for w in self.weights),
            'sg2im_split_weights', _p(self.table), len(self._rows), self.tiles, _stream())
    _count()


def _split_of(weight):
  if USE_SPLIT_SHADOWS and CONV_MATH in ('bf16x3', 'bf16'):
    f = getattr(weight, '_split_fwd', None)
    if f is not None:
      return f, weight._split_dgrad
  return None


class ConvKCC(torch.autograd.Function):
  """y = act(conv(x, W) + b) on the tensor-core kernels with W given as w_kcc
  (T, Ci_w, Co) — the weight-gradient layout — so that neither the forward, nor the
  data gradient, nor the weight gradient needs a layout pass: the forward reads it
  MN-major, the data gradient K-major with the tap index flipped, and the weight
  gradient kernel's output IS the gradient w.r.t. w_kcc.  Stride 1, tf32 only;
  `in_ch` < Ci_w uses a channel prefix (the CRN's first stage)."""

  @staticmethod
  def forward(ctx, x, w_kcc, bias, KH, KW, pad, act, slope, in_ch, out_hw, zero_bias_grad,
              stats_out, round_out, grad_into=None, w_read=None, split=None, s2d_c=0):
    T, Ci_w, Co = w_kcc.shape
    if s2d_c:
      # w_kcc is a 4x4 stride-2 filter (16, C, Co) and x the space-to-depth form of its input:
      # the pre-split copies are already re-tiled (SplitShadows), the weight gradient is written
      # in the filter's own order into its bucket slot — w_kcc itself is never read
      assert (T, Ci_w) == (16, s2d_c) and split is not None and w_read is None and in_ch is None \
          and (grad_into is not None or not ctx.needs_input_grad[1])
      T, Ci_w = 4, 4 * s2d_c
    if w_read is not None:
      # RN-TF32 shadow of the same weights (FlatAdam keeps it current): what the kernels read
      assert w_read.shape == w_kcc.shape and w_read.is_contiguous()
    else:
      w_read = w_kcc
    Ci = Ci_w if in_ch is None else in_ch
    assert T == KH * KW and x.size(3) == Ci and w_kcc.is_contiguous()
    Hout = x.size(1) + 2 * pad - KH + 1 if out_hw is None else out_hw[0]
    Wout = x.size(2) + 2 * pad - KW + 1 if out_hw is None else out_hw[1]
    fused_stats = stats_out is not None and act == 0 and Co <= 1024
    if split is not None:
      # pre-split operand copies of the same weights (SplitShadows): (T, Co, cin_pad) forward
      y = conv_tc_presplit(x, split[0], Co, bias, KH, KW, pad, Co, act, slope, (Hout, Wout),
                           stats_out if fused_stats else None)
    else:
      y = conv_tc_kcc(x, w_read, Ci_w, 0, bias, KH, KW, pad, Co, act, slope, (Hout, Wout),
                      stats_out if fused_stats else None, round_out)
    if stats_out is not None and not fused_stats:
      _call_b(4 * y.numel(), 'sg2im_bn_stats', _p(y), y.numel() // Co, Co, _p(stats_out), _stream())
      _count()
    ctx.cfg = (KH, KW, pad, act, slope, Ci, Ci_w, Co, int(s2d_c))
    ctx.split_dgrad = None if split is None else split[1]
    ctx.save_for_backward(x, w_read, y if act else None)
    ctx.bias_ref = bias
    ctx.has_bias = bias is not None
    ctx.zero_bias_grad = bool(zero_bias_grad)
    # the (T, Ci_w, Co) view of the parameter's slot in the flat gradient bucket, or None
    ctx.grad_into = grad_into if (grad_into is not None and Ci == Ci_w) else None
    return y

  @staticmethod
  def backward(ctx, dy):
    KH, KW, pad, act, slope, Ci, Ci_w, Co, s2d_c = ctx.cfg
    x, w_kcc, y = ctx.saved_tensors
    dy = dy.contiguous()
    dx = dw = db = None
    db_done = False
    if act:
      if FUSE_ACT_BWD and ctx.has_bias and ctx.needs_input_grad[2] and not ctx.zero_bias_grad:
        dy, db = act_bwd_bias(dy, y, slope, ctx.bias_ref)
        db_done = True
      else:
        dy = act_bwd(dy, y, slope)
    if ctx.needs_input_grad[0]:
      pad_t = KH - 1 - pad
      if (KH == KW and pad_t >= 0
          and conv_tc_ok(dy, KH, KW, 1, pad_t, Ci, (x.size(1), x.size(2)))):
        if ctx.split_dgrad is not None:                      # (T flipped, Ci_w, cout_pad)
          dx = conv_tc_presplit(dy, ctx.split_dgrad, Ci_w, None, KH, KW, pad_t, Ci,
                                out_hw=(x.size(1), x.size(2)), tag='conv_dgrad_tc')
        else:
          dx = conv_tc_kcc(dy, w_kcc, Ci_w, 1, None, KH, KW, pad_t, Ci, out_hw=(x.size(1), x.size(2)),
                           tag='conv_dgrad_tc')
      else:
        assert not s2d_c
        wd = w_kcc[:, :Ci].permute(0, 2, 1).reshape(KH * KW * Co, Ci).contiguous()   # exact-fp32 kernel
        dx = conv_igemm(1, dy, wd, None, KH, KW, 1, pad, (x.size(1), x.size(2)), Ci)
    if ctx.needs_input_grad[1] and ctx.grad_into is not None:
      # the weight-gradient kernel adds straight into the gradient bucket: no temporary, no zero
      # fill, no autograd accumulation kernel (the returned gradient is None)
      conv_wgrad(x, dy, KH, KW, 1, pad, accumulate_into=ctx.grad_into, s2d_c=s2d_c)
    elif ctx.needs_input_grad[1]:
      dwp = conv_wgrad(x, dy, KH, KW, 1, pad)                # (T*Ci, Co): already the w_kcc layout
      if Ci == Ci_w:
        dw = dwp.view(KH * KW, Ci_w, Co)
      else:
        dw = torch.zeros(KH * KW, Ci_w, Co, dtype=torch.float32, device=dwp.device)
        dw[:, :Ci] = dwp.view(KH * KW, Ci, Co)
    if ctx.has_bias and ctx.needs_input_grad[2] and not db_done:
      if ctx.zero_bias_grad:
        # exact zero (conv feeds a train-mode BatchNorm).  With the flat gradient bucket zeroed at
        # the start of the step there is nothing to add: skip the fill and the accumulate kernel
        db = None if DIRECT_WGRAD else torch.zeros(Co, dtype=torch.float32, device=dy.device)
      else:
        db = colsum(dy.view(-1, Co))
    return (dx, dw, db) + (None,) * 14


class Pool2d(torch.autograd.Function):
  """nn.AvgPool2d (mode 0) / nn.MaxPool2d (mode 1) with kernel_size = stride = factor on an
  NHWC tensor — build_cnn's 'PX' token (sg2im/layers.py:195-201).  Floor mode: trailing rows /
  columns that do not fill a window are dropped and get zero gradient."""

  @staticmethod
  def forward(ctx, h, factor, mode):
    h = _chk(h).contiguous()
    N, H, W, C = h.shape
    out = torch.empty(N, H // factor, W // factor, C, dtype=torch.float32, device=h.device)
    _call_b(4 * (h.numel() + out.numel()),
            'sg2im_pool2d_fwd', _p(h), N, H, W, C, int(factor), int(mode), _p(out), _stream())
    _count()
    ctx.cfg = (N, H, W, C, int(factor), int(mode))
    ctx.save_for_backward(h if mode == 1 else None)        # max: the arg-max is re-derived from x
    return out

  @staticmethod
  def backward(ctx, dy):
    N, H, W, C, f, mode = ctx.cfg
    x, = ctx.saved_tensors
    dy = dy.contiguous()
    ragged = H % f != 0 or W % f != 0
    dx = (torch.zeros if ragged else torch.empty)(N, H, W, C, dtype=torch.float32, device=dy.device)
    _call_b(4 * (dx.numel() * (2 if mode == 1 else 1) + dy.numel()),
            'sg2im_pool2d_bwd', _p(dy), _p(x), N, H, W, C, f, mode, _p(dx), _stream())
    _count()
    return dx, None, None


class S2D(torch.autograd.Function):
  """Space-to-depth by 2 with zero padding to even size: (N,H,W,C) ->
  (N,ceil(H/2),ceil(W/2),4C), channel = ((y&1)*2+(x&1))*C + c."""

  @staticmethod
  def forward(ctx, x):
    _chk(x)
    N, H, W, C = x.shape
    out = torch.empty(N, (H + 1) // 2, (W + 1) // 2, 4 * C, dtype=torch.float32, device=x.device)
    sn, sh, sw, sc = x.stride()
    _call_b(4 * (x.numel() + out.numel()),
            'sg2im_s2d_fwd', _p(x), sn, sh, sw, sc, N, H, W, C, _p(out), _stream())
    _count()
    ctx.shape = (N, H, W, C)
    return out

  @staticmethod
  def backward(ctx, dout):
    N, H, W, C = ctx.shape
    dout = dout.contiguous()
    dx = torch.empty(N, H, W, C, dtype=torch.float32, device=dout.device)
    _call_b(4 * (dx.numel() + dout.numel()),
            'sg2im_s2d_bwd', _p(dout), N, H, W, C, _p(dx), _stream())
    _count()
    return dx


# Training with flat gradient buckets (train_step.FlatGrads) in the kcc weight layout: let the
# weight-gradient kernels accumulate directly into the parameter's .grad slot.  Switched on by
# TrainStep(weights='kcc'); off for plain autograd use (hooks / create_graph expect returned grads).
DIRECT_WGRAD = False


def _grad_slot(weight):
  g = weight.grad
  if DIRECT_WGRAD and g is not None and weight.requires_grad and g.shape == weight.shape \
      and g.stride() == weight.stride() and is_kcc(g):
    return _kcc_view(g)
  return None


def _shadow(weight):
  """RN-TF32 shadow of a parameter (FlatAdam keeps it current): what the 'tf32' kernels read —
  the hardware would truncate the fp32 master.  The bf16 modes split the master itself."""
  sh = getattr(weight, '_tc_shadow', None) if CONV_MATH == 'tf32' else None
  return None if sh is None else _kcc_view(sh)


def _kcc_view(weight):
  """(T, Ci, Co) view of a weight stored in the weight-gradient layout (zero-copy,
  differentiable: the gradient flows back through the view ops with matching strides)."""
  if weight.dim() == 2:
    return weight.t().unsqueeze(0)
  Co, Ci, KH, KW = weight.shape
  return weight.permute(2, 3, 1, 0).reshape(KH * KW, Ci, Co)


def conv2d(x, weight, bias, stride=1, pad=0, act=0, slope=0.0, in_ch=None, feeds_bn=False,
           stats_out=None, round_out=False):
  """stats_out: zeroed float64 [2*Cout]; receives the per-channel sum and sum of
  squares of the output (fused into the tensor-core epilogue when possible).
  round_out: the output is consumed directly by another tensor-core op (no
  normalise/activate pass in between that would round it): write RN-TF32 values."""
  round_out = bool(round_out) and CONV_MATH == 'tf32'
  Co, C, KH, KW = weight.shape
  kcc = _tc_math() and is_kcc(weight)
  s2d = (_tc_math() and stride == 2 and pad == 0 and in_ch is None
         and KH == 4 and KW == 4 and x.size(1) >= 4 and x.size(2) >= 4 and Co % 32 == 0)
  if s2d:
    # 4x4 stride-2 'valid' conv (the discriminators, scripts/train.py:122-130) ==
    # 2x2 stride-1 conv on the space-to-depth input: runs on the tensor-core
    # kernels (forward, dgrad, wgrad) with no strided gathers.
    Ho, Wo = conv_out_size(x.size(1), 4, 2, 0), conv_out_size(x.size(2), 4, 2, 0)
    xs = S2D.apply(x)
    split, slot = _split_of(weight), _grad_slot(weight)
    if (kcc and split is not None and (slot is not None or not weight.requires_grad)
        and split[0].size(0) == 4
        and conv_tc_ok(xs, 2, 2, 1, 0, Co, (Ho, Wo)) and xs.data_ptr() % 16 == 0
        and _lib.load().sg2im_conv_wgrad_tc_supported(xs.size(0), xs.size(1), xs.size(2), 4 * C, 4 * C,
                                                      2, 2, 1, 0, Ho, Wo, Co)):
      # training step: re-tiled pre-split copies (no per-call copy of the filter) and the weight
      # gradient accumulated in the filter's own order into its slot of the gradient bucket
      return ConvKCC.apply(xs, _kcc_view(weight), bias, 2, 2, 0, act, slope, None, (Ho, Wo), feeds_bn,
                           stats_out, round_out, slot, None, split, C)
    if kcc:
      # [ky][kx][c][co] -> [(ty,tx)][(py,px,c)][co]: one small copy, no pack pass
      w2 = weight.permute(2, 3, 1, 0).reshape(2, 2, 2, 2, C, Co).permute(0, 2, 1, 3, 4, 5)
      w2 = w2.reshape(4, 4 * C, Co)
      if conv_tc_ok(xs, 2, 2, 1, 0, Co, (Ho, Wo)):
        sh = getattr(weight, '_tc_shadow', None) if CONV_MATH == 'tf32' else None
        w2r = None
        if sh is not None:
          w2r = sh.detach().permute(2, 3, 1, 0).reshape(2, 2, 2, 2, C, Co).permute(0, 2, 1, 3, 4, 5)
          w2r = w2r.reshape(4, 4 * C, Co)
        return ConvKCC.apply(xs, w2, bias, 2, 2, 0, act, slope, None, (Ho, Wo), feeds_bn, stats_out,
                             round_out, None, w2r)
    wc = weight if weight.is_contiguous() else weight.contiguous()
    w2 = wc.view(Co, C, 2, 2, 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(Co, 4 * C, 2, 2)
    return Conv.apply(xs, w2, bias, 1, 0, act, slope, None, (Ho, Wo), feeds_bn, stats_out,
                      round_out)
  if kcc and stride == 1:
    Ci = C if in_ch is None else in_ch
    Hout, Wout = conv_out_size(x.size(1), KH, 1, pad), conv_out_size(x.size(2), KW, 1, pad)
    if conv_tc_ok(x, KH, KW, 1, pad, Co, (Hout, Wout)) and x.size(3) == Ci:
      return ConvKCC.apply(x, _kcc_view(weight), bias, KH, KW, pad, act, slope, in_ch, None, feeds_bn,
                           stats_out, round_out, _grad_slot(weight), _shadow(weight), _split_of(weight))
  return Conv.apply(x, weight, bias, stride, pad, act, slope, in_ch, None, feeds_bn, stats_out,
                    round_out)


def linear(x2d, weight, bias, act=0, slope=0.0, round_out=False):
  """nn.Linear (+ fused ReLU/LeakyReLU) as a 1x1 convolution over rows.
  round_out: the output feeds another tensor-core GEMM (hand over RN-TF32 values)."""
  M, K = x2d.shape
  rnd = bool(round_out) and CONV_MATH == 'tf32'
  x4 = x2d.reshape(M, 1, 1, K)
  if _tc_math() and is_kcc(weight) and conv_tc_ok(x4, 1, 1, 1, 0, weight.size(0), (1, 1)):
    y = ConvKCC.apply(x4, _kcc_view(weight), bias, 1, 1, 0, act, slope, None, None, False, None, rnd,
                      _grad_slot(weight), _shadow(weight), _split_of(weight))
  else:
    w4 = weight.reshape(weight.size(0), K, 1, 1) if not weight.is_contiguous() else \
        weight.view(weight.size(0), K, 1, 1)
    y = Conv.apply(x4, w4, bias, 1, 0, act, slope, None, None, False, None, rnd)
  return y.view(M, weight.size(0))


class BNAct(torch.autograd.Function):
  """y = up_x{up}( leaky_slope( BN(x) ) ) written into out[..., coff:coff+C]
  (out=None -> fresh tensor).  BN optional (gamma None and no running stats ->
  plain activation/upsample).  Train mode uses batch statistics and updates the
  running buffers in place, like nn.BatchNorm2d."""

  @staticmethod
  def forward(ctx, x, gamma, beta, running_mean, running_var, use_bn, training, momentum, eps,
              slope, up, unbias_mult, out, out_coff, sums=None, nbt=None):
    x = _chk(x).contiguous()
    scale = shift = save = None
    if use_bn:
      scale, shift, save = bn_scale_shift(x, gamma, beta, running_mean, running_var, training,
                                          momentum, eps, unbias_mult, sums if training else None,
                                          nbt if training else None)
    y = scale_act_fwd(x, scale, shift, slope, up, out, out_coff)
    if out is not None:
      ctx.mark_dirty(out)
    ctx.cfg = (use_bn, training, slope, up, out_coff, out is not None)
    ctx.params = (gamma, beta)                    # the Parameters themselves (their .grad slots)
    ctx.save_for_backward(x, scale, shift, save)
    return y

  @staticmethod
  def backward(ctx, dy):
    use_bn, training, slope, up, coff, sliced = ctx.cfg
    x, scale, shift, save = ctx.saved_tensors
    dy = dy.contiguous()
    want_pg = use_bn and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
    into = None
    if want_pg and DIRECT_WGRAD and ctx.needs_input_grad[1] and ctx.needs_input_grad[2]:
      gg, gb = (getattr(p, 'grad', None) for p in ctx.params)
      C = x.size(3)
      if (gg is not None and gb is not None and gg.is_contiguous() and gb.is_contiguous()
          and gg.numel() == C and gb.numel() == C and gg.dtype == torch.float32):
        into = (gg, gb)                           # slots of the flat gradient bucket: add in place
    dx, dgamma, dbeta = scale_act_bwd(dy, coff, x, scale, shift, save, slope, up,
                                      training and use_bn, want_pg, into)
    dout = dy if sliced else None
    return (dx, dgamma, dbeta, None, None, None, None, None, None, None, None, None, dout, None,
            None, None)


def bn_act(x, bn=None, slope=1.0, up=1, unbias_mult=1, out=None, out_coff=0, sums=None):
  """bn: an nn.BatchNorm2d-like module (weight, bias, running_mean, running_var,
  training, momentum, eps, num_batches_tracked) or None."""
  if bn is None:
    return BNAct.apply(x, None, None, None, None, False, False, 0.0, 0.0, slope, up, 1, out,
                       out_coff)
  training = bn.training or bn.running_mean is None
  momentum = 0.1 if bn.momentum is None else bn.momentum
  # nn.BatchNorm's num_batches_tracked += 1 happens inside the finalize kernel (one launch less per layer)
  nbt = bn.num_batches_tracked if (training and bn.num_batches_tracked is not None) else None
  if nbt is not None and nbt.dtype != torch.int64:
    nbt.add_(1)
    nbt = None
  return BNAct.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, True, training,
                     momentum, bn.eps, slope, up, unbias_mult, out, out_coff, sums, nbt)


def new_stats(channels, device):
  return _zeros(2 * channels, torch.float64, device)


class TripleGather(torch.autograd.Function):
  """cur_t_vecs = cat([obj[s], pred, obj[o]], 1)   (sg2im/graph.py:77-82)."""

  @staticmethod
  def forward(ctx, obj_vecs, pred_vecs, edges, csr):
    out = triple_gather(obj_vecs, pred_vecs, edges, pred_vecs.size(1))
    ctx.csr = csr
    ctx.dims = (obj_vecs.size(0), obj_vecs.size(1), pred_vecs.size(1))
    return out

  @staticmethod
  def backward(ctx, dcur):
    O, D, Dp = ctx.dims
    row_ptr, entries = ctx.csr
    dcur = dcur.contiguous()
    dobj = segment_sum(dcur, 0, D + Dp, D, row_ptr, entries, O, False)
    dpred = dcur[:, D:D + Dp]
    return dobj, dpred, None, None


class GraphPool(torch.autograd.Function):
  """(pooled (O,H), new_p (T,Dout)) from new_t_vecs (T, 2H+Dout)
  (sg2im/graph.py:85-114); bit-exact vs the CPU scatter_add order."""

  @staticmethod
  def forward(ctx, new_t, edges, csr, H, Dout, num_objs, avg):
    row_ptr, entries = csr
    pooled = segment_sum(new_t, 0, H + Dout, H, row_ptr, entries, num_objs, avg)
    new_p = new_t[:, H:H + Dout].contiguous()
    ctx.csr = csr
    ctx.edges = edges
    ctx.cfg = (H, Dout, avg)
    return pooled, new_p

  @staticmethod
  def backward(ctx, dpooled, dnew_p):
    H, Dout, avg = ctx.cfg
    row_ptr, _ = ctx.csr
    if dpooled is None:
      dpooled = _zeros((row_ptr.numel() - 1, H), torch.float32, row_ptr.device)
    dnew_t = triple_gather(dpooled, dnew_p, ctx.edges, Dout, row_ptr if avg else None)
    return dnew_t, None, None, None, None, None, None


class Layout(torch.autograd.Function):
  """Fused masks_to_layout / boxes_to_layout (sg2im/layout.py:30-91) + the
  noise concat of model.py:164-169.  Output (N,H,W,D+noise_c) NHWC."""

  @staticmethod
  def forward(ctx, vecs, boxes, masks, obj_to_img, N, H, W, noise, align_corners):
    vecs = _chk(vecs).contiguous()
    boxes = _chk(boxes).contiguous()
    O, D = vecs.shape
    M = 0
    if masks is not None:
      masks = _chk(masks.float() if masks.dtype != torch.float32 else masks).contiguous()
      M = masks.size(1)
    nc = 0 if noise is None else noise.size(1)
    out = torch.empty(N, H, W, D + nc, dtype=torch.float32, device=vecs.device)
    _layout_launch(vecs, boxes, masks, obj_to_img, N, H, W, noise, align_corners, out)
    ctx.save_for_backward(vecs, boxes, masks, obj_to_img)
    ctx.cfg = (N, H, W, M, align_corners)
    return out

  @staticmethod
  def backward(ctx, dout):
    vecs, boxes, masks, obj_to_img = ctx.saved_tensors
    N, H, W, M, align = ctx.cfg
    dout = dout.contiguous()
    O, D = vecs.shape
    dvecs = _zeros(vecs.shape, torch.float32, vecs.device)
    dmasks = None
    if masks is not None and ctx.needs_input_grad[2]:
      dmasks = _zeros(masks.shape, torch.float32, masks.device)
    _call_b(4 * N * H * W * D + 8 * O * D,
            'sg2im_layout_bwd', _p(dout), dout.size(3), _p(vecs), _p(boxes), _p(masks), M,
            _p(obj_to_img), N, O, D, H, W, int(align), _p(dvecs), _p(dmasks), _stream())
    _count()
    dboxes = _layout_dboxes(ctx, dout, vecs, boxes, masks, M, obj_to_img, N, H, W, align)
    return dvecs, dboxes, dmasks, None, None, None, None, None, None


def _layout_dboxes(ctx, dout, vecs, boxes, masks, M, obj_to_img, N, H, W, align):
  """Gradient w.r.t. the boxes (training on predicted boxes, model.py:151-160) or None."""
  if not ctx.needs_input_grad[1]:
    return None
  O, D = vecs.shape
  dboxes = torch.empty(O, 4, dtype=torch.float32, device=vecs.device)
  _call_b(4 * N * H * W * D + 4 * O * (D + 8), 'sg2im_layout_bwd_boxes', _p(dout), dout.size(3),
          _p(vecs), _p(boxes), _p(masks), M, _p(obj_to_img), N, O, D, H, W, int(align), _p(dboxes),
          _stream())
  _count()
  return dboxes


def _layout_launch(vecs, boxes, masks, obj_to_img, N, H, W, noise, align_corners, out,
                   round_tf32=False):
  O, D = vecs.shape
  M = 0 if masks is None else masks.size(1)
  img_ptr, img_ent = csr_build(obj_to_img, 1, N)
  nc = 0 if noise is None else noise.size(1)
  ns = (0, 0, 0, 0) if noise is None else noise.stride()      # (n, c, h, w)
  # algorithmic bytes: the D + nc channels WRITTEN per pixel (out may be a wider stage buffer whose
  # other channels another kernel fills), the nc noise channels read, vectors / boxes / masks read
  _call_b(4 * N * H * W * (D + 2 * nc) + 4 * O * (D + 4 + M * M),
          'sg2im_layout_fwd', _p(vecs), _p(boxes), _p(masks), M, _p(img_ptr), _p(img_ent), N, O,
          D, H, W, int(align_corners), _p(noise), nc, ns[0], ns[1], ns[2], ns[3], _p(out),
          out.size(3), int(round_tf32), _stream())
  _count()


class LayoutStack(torch.autograd.Function):
  """Layout + the input buffers of every cascaded-refinement stage in one go.

  Level k (k = 0 coarsest ... L-1 full resolution) is an NHWC buffer of
  C + extras[k] channels, C = D + noise channels: the first C hold the layout
  average-pooled by a 2x2 cascade (F.avg_pool2d, sg2im/crn.py:58-62; the
  full-resolution level is written by the layout kernel itself), the trailing
  ``extras[k]`` channels are left for the previous stage's upsampled features,
  which BNAct later writes in place (the torch.cat of crn.py:63 never runs).
  Backward folds the per-level layout gradients back down the cascade into the
  finest level's buffer and runs the layout backward on that slice."""

  @staticmethod
  def forward(ctx, vecs, boxes, masks, obj_to_img, N, H, W, noise, align_corners, extras):
    vecs = _chk(vecs).contiguous()
    boxes = _chk(boxes).contiguous()
    if masks is not None:
      masks = _chk(masks.float() if masks.dtype != torch.float32 else masks).contiguous()
    D = vecs.size(1)
    C = D + (0 if noise is None else noise.size(1))
    L = len(extras)
    dev = vecs.device
    bufs = [None] * L
    bufs[L - 1] = torch.empty(N, H, W, C + extras[L - 1], dtype=torch.float32, device=dev)
    _layout_launch(vecs, boxes, masks, obj_to_img, N, H, W, noise, align_corners, bufs[L - 1],
                   round_tf32=CONV_MATH == 'tf32')
    for k in range(L - 2, -1, -1):
      f = L - 1 - k
      bufs[k] = torch.empty(N, H >> f, W >> f, C + extras[k], dtype=torch.float32, device=dev)
      avgpool2_fwd(bufs[k + 1], 0, C, bufs[k], 0)
    ctx.save_for_backward(vecs, boxes, masks, obj_to_img)
    ctx.cfg = (N, H, W, C, align_corners)
    return tuple(bufs)

  @staticmethod
  def backward(ctx, *douts):
    if GRAD_READY_HOOK is not None:
      GRAD_READY_HOOK()
    vecs, boxes, masks, obj_to_img = ctx.saved_tensors
    N, H, W, C, align = ctx.cfg
    L = len(douts)
    g = [d.contiguous() for d in douts]
    # the stage gradients are private temporaries of this backward pass (fresh
    # conv-dgrad outputs): accumulate the coarser levels into the finer in place
    for k in range(1, L):
      avgpool2_bwd(g[k - 1], 0, C, g[k], 0, True)
    O, D = vecs.shape
    M = 0 if masks is None else masks.size(1)
    dvecs = _zeros(vecs.shape, torch.float32, vecs.device)
    dmasks = None
    if masks is not None and ctx.needs_input_grad[2]:
      dmasks = _zeros(masks.shape, torch.float32, masks.device)
    _call_b(4 * N * H * W * D + 8 * O * D,
            'sg2im_layout_bwd', _p(g[L - 1]), g[L - 1].size(3), _p(vecs), _p(boxes), _p(masks), M,
            _p(obj_to_img), N, O, D, H, W, int(align), _p(dvecs), _p(dmasks), _stream())
    _count()
    dboxes = _layout_dboxes(ctx, g[L - 1], vecs, boxes, masks, M, obj_to_img, N, H, W, align)
    return dvecs, dboxes, dmasks, None, None, None, None, None, None, None


class BCELogitsMean(torch.autograd.Function):
  """mean(max(x, 0) - x * t + log(1 + exp(-|x|))) for a constant target t — bce_loss(scores,
  ones / zeros) of sg2im/losses.py:39-57 — as one forward and one backward pass."""

  @staticmethod
  def forward(ctx, x, target):
    x = _chk(x).contiguous()
    n = x.numel()
    out = torch.empty((), dtype=torch.float32, device=x.device)
    scratch = _zeros(1, torch.float64, x.device)
    _call('sg2im_bce_logits_mean_fwd', _p(x), n, float(target), _p(scratch), _p(out), _stream())
    _count(2)
    ctx.save_for_backward(x)
    ctx.target = float(target)
    return out

  @staticmethod
  def backward(ctx, gout):
    x, = ctx.saved_tensors
    gout = gout.contiguous().to(torch.float32)
    dx = torch.empty_like(x)
    _call('sg2im_bce_logits_mean_bwd', _p(x), x.numel(), ctx.target, _p(gout), _p(dx), _stream())
    _count()
    return dx, None


class Crop(torch.autograd.Function):
  """crop_bbox_batch (sg2im/bilinear.py:28-132).  feats: (N,H,W,C)-shaped view
  with any strides; returns (B,HH,WW,C) NHWC."""

  @staticmethod
  def forward(ctx, feats, boxes, idx, HH, WW, align_corners):
    _chk(feats)
    if ctx.needs_input_grad[1]:
      # the reference crops with the batch's ground-truth boxes (scripts/train.py:539,569-570); a
      # gradient w.r.t. the crop boxes is not implemented and must not be dropped silently
      raise RuntimeError('sg2im_b200: crop boxes that require grad are not supported')
    boxes = _chk(boxes).contiguous()
    idx = _chk(idx, torch.int64, 'bbox_to_feats').contiguous()
    N, H, W, C = feats.shape
    B = boxes.size(0)
    out = torch.empty(B, HH, WW, C, dtype=torch.float32, device=feats.device)
    sn, sh, sw, sc = feats.stride()
    _call_b(4 * (N * H * W * C + out.numel()),
            'sg2im_crop_fwd', _p(feats), sn, sh, sw, sc, N, H, W, C, _p(boxes), _p(idx), B, HH, WW,
            int(align_corners), _p(out), _stream())
    _count()
    ctx.save_for_backward(boxes, idx)
    ctx.cfg = (N, H, W, C, HH, WW, align_corners)
    return out

  @staticmethod
  def backward(ctx, dout):
    boxes, idx = ctx.saved_tensors
    N, H, W, C, HH, WW, align = ctx.cfg
    dout = dout.contiguous()
    dfeats = torch.zeros(N, H, W, C, dtype=torch.float32, device=dout.device)
    _call_b(4 * (dfeats.numel() + dout.numel()),
            'sg2im_crop_bwd', _p(dout), _p(boxes), _p(idx), N, H, W, C, boxes.size(0), HH, WW,
            int(align), _p(dfeats), _stream())
    _count()
    return dfeats, None, None, None, None, None
