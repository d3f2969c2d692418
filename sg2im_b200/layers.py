"""Layer factories with the reference's surface (sg2im/layers.py): ``build_mlp``,
``build_cnn``, ``get_normalization_2d``, ``get_activation`` and the small
modules they return — same constructor arguments, same ``nn.Sequential``
structure and therefore the same ``state_dict`` keys — whose forward passes run
on the sm_100a kernels of libsg2im_b200.so.

Callers see the reference's NCHW shapes; internally every 4-D activation is an
NHWC tensor (a permuted view, i.e. torch ``channels_last`` memory), so no
layout conversion kernels run between layers.
"""
import torch
import torch.nn as nn

from . import ops


def _to_nhwc(x):
  """NCHW-shaped tensor (any strides) -> (N,H,W,C)-shaped view."""
  return x.permute(0, 2, 3, 1)


def _to_nchw(h):
  return h.permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------
# leaf modules: parameter containers identical to torch's, forward on our kernels
# ----------------------------------------------------------------------------

class Conv2d(nn.Conv2d):
  """nn.Conv2d parameters (OIHW fp32 master weights); square stride/padding,
  dilation 1, groups 1 — everything sg2im builds (layers.py:178, crn.py:41-45)."""

  def _cfg(self):
    if (self.dilation != (1, 1) or self.groups != 1 or self.stride[0] != self.stride[1]
        or self.padding[0] != self.padding[1] or self.padding_mode != 'zeros'):
      raise NotImplementedError('sg2im_b200.Conv2d: only square stride/zero padding, no '
                                'dilation/groups (all the reference uses)')
    return self.stride[0], self.padding[0]

  def forward_nhwc(self, h, act=0, slope=0.0, in_ch=None, feeds_bn=False, stats_out=None):
    """feeds_bn: the output goes straight into a train-mode BatchNorm (its
    bias gradient is identically zero and is not computed); stats_out receives
    that BatchNorm's batch statistics from the conv epilogue."""
    stride, pad = self._cfg()
    return ops.conv2d(h, self.weight, self.bias, stride, pad, act, slope, in_ch, feeds_bn,
                      stats_out)

  def forward(self, x):
    return _to_nchw(self.forward_nhwc(_to_nhwc(x)))


class Linear(nn.Linear):
  def forward_act(self, x, act=0, slope=0.0, round_out=False):
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.size(-1))
    nout = self.out_features
    pad = (-nout) % 4
    if pad and ops._tc_math() and nout >= 32 and self.in_features % 4 == 0:
      # e.g. the object classifier (1024 -> num_objects = 179): pad the output
      # width to a multiple of 4 with zero rows so the GEMM and its data
      # gradient run on the tensor-core kernel; the extra columns are sliced off
      w = torch.nn.functional.pad(self.weight, (0, 0, 0, pad))
      b = None if self.bias is None else torch.nn.functional.pad(self.bias, (0, pad))
      y = ops.linear(x2, w, b, act, slope)[:, :nout]
    else:
      y = ops.linear(x2, self.weight, self.bias, act, slope, round_out)
    return y.reshape(*lead, nout)

  def forward(self, x):
    return self.forward_act(x)


class BatchNorm2d(nn.BatchNorm2d):
  def forward_nhwc(self, h, slope=1.0, up=1, unbias_mult=1, out=None, out_coff=0, sums=None):
    return ops.bn_act(h, self, slope, up, unbias_mult, out, out_coff, sums)

  def forward(self, x):
    return _to_nchw(self.forward_nhwc(_to_nhwc(x).contiguous()))


class InstanceNorm2d(nn.InstanceNorm2d):
  """get_normalization_2d('instance') (sg2im/layers.py:24): torch's defaults — no affine
  parameters, no running statistics, instance statistics in train AND eval mode, biased variance,
  eps 1e-5; nothing in the state_dict.  Per image this is exactly the train-mode BatchNorm
  arithmetic over that image's H*W rows with gamma = 1, beta = 0, so it runs on the same
  statistics / finalize / normalise-activate(-upsample) kernels, one image per launch set."""

  def forward_nhwc(self, h, slope=1.0, up=1, out=None, out_coff=0):
    if self.affine or self.track_running_stats:
      raise NotImplementedError('sg2im_b200.InstanceNorm2d: torch defaults only (what '
                                'get_normalization_2d builds)')
    h = h.contiguous()
    y = torch.cat([ops.bn_act(h[n:n + 1], self, slope, up) for n in range(h.size(0))], 0)
    if out is None:
      return y
    out[..., out_coff:out_coff + y.size(3)] = y
    return out

  def forward(self, x):
    return _to_nchw(self.forward_nhwc(_to_nhwc(x)))


def norm_act(h, norm, slope=1.0, up=1, out=None, out_coff=0, sums=None):
  """leaky_slope(norm(h)) (+ nearest upsample, optionally written into a channel slice of `out`)
  for norm = BatchNorm2d | InstanceNorm2d | None."""
  if isinstance(norm, InstanceNorm2d):
    return norm.forward_nhwc(h, slope, up, out, out_coff)
  return ops.bn_act(h, norm, slope, up, 1, out, out_coff, sums)


class BatchNorm1d(nn.BatchNorm1d):
  """(rows, C) batch norm for build_mlp(batch_norm='batch')."""

  def forward_act(self, x, slope=1.0):
    M, C = x.shape
    return ops.bn_act(x.contiguous().view(M, 1, 1, C), self, slope).view(M, C)

  def forward(self, x):
    return self.forward_act(x)


class LeakyReLU(nn.LeakyReLU):
  def forward(self, x):
    if x.dim() == 4:
      return _to_nchw(ops.bn_act(_to_nhwc(x).contiguous(), None, self.negative_slope))
    M = x.numel() // x.size(-1)
    return ops.bn_act(x.contiguous().view(M, 1, 1, x.size(-1)), None,
                      self.negative_slope).view(x.shape)


class ReLU(nn.ReLU):
  negative_slope = 0.0

  def forward(self, x):
    return LeakyReLU.forward(self, x)


class Upsample(nn.Upsample):
  """Nearest-neighbour integer upsampling (model.py:98, layers.py 'UX')."""

  def _factor(self):
    f = self.scale_factor
    if self.mode != 'nearest' or f is None or int(f) != f:
      raise NotImplementedError('sg2im_b200.Upsample: integer nearest upsampling only')
    return int(f)

  def forward(self, x):
    return _to_nchw(ops.bn_act(_to_nhwc(x).contiguous(), None, 1.0, self._factor()))


class GlobalAvgPool(nn.Module):
  """sg2im/layers.py:83-86 (layout-agnostic form of view(N,C,-1).mean(2))."""

  def forward(self, x):
    return x.mean(dim=(2, 3))


class Flatten(nn.Module):
  def forward(self, x):
    return x.reshape(x.size(0), -1)

  def __repr__(self):
    return 'Flatten()'


class Unflatten(nn.Module):
  def __init__(self, size):
    super(Unflatten, self).__init__()
    self.size = size

  def forward(self, x):
    return x.view(*self.size)

  def __repr__(self):
    return 'Unflatten(%s)' % ', '.join('%d' % d for d in self.size)


class AvgPool2(nn.AvgPool2d):
  """'P2' with pooling='avg' (layers.py:195-201)."""

  def forward(self, x):
    h = _to_nhwc(x).contiguous()
    N, H, W, C = h.shape
    return _to_nchw(_AvgPool2Fn.apply(h))


def _pool_factor(m):
  """kernel_size = stride = f, no padding / dilation / ceil mode: what build_cnn builds."""
  as2 = lambda v: (v, v) if isinstance(v, int) else tuple(v)            # noqa: E731
  k, st = as2(m.kernel_size), as2(m.stride if m.stride is not None else m.kernel_size)
  ok = (k[0] == k[1] and st == k and as2(m.padding) == (0, 0) and not m.ceil_mode
        and as2(getattr(m, 'dilation', 1)) == (1, 1)
        and getattr(m, 'divisor_override', None) is None and not getattr(m, 'return_indices', False))
  if not ok:
    raise NotImplementedError('sg2im_b200: pooling with kernel_size = stride, no padding only '
                              "(what build_cnn's 'PX' builds)")
  return int(k[0])


class AvgPool2d(nn.AvgPool2d):
  """'PX' with pooling='avg' (layers.py:198-199), any factor."""

  def forward(self, x):
    return _to_nchw(ops.Pool2d.apply(_to_nhwc(x), _pool_factor(self), 0))


class MaxPool2d(nn.MaxPool2d):
  """'PX' with pooling='max' (layers.py:196-197; build_cnn's default pooling)."""

  def forward(self, x):
    return _to_nchw(ops.Pool2d.apply(_to_nhwc(x), _pool_factor(self), 1))


class _AvgPool2Fn(torch.autograd.Function):
  @staticmethod
  def forward(ctx, h):
    N, H, W, C = h.shape
    out = torch.empty(N, H // 2, W // 2, C, dtype=h.dtype, device=h.device)
    ops.avgpool2_fwd(h, 0, C, out, 0)
    ctx.shape = h.shape
    return out

  @staticmethod
  def backward(ctx, dy):
    N, H, W, C = ctx.shape
    dx = torch.empty(N, H, W, C, dtype=dy.dtype, device=dy.device)
    ops.avgpool2_bwd(dy.contiguous(), 0, C, dx, 0, False)
    return dx


def to_kcc_(module):
  """Re-store every Conv2d / Linear weight of `module` tap-major with the output channel
  fastest ([KH][KW][Cin][Cout]; Linear: [in][out]) — the layout the tensor-core weight-gradient
  kernel writes and `sg2im_conv_tc_kcc` reads in place, so a training step needs no weight
  pack / unpack pass (ops.ConvKCC).  The parameters keep their OIHW / (out, in) SHAPE as
  permuted views: `state_dict` keys, shapes and values are unchanged, `load_state_dict`
  copies into the new storage.  TrainStep(weights='kcc') — what bench.py measures; parity on
  hardware in tests/test_gpu_bf16x3.py and tests/test_gpu_next_rows.py.  Returns `module`."""
  with torch.no_grad():
    for m in module.modules():
      if isinstance(m, nn.Conv2d):
        store = m.weight.detach().permute(2, 3, 1, 0).contiguous()
        m.weight.data = store.permute(3, 2, 0, 1)
      elif isinstance(m, nn.Linear):
        store = m.weight.detach().t().contiguous()
        m.weight.data = store.t()
  return module


# ----------------------------------------------------------------------------
# Sequential with peephole fusion
# ----------------------------------------------------------------------------

def _slope_of(m):
  if isinstance(m, nn.LeakyReLU):
    return m.negative_slope
  if isinstance(m, nn.ReLU):
    return 0.0
  return None


class FusedSequential(nn.Sequential):
  """An nn.Sequential (same children, same state_dict keys) whose forward
  walks the children with three fusions, each one kernel instead of two/three:
    Conv2d|Linear -> ReLU|LeakyReLU      : activation in the GEMM epilogue
    BatchNorm -> ReLU|LeakyReLU          : one normalise+activate pass
    Upsample(x2) -> BatchNorm2d          : upsample folded into the BN apply
  Unknown children fall back to their own forward (NCHW view in / out)."""

  def forward(self, x):
    mods = list(self)
    four_d = x.dim() == 4
    h = _to_nhwc(x) if four_d else x
    i = 0
    sums = None                       # batch statistics handed from a conv to the BN after it
    while i < len(mods):
      m = mods[i]
      nxt = mods[i + 1] if i + 1 < len(mods) else None
      s = _slope_of(nxt) if nxt is not None else None
      if isinstance(m, Conv2d) and four_d:
        if s is not None:
          h = m.forward_nhwc(h, 1, s); i += 2
        else:
          fb = isinstance(nxt, BatchNorm2d) and nxt.training
          sums = ops.new_stats(m.out_channels, h.device) if fb else None
          h = m.forward_nhwc(h, feeds_bn=fb, stats_out=sums); i += 1
      elif isinstance(m, Linear) and not four_d:
        if s is not None:
          # the next Linear of the MLP consumes this output on the tensor core
          feeds_gemm = i + 2 < len(mods) and isinstance(mods[i + 2], Linear)
          h = m.forward_act(h, 1, s, round_out=feeds_gemm); i += 2
        else:
          h = m.forward_act(h); i += 1
      elif isinstance(m, BatchNorm2d) and four_d:
        if s is not None:
          h = m.forward_nhwc(h.contiguous(), s, sums=sums); i += 2
        else:
          h = m.forward_nhwc(h.contiguous(), sums=sums); i += 1
        sums = None
      elif isinstance(m, InstanceNorm2d) and four_d:
        if s is not None:
          h = m.forward_nhwc(h, s); i += 2
        else:
          h = m.forward_nhwc(h); i += 1
      elif isinstance(m, BatchNorm1d) and not four_d:
        if s is not None:
          h = m.forward_act(h, s); i += 2
        else:
          h = m.forward_act(h); i += 1
      elif isinstance(m, Upsample) and four_d and isinstance(nxt, BatchNorm2d):
        f = m._factor()
        h = nxt.forward_nhwc(h.contiguous(), 1.0, f, f * f); i += 2
      elif isinstance(m, (Flatten, Unflatten)):
        h = m(_to_nchw(h) if four_d else h)
        four_d = h.dim() == 4
        h = _to_nhwc(h) if four_d else h
        i += 1
      else:
        y = m(_to_nchw(h) if four_d else h)
        four_d = y.dim() == 4
        h = _to_nhwc(y) if four_d else y
        i += 1
    return _to_nchw(h) if four_d else h


# ----------------------------------------------------------------------------
# factories (reference signatures)
# ----------------------------------------------------------------------------

def get_normalization_2d(channels, normalization):
  """sg2im/layers.py:22-30."""
  if normalization == 'instance':
    return InstanceNorm2d(channels)
  elif normalization == 'batch':
    return BatchNorm2d(channels)
  elif normalization == 'none':
    return None
  else:
    raise ValueError('Unrecognized normalization type "%s"' % normalization)


def get_activation(name):
  """sg2im/layers.py:33-46 — including its quirk: the reference reassigns
  name = 'leakyrelu' before the lookup, so 'relu' also yields LeakyReLU (slope
  0.01).  Reproduced for result compatibility."""
  kwargs = {}
  if name.lower().startswith('leakyrelu'):
    if '-' in name:
      kwargs = {'negative_slope': float(name.split('-')[1])}
  return LeakyReLU(**kwargs)


def _init_conv(layer, method):
  if not isinstance(layer, nn.Conv2d) or method == 'default':
    return
  if method == 'kaiming-normal':
    nn.init.kaiming_normal_(layer.weight)
  elif method == 'kaiming-uniform':
    nn.init.kaiming_uniform_(layer.weight)


def _get_padding(K, mode):
  """sg2im/layers.py:120-126."""
  if mode == 'valid':
    return 0
  elif mode == 'same':
    assert K % 2 == 1, 'Invalid kernel size %d for "same" padding' % K
    return (K - 1) // 2


class ResidualBlock(nn.Module):
  """sg2im/layers.py:89-117, quirks included: `self.net` is evaluated TWICE per forward
  (:115-116; the first result is discarded, but train-mode BatchNorm running statistics and
  `num_batches_tracked` advance twice), and with padding 0 the shortcut is the EMPTY slice
  x[:, :, 0:-0, 0:-0] (:112-114), so a 'valid' residual block fails in the addition exactly like
  the reference's.  Same children / state_dict keys (net.0 .. net.5)."""

  def __init__(self, channels, normalization='batch', activation='relu',
               padding='same', kernel_size=3, init='default'):
    super(ResidualBlock, self).__init__()
    K = kernel_size
    P = _get_padding(K, padding)
    C = channels
    self.padding = P
    layers = [
      get_normalization_2d(C, normalization),
      get_activation(activation),
      Conv2d(C, C, kernel_size=K, padding=P),
      get_normalization_2d(C, normalization),
      get_activation(activation),
      Conv2d(C, C, kernel_size=K, padding=P),
    ]
    layers = [layer for layer in layers if layer is not None]
    for layer in layers:
      _init_conv(layer, method=init)
    self.net = FusedSequential(*layers)

  def forward(self, x):
    P = self.padding
    shortcut = x
    if P == 0:
      shortcut = x[:, :, P:-P, P:-P]
    if self.training and any(isinstance(m, nn.BatchNorm2d) for m in self.net):
      with torch.no_grad():
        self.net(x)                  # the reference's discarded first evaluation (side effects only)
    return shortcut + self.net(x)


def build_cnn(arch, normalization='batch', activation='relu', padding='same',
              pooling='max', init='default'):
  """sg2im/layers.py:129-213: architecture-string CNN builder.
  Every token — IX / CK-X[-S] / R / UX / PX (max or average) / FC-X-Y — is built from modules
  whose forward and backward run on the sm_100a kernels.  Returns (nn.Sequential, channels)."""
  if isinstance(arch, str):
    arch = arch.split(',')
  cur_C = 3
  if len(arch) > 0 and arch[0][0] == 'I':
    cur_C = int(arch[0][1:])
    arch = arch[1:]

  first_conv = True
  flat = False
  layers = []
  for i, s in enumerate(arch):
    if s[0] == 'C':
      if not first_conv:
        layers.append(get_normalization_2d(cur_C, normalization))
        layers.append(get_activation(activation))
      first_conv = False
      vals = [int(v) for v in s[1:].split('-')]
      if len(vals) == 2:
        K, next_C = vals
        stride = 1
      elif len(vals) == 3:
        K, next_C, stride = vals
      P = _get_padding(K, padding)
      conv = Conv2d(cur_C, next_C, kernel_size=K, padding=P, stride=stride)
      layers.append(conv)
      _init_conv(layers[-1], init)
      cur_C = next_C
    elif s[0] == 'R':
      norm = 'none' if first_conv else normalization
      res = ResidualBlock(cur_C, normalization=norm, activation=activation,
                          padding=padding, init=init)
      layers.append(res)
      first_conv = False
    elif s[0] == 'U':
      layers.append(Upsample(scale_factor=int(s[1:]), mode='nearest'))
    elif s[0] == 'P':
      factor = int(s[1:])
      if pooling == 'avg' and factor == 2:
        layers.append(AvgPool2(kernel_size=2, stride=2))
      elif pooling == 'avg':
        layers.append(AvgPool2d(kernel_size=factor, stride=factor))
      elif pooling == 'max':
        layers.append(MaxPool2d(kernel_size=factor, stride=factor))
      # any other pooling string: the reference leaves `pool` unbound (layers.py:196-201)
    elif s[:2] == 'FC':
      _, Din, Dout = s.split('-')
      Din, Dout = int(Din), int(Dout)
      if not flat:
        layers.append(Flatten())
      flat = True
      layers.append(Linear(Din, Dout))
      if i + 1 < len(arch):
        layers.append(get_activation(activation))
      cur_C = Dout
    else:
      raise ValueError('Invalid layer "%s"' % s)
  layers = [layer for layer in layers if layer is not None]
  for layer in layers:
    print(layer)                                   # the reference prints them too (:211-212)
  return FusedSequential(*layers), cur_C


def build_mlp(dim_list, activation='relu', batch_norm='none',
              dropout=0, final_nonlinearity=True):
  """sg2im/layers.py:216-232."""
  layers = []
  for i in range(len(dim_list) - 1):
    dim_in, dim_out = dim_list[i], dim_list[i + 1]
    layers.append(Linear(dim_in, dim_out))
    final_layer = (i == len(dim_list) - 2)
    if not final_layer or final_nonlinearity:
      if batch_norm == 'batch':
        layers.append(BatchNorm1d(dim_out))
      if activation == 'relu':
        layers.append(ReLU())
      elif activation == 'leakyrelu':
        layers.append(LeakyReLU())
    if dropout > 0:
      layers.append(nn.Dropout(p=dropout))
  return FusedSequential(*layers)
