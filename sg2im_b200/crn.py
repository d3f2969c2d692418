"""Cascaded refinement network — reference surface of sg2im/crn.py.

Same module tree / state_dict keys; the forward is re-planned around the
kernels: every stage's input buffer (pooled layout ++ upsampled features) is
allocated once by ops.LayoutStack, the previous stage's
BatchNorm -> LeakyReLU -> nearest-x2 upsample is one pass that writes straight
into that buffer's trailing channels (no torch.cat, no F.upsample, no separate
BN/activation passes), and the first stage skips its all-zero extra input
channel."""
import torch
import torch.nn as nn

from . import ops
from .layers import (get_normalization_2d, get_activation, Conv2d, BatchNorm2d, InstanceNorm2d,
                     FusedSequential, norm_act, _to_nhwc, _to_nchw)


# Inference: fold an eval-mode BatchNorm into the convolution before it
# (w' = w * g/sqrt(var+eps), b' = (b - mean) * g/sqrt(var+eps) + beta) so that
# conv -> BN -> LeakyReLU is ONE kernel (activation in the conv epilogue) and the
# only elementwise pass left per stage is the x2 upsample into the next stage's
# buffer.  Off until it has been timed and parity-checked on hardware
# (tests/test_gpu_next_rows.py); the wiring is covered on CPU.
FOLD_EVAL_BN = False


def _can_fold(bn):
  return (FOLD_EVAL_BN and isinstance(bn, BatchNorm2d) and not bn.training
          and bn.running_mean is not None)


def _folded(conv, bn):
  scale = torch.rsqrt(bn.running_var + bn.eps)
  if bn.weight is not None:
    scale = scale * bn.weight
  w = conv.weight * scale.view(-1, 1, 1, 1)
  b = -bn.running_mean if conv.bias is None else conv.bias - bn.running_mean
  b = b * scale
  if bn.bias is not None:
    b = b + bn.bias
  return w, b


def _conv_bn_act(conv, bn, slope, h, in_ch, round_out):
  """act(BN_eval(conv(h))) as one convolution with folded parameters."""
  stride, pad = conv._cfg()
  w, b = _folded(conv, bn)
  return ops.conv2d(h, w, b, stride, pad, 1, slope, in_ch, round_out=round_out)


class RefinementModule(nn.Module):
  """sg2im/crn.py:35-65."""

  def __init__(self, layout_dim, input_dim, output_dim,
               normalization='instance', activation='leakyrelu'):
    super(RefinementModule, self).__init__()
    layers = []
    layers.append(Conv2d(layout_dim + input_dim, output_dim, kernel_size=3, padding=1))
    layers.append(get_normalization_2d(output_dim, normalization))
    layers.append(get_activation(activation))
    layers.append(Conv2d(output_dim, output_dim, kernel_size=3, padding=1))
    layers.append(get_normalization_2d(output_dim, normalization))
    layers.append(get_activation(activation))
    layers = [layer for layer in layers if layer is not None]
    for layer in layers:
      if isinstance(layer, nn.Conv2d):
        nn.init.kaiming_normal_(layer.weight)
    self.net = FusedSequential(*layers)
    self.layout_dim = layout_dim
    self.input_dim = input_dim

  def parts(self):
    """(conv1, norm1|None, slope1, conv2, norm2|None, slope2); norm = BatchNorm2d or
    InstanceNorm2d."""
    mods = list(self.net)
    convs = [m for m in mods if isinstance(m, Conv2d)]
    bns = [m for m in mods if isinstance(m, (BatchNorm2d, InstanceNorm2d))] or [None, None]
    acts = [m for m in mods if isinstance(m, nn.LeakyReLU)]
    return convs[0], bns[0], acts[0].negative_slope, convs[1], bns[1], acts[1].negative_slope

  def forward(self, layout, feats):
    """Stand-alone module call with the reference's semantics (crn.py:54-65);
    RefinementNetwork.forward does not go through here."""
    _, _, HH, WW = layout.size()
    _, _, H, W = feats.size()
    assert HH >= H
    if HH > H:
      factor = round(HH // H)
      assert HH % factor == 0
      assert WW % factor == 0 and WW // factor == W
      lay = _to_nhwc(layout).contiguous()
      while lay.size(1) > H:
        nxt = torch.empty(lay.size(0), lay.size(1) // 2, lay.size(2) // 2, lay.size(3),
                          dtype=lay.dtype, device=lay.device)
        ops.avgpool2_fwd(lay, 0, lay.size(3), nxt, 0)      # inference-only helper
        lay = nxt
      layout = _to_nchw(lay)
    net_input = torch.cat([layout, feats], dim=1)
    return self.net(net_input)


class RefinementNetwork(nn.Module):
  """sg2im/crn.py:68-111."""

  def __init__(self, dims, normalization='instance', activation='leakyrelu'):
    super(RefinementNetwork, self).__init__()
    layout_dim = dims[0]
    self.refinement_modules = nn.ModuleList()
    for i in range(1, len(dims)):
      input_dim = 1 if i == 1 else dims[i - 1]
      output_dim = dims[i]
      mod = RefinementModule(layout_dim, input_dim, output_dim,
                             normalization=normalization, activation=activation)
      self.refinement_modules.append(mod)
    output_conv_layers = [
      Conv2d(dims[-1], dims[-1], kernel_size=3, padding=1),
      get_activation(activation),
      Conv2d(dims[-1], 3, kernel_size=1, padding=0)
    ]
    nn.init.kaiming_normal_(output_conv_layers[0].weight)
    nn.init.kaiming_normal_(output_conv_layers[2].weight)
    self.output_conv = FusedSequential(*output_conv_layers)
    self.layout_dim = layout_dim

  def stage_extras(self):
    """Trailing feature channels of each stage's input buffer: stage 0 takes
    none (its 1-channel zero input is skipped), stage i takes dims[i-1]."""
    return [0] + [m.input_dim for m in list(self.refinement_modules)[1:]]

  def forward_stack(self, bufs):
    """bufs: per-stage NHWC input buffers from ops.LayoutStack (layout slice
    filled, feature slice pending).  Returns the image, NHWC."""
    mods = list(self.refinement_modules)
    C = self.layout_dim
    h = bufs[0]
    a = None
    for i, mod in enumerate(mods):
      conv1, bn1, s1, conv2, bn2, s2 = mod.parts()
      if _can_fold(bn1) and _can_fold(bn2):
        last = i + 1 == len(mods)
        a1 = _conv_bn_act(conv1, bn1, s1, h, C if i == 0 else None, True)
        a2 = _conv_bn_act(conv2, bn2, s2, a1, None, last)
        if last:
          a = a2
        else:
          h = ops.bn_act(a2, None, 1.0, up=2, out=bufs[i + 1], out_coff=C)
        continue
      fb1 = isinstance(bn1, BatchNorm2d) and bn1.training
      fb2 = isinstance(bn2, BatchNorm2d) and bn2.training
      st1 = ops.new_stats(conv1.out_channels, h.device) if fb1 else None
      z1 = conv1.forward_nhwc(h, in_ch=C if i == 0 else None, feeds_bn=fb1, stats_out=st1)
      a1 = norm_act(z1, bn1, s1, sums=st1)
      st2 = ops.new_stats(conv2.out_channels, h.device) if fb2 else None
      z2 = conv2.forward_nhwc(a1, feeds_bn=fb2, stats_out=st2)
      if i + 1 < len(mods):
        # BN + LeakyReLU + nearest x2 upsample, written into the next stage's
        # buffer behind its layout channels (crn.py:107 + :63)
        h = norm_act(z2, bn2, s2, up=2, out=bufs[i + 1], out_coff=C, sums=st2)
      else:
        a = norm_act(z2, bn2, s2, sums=st2)
    oc = list(self.output_conv)
    t = oc[0].forward_nhwc(a, 1, oc[1].negative_slope)
    return oc[2].forward_nhwc(t)

  def forward(self, layout):
    """Reference call shape: layout (N, C, H, W) -> image (N, 3, H, W).  Builds
    the stage buffers from an existing layout tensor (one copy); the model
    uses forward_stack directly and never materialises the layout twice."""
    N, C, H, W = layout.size()
    L = len(self.refinement_modules)
    assert (H >> L) != 0 and (W >> L) != 0
    bufs = _StackFromLayout.apply(_to_nhwc(layout).contiguous(), tuple(self.stage_extras()))
    return _to_nchw(self.forward_stack(list(bufs)))


class _StackFromLayout(torch.autograd.Function):
  """Stage buffers from an already computed NHWC layout (stand-alone
  RefinementNetwork calls)."""

  @staticmethod
  def forward(ctx, layout, extras):
    N, H, W, C = layout.shape
    L = len(extras)
    bufs = [None] * L
    bufs[L - 1] = torch.empty(N, H, W, C + extras[L - 1], dtype=layout.dtype,
                              device=layout.device)
    bufs[L - 1][..., :C] = layout
    for k in range(L - 2, -1, -1):
      f = L - 1 - k
      bufs[k] = torch.empty(N, H >> f, W >> f, C + extras[k], dtype=layout.dtype,
                            device=layout.device)
      ops.avgpool2_fwd(bufs[k + 1], 0, C, bufs[k], 0)
    ctx.C = C
    return tuple(bufs)

  @staticmethod
  def backward(ctx, *douts):
    C = ctx.C
    g = [d.contiguous() for d in douts]
    for k in range(1, len(g)):
      ops.avgpool2_bwd(g[k - 1], 0, C, g[k], 0, True)
    return g[-1][..., :C].contiguous(), None
