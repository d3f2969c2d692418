"""TEST INFRASTRUCTURE ONLY: a CPU stand-in for the op boundary of
``sg2im_b200.ops`` (the autograd Functions / wrappers that launch the CUDA
kernels), built from torch CPU primitives and the oracle.

Purpose: ``-m "not gpu"`` tests of the HOST LOGIC — that the nn.Module mirror
(`model`, `crn`, `graph`, `layers`, `discriminators`, `train_step`) wires the ops
together exactly like the reference (which slices, which order, which buffers,
which BatchNorm sees what) — by replaying the golden fixtures with the kernels
swapped for their mathematical definition.  It is installed with
``with cpu_ops(): ...`` (monkeypatching, restored on exit) and is never
importable from the product; the product path has no CPU fallback.
"""
import contextlib

import torch
import torch.nn.functional as F

from oracle import sg2im_oracle as orc


def _leaky(x, slope):
  return x if slope == 1.0 else F.leaky_relu(x, slope)


def _conv2d(x, weight, bias, stride=1, pad=0, act=0, slope=0.0, in_ch=None, feeds_bn=False,
            stats_out=None, round_out=False):
  w = weight if in_ch is None else weight[:, :in_ch]
  y = F.conv2d(x.permute(0, 3, 1, 2), w, bias, stride=stride, padding=pad)
  if stats_out is not None:
    with torch.no_grad():
      C = y.size(1)
      stats_out[:C] += y.double().sum(dim=(0, 2, 3))
      stats_out[C:] += (y.double() ** 2).sum(dim=(0, 2, 3))
  if act:
    y = F.leaky_relu(y, slope)
  return y.permute(0, 2, 3, 1)


def _linear(x2d, weight, bias, act=0, slope=0.0, round_out=False):
  y = F.linear(x2d, weight, bias)
  return F.leaky_relu(y, slope) if act else y


def _new_stats(channels, device):
  return torch.zeros(2 * channels, dtype=torch.float64, device=device)


def _bn_act(x, bn=None, slope=1.0, up=1, unbias_mult=1, out=None, out_coff=0, sums=None):
  """x NHWC.  Same contract as ops.bn_act: BN (train: batch statistics + running
  update) -> leaky(slope) -> nearest upsample, optionally written behind
  out[..., :out_coff]."""
  h = x.permute(0, 3, 1, 2)
  bn_after_up = bn is not None and up > 1 and unbias_mult == up * up
  if bn_after_up:                                   # mask head order: Upsample -> BN
    h = F.interpolate(h, scale_factor=up, mode='nearest')
  if bn is not None:
    training = bn.training or bn.running_mean is None
    if training and bn.num_batches_tracked is not None:
      bn.num_batches_tracked.add_(1)
    if sums is not None and training and not bn_after_up:
      # the statistics handed over by the conv must describe this very tensor
      C = h.size(1)
      cnt = h.numel() // C
      mean = (sums[:C] / cnt).float()
      assert torch.allclose(mean, h.mean(dim=(0, 2, 3)), rtol=1e-4, atol=1e-5)
    h = F.batch_norm(h, bn.running_mean, bn.running_var, bn.weight, bn.bias, training,
                     0.1 if bn.momentum is None else bn.momentum, bn.eps)
  h = _leaky(h, slope)
  if up > 1 and not bn_after_up:
    h = F.interpolate(h, scale_factor=up, mode='nearest')
  y = h.permute(0, 2, 3, 1)
  if out is None:
    return y
  C = y.size(3)
  return torch.cat([out[..., :out_coff], y, out[..., out_coff + C:]], dim=3)


class _Apply(object):
  """Object with an .apply like an autograd Function."""

  def __init__(self, fn):
    self.apply = fn


def _layout(vecs, boxes, masks, obj_to_img, N, H, W, noise, align_corners):
  assert not align_corners
  if masks is None:
    lay = orc.boxes_to_layout(vecs, boxes, obj_to_img, H, W, N)
  else:
    lay = orc.masks_to_layout(vecs, boxes, masks.float(), obj_to_img, H, W, N)
  if noise is not None:
    lay = torch.cat([lay, noise], dim=1)
  return lay.permute(0, 2, 3, 1)


def _layout_stack(vecs, boxes, masks, obj_to_img, N, H, W, noise, align_corners, extras):
  lay = _layout(vecs, boxes, masks, obj_to_img, N, H, W, noise, align_corners).permute(0, 3, 1, 2)
  L = len(extras)
  bufs = [None] * L
  cur = lay
  for k in range(L - 1, -1, -1):
    if k < L - 1:
      cur = F.avg_pool2d(cur, 2, 2)                 # factor-2 cascade, as the kernels do
    pad = torch.zeros(cur.size(0), extras[k], cur.size(2), cur.size(3))
    bufs[k] = torch.cat([cur, pad], dim=1).permute(0, 2, 3, 1)
  return tuple(bufs)


def _crop(feats, boxes, idx, HH, WW, align_corners):
  assert not align_corners
  return orc.crop_bbox_batch(feats.permute(0, 3, 1, 2), boxes, idx, HH, WW).permute(0, 2, 3, 1)


def _triple_gather(obj_vecs, pred_vecs, edges, csr):
  return torch.cat([obj_vecs[edges[:, 0]], pred_vecs, obj_vecs[edges[:, 1]]], dim=1)


def _graph_pool(new_t, edges, csr, H, Dout, num_objs, avg):
  pooled = orc.graph_pool(new_t, edges, num_objs, H, Dout, 'avg' if avg else 'sum')
  return pooled, new_t[:, H:H + Dout]


def _avgpool2_fwd(x, x_coff, C, out, out_coff):
  src = x[..., x_coff:x_coff + C].permute(0, 3, 1, 2)
  out[..., out_coff:out_coff + C] = F.avg_pool2d(src, 2, 2).permute(0, 2, 3, 1)


def _avgpool2_bwd(dcoarse, dc_coff, C, dfine, df_coff, accumulate):
  g = dcoarse[..., dc_coff:dc_coff + C] * 0.25
  g = g.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
  if accumulate:
    dfine[..., df_coff:df_coff + C] += g
  else:
    dfine[..., df_coff:df_coff + C] = g


def _pool2d(h, factor, mode):
  x = h.permute(0, 3, 1, 2)
  y = F.avg_pool2d(x, factor, factor) if mode == 0 else F.max_pool2d(x, factor, factor)
  return y.permute(0, 2, 3, 1)


def _round_tf32(src, dst):
  u = src.view(torch.int32)
  finite = (u & 0x7f800000) != 0x7f800000
  dst.copy_(torch.where(finite, (u + 0x1000) & ~0x1fff, u & ~0x1fff).view(torch.float32))


def _adam_flat(params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, weight_decay=0.0,
               found_inf=None, shadow=None, grad_scale=1.0):
  """Mathematical definition of sg2im_adam_flat (csrc/adam.cu), torch/optim/adam.py
  single-tensor arithmetic, in place."""
  if found_inf is not None and float(found_inf) != 0.0:
    return
  with torch.no_grad():
    step += 1
    t = float(step)
    g = grads * grad_scale
    g = g if weight_decay == 0 else g + weight_decay * params
    exp_avg.add_((g - exp_avg) * (1 - beta1))
    exp_avg_sq.mul_(beta2).add_((1 - beta2) * g * g)
    bc1 = 1 - beta1 ** t
    bc2_sqrt = (1 - beta2 ** t) ** 0.5
    denom = exp_avg_sq.sqrt() / bc2_sqrt + eps
    params.sub_((lr / bc1) * (exp_avg / denom))
    if shadow is not None:
      _round_tf32(params, shadow)


@contextlib.contextmanager
def cpu_ops():
  from sg2im_b200 import ops
  saved = {}
  repl = dict(conv2d=_conv2d, linear=_linear, bn_act=_bn_act, new_stats=_new_stats,
              adam_flat=_adam_flat, round_tf32=_round_tf32, avgpool2_fwd=_avgpool2_fwd, avgpool2_bwd=_avgpool2_bwd,
              csr_build=lambda idx, nroles, num_rows: (None, None),
              LayoutStack=_Apply(_layout_stack), Layout=_Apply(_layout), Crop=_Apply(_crop),
              TripleGather=_Apply(_triple_gather), GraphPool=_Apply(_graph_pool),
              Pool2d=_Apply(_pool2d))
  for k, v in repl.items():
    saved[k] = getattr(ops, k)
    setattr(ops, k, v)
  try:
    yield
  finally:
    for k, v in saved.items():
      setattr(ops, k, v)
