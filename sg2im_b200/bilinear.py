"""Differentiable box crops — reference surface of sg2im/bilinear.py's default
path (crop_bbox_batch -> crop_bbox_batch_cudnn -> crop_bbox) as one gather
kernel: no per-image Python loop, no nonzero() host syncs, any object order."""
from . import ops
from . import layout as _layout


def crop_bbox_batch(feats, bbox, bbox_to_feats, HH, WW=None, backend='cudnn'):
  """feats (N,C,H,W), bbox (B,4) xyxy in [0,1], bbox_to_feats (B,) ->
  crops (B,C,HH,WW) with crops[i] cut from feats[bbox_to_feats[i]]
  (sg2im/bilinear.py:28-43)."""
  if backend != 'cudnn':
    raise NotImplementedError("sg2im_b200: only the default backend ('cudnn' semantics)")
  if WW is None:
    WW = HH
  out = ops.Crop.apply(feats.permute(0, 2, 3, 1), bbox, bbox_to_feats, HH, WW,
                       _layout.ALIGN_CORNERS)
  return out.permute(0, 3, 1, 2)


def crop_bbox(feats, bbox, HH, WW=None, backend='cudnn'):
  """sg2im/bilinear.py:103-132: crop i from feats[i]."""
  import torch
  idx = torch.arange(feats.size(0), device=feats.device, dtype=torch.int64)
  return crop_bbox_batch(feats, bbox, idx, HH, WW, backend)


def crop_bbox_batch_cudnn(feats, bbox, bbox_to_feats, HH, WW=None):
  """sg2im/bilinear.py:69-100 — the reference groups the boxes by image, crops per image and
  inverse-permutes; the gather kernel needs none of that, so this is crop_bbox_batch."""
  return crop_bbox_batch(feats, bbox, bbox_to_feats, HH, WW)


def tensor_linspace(start, end, steps=10):
  """sg2im/bilinear.py:249-278: out[..., i] = w0[i] * start + w1[i] * end with
  w0 = linspace(1, 0, steps), w1 = linspace(0, 1, steps); result shape start.shape + (steps,).
  Host helper (tiny tensors); the crop kernel evaluates the same expression per sample."""
  import torch
  assert start.size() == end.size()
  w0 = torch.linspace(1, 0, steps=steps).to(start)
  w1 = torch.linspace(0, 1, steps=steps).to(start)
  return w0 * start.unsqueeze(-1) + w1 * end.unsqueeze(-1)
