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
