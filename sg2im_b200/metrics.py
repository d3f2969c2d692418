"""Box metrics of the validation step (reference surface: sg2im/metrics.py:20-35).
O(objects) scalar work: stays in PyTorch, like the losses."""
import torch


def intersection(bbox_pred, bbox_gt):
  """Overlap area of matching rows of two (O, 4) xyxy box lists."""
  lo = torch.maximum(bbox_pred[:, :2], bbox_gt[:, :2])
  hi = torch.minimum(bbox_pred[:, 2:], bbox_gt[:, 2:])
  wh = (hi - lo).clamp_(min=0)
  return wh[:, 0] * wh[:, 1]


def _area(b):
  return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])


def jaccard(bbox_pred, bbox_gt):
  """SUM over rows of intersection-over-union (the caller divides by the number
  of boxes, scripts/train.py:343-344,369)."""
  inter = intersection(bbox_pred, bbox_gt)
  union = _area(bbox_pred) + _area(bbox_gt) - inter
  return (inter / union).sum()
