"""Scene layout from object vectors, boxes and masks — reference surface of
sg2im/layout.py on the fused sm_100a layout kernel (no (O,D,H,W) temporary,
no host sync).  Returned tensors have the reference's (N, D, H, W) shape."""
import torch

from . import ops

# torch >= 1.3 semantics of the reference's bare F.grid_sample calls
# (SURVEY.md §0.5).  Set True to reproduce checkpoints trained under torch 0.4.
ALIGN_CORNERS = False


def _num_imgs(obj_to_img, num_imgs):
  if num_imgs is not None:
    return int(num_imgs)
  # the reference's own way (layout.py:143): one device->host sync.  Callers on
  # the training path pass num_imgs (= imgs.size(0)) and never reach this.
  return int(obj_to_img.max().item()) + 1


def _pooling_weights(vecs, obj_to_img, N, pooling):
  """_pool_samples (sg2im/layout.py:131-162): 'sum', or 'avg' = the per-image sum divided by
  clamp(number of objects of the image, 1).  The layout is linear in the object vectors, so the
  division is applied to the (O, D) vectors (each by its image's count) instead of to the
  (N, D, H, W) result: same value up to the last rounding, no extra pass over the layout.  The
  reference prints the counts in 'avg' mode (:156); so does this."""
  if pooling == 'sum':
    return vecs
  if pooling != 'avg':
    raise ValueError('Invalid pooling "%s"' % pooling)
  counts = torch.bincount(obj_to_img, minlength=N).to(vecs.dtype)
  print(counts)
  return vecs / counts.clamp(min=1)[obj_to_img].unsqueeze(1)


def boxes_to_layout(vecs, boxes, obj_to_img, H, W=None, pooling='sum', num_imgs=None):
  """sg2im/layout.py:30-63."""
  if W is None:
    W = H
  N = _num_imgs(obj_to_img, num_imgs)
  vecs = _pooling_weights(vecs, obj_to_img, N, pooling)
  out = ops.Layout.apply(vecs, boxes, None, obj_to_img, N, H, W, None, ALIGN_CORNERS)
  return out.permute(0, 3, 1, 2)


def masks_to_layout(vecs, boxes, masks, obj_to_img, H, W=None, pooling='sum', num_imgs=None):
  """sg2im/layout.py:66-91."""
  O, D = vecs.size()
  M = masks.size(1)
  assert masks.size() == (O, M, M)
  if W is None:
    W = H
  N = _num_imgs(obj_to_img, num_imgs)
  vecs = _pooling_weights(vecs, obj_to_img, N, pooling)
  out = ops.Layout.apply(vecs, boxes, masks, obj_to_img, N, H, W, None, ALIGN_CORNERS)
  return out.permute(0, 3, 1, 2)
