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


def boxes_to_layout(vecs, boxes, obj_to_img, H, W=None, pooling='sum', num_imgs=None):
  """sg2im/layout.py:30-63."""
  if pooling != 'sum':
    raise NotImplementedError("sg2im_b200: only pooling='sum' (what the model uses)")
  if W is None:
    W = H
  N = _num_imgs(obj_to_img, num_imgs)
  out = ops.Layout.apply(vecs, boxes, None, obj_to_img, N, H, W, None, ALIGN_CORNERS)
  return out.permute(0, 3, 1, 2)


def masks_to_layout(vecs, boxes, masks, obj_to_img, H, W=None, pooling='sum', num_imgs=None):
  """sg2im/layout.py:66-91."""
  if pooling != 'sum':
    raise NotImplementedError("sg2im_b200: only pooling='sum' (what the model uses)")
  O, D = vecs.size()
  M = masks.size(1)
  assert masks.size() == (O, M, M)
  if W is None:
    W = H
  N = _num_imgs(obj_to_img, num_imgs)
  out = ops.Layout.apply(vecs, boxes, masks, obj_to_img, N, H, W, None, ALIGN_CORNERS)
  return out.permute(0, 3, 1, 2)
