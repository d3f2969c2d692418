"""The validation pass of the training script — reference surface
``check_model`` (scripts/train.py:309-384): the generator's losses and box IoU
over ``num_val_samples`` images, then three sample renderings of the last batch
(ground-truth boxes + masks, ground-truth boxes + predicted masks, everything
predicted), de-normalised to uint8.

Same calls in the same order as the reference (the model is left in whatever
train/eval mode the caller set — train.py calls this in train mode), minus its
host churn: losses stay on the device until one read-back at the end, images are
de-normalised on the device and only bytes cross PCIe.
"""
from collections import defaultdict

import torch
import torch.nn.functional as F

from .images import imagenet_deprocess_batch
from .metrics import jaccard

_LOSS_DEFAULTS = dict(l1_pixel_loss_weight=1.0, bbox_pred_loss_weight=10.0,
                      predicate_pred_loss_weight=0.0, mask_loss_weight=0.0,
                      num_val_samples=1024)


def _arg(args, name):
  if isinstance(args, dict):
    return args.get(name, _LOSS_DEFAULTS[name])
  return getattr(args, name, _LOSS_DEFAULTS[name])


def calculate_model_losses(args, skip_pixel_loss, model, img, img_pred, bbox, bbox_pred, masks,
                           masks_pred, predicates, predicate_scores):
  """scripts/train.py:387-412, values kept as 0-dim device tensors (the
  reference stores ``.item()`` floats: one host sync per loss term)."""
  losses = {}
  total = torch.zeros(1, dtype=img.dtype, device=img.device)

  def add(name, value, weight):
    nonlocal total
    value = value * weight
    losses[name] = value.detach()
    total = total + value

  add('L1_pixel_loss', F.l1_loss(img_pred, img),
      0 if skip_pixel_loss else _arg(args, 'l1_pixel_loss_weight'))
  add('bbox_pred', F.mse_loss(bbox_pred, bbox), _arg(args, 'bbox_pred_loss_weight'))
  if _arg(args, 'predicate_pred_loss_weight') > 0:
    add('predicate_pred', F.cross_entropy(predicate_scores, predicates),
        _arg(args, 'predicate_pred_loss_weight'))
  if _arg(args, 'mask_loss_weight') > 0 and masks is not None and masks_pred is not None:
    add('mask_loss', F.binary_cross_entropy(masks_pred, masks.float()),
        _arg(args, 'mask_loss_weight'))
  return total, losses


def check_model(args, t, loader, model, device=None, deprocess=imagenet_deprocess_batch):
  """Returns ``(mean_losses, samples, batch_data, avg_iou)`` like the reference:
  dict of float means, dict of uint8 (N,3,H,W) CPU image batches, dict of CPU
  copies of the last batch, and the mean box IoU (0-dim CPU tensor)."""
  if device is None:
    device = next(model.parameters()).device
  per_loss = defaultdict(list)
  iou_sum, n_boxes, n_imgs = None, 0, 0
  batch = None
  with torch.no_grad():
    for cpu_batch in loader:
      batch = [x.to(device, non_blocking=True) for x in cpu_batch]
      masks = None
      if len(batch) == 6:
        imgs, objs, boxes, triples, obj_to_img, triple_to_img = batch
      elif len(batch) == 7:
        imgs, objs, boxes, masks, triples, obj_to_img, triple_to_img = batch
      else:
        raise ValueError('check_model: batches are 6- or 7-tuples (vg / coco collate)')
      imgs_pred, boxes_pred, masks_pred, predicate_scores = model(
          objs, triples, obj_to_img, boxes_gt=boxes, masks_gt=masks)
      _, losses = calculate_model_losses(args, False, model, imgs, imgs_pred, boxes, boxes_pred,
                                         masks, masks_pred, triples[:, 1], predicate_scores)
      iou = jaccard(boxes_pred, boxes)
      iou_sum = iou if iou_sum is None else iou_sum + iou
      n_boxes += boxes_pred.size(0)
      for name, value in losses.items():
        per_loss[name].append(value.reshape(()))
      n_imgs += imgs.size(0)
      if n_imgs >= _arg(args, 'num_val_samples'):
        break
    if batch is None:
      raise ValueError('check_model: empty loader')

    renders = {'gt_img': imgs}
    renders['gt_box_gt_mask'] = model(objs, triples, obj_to_img, boxes_gt=boxes, masks_gt=masks)[0]
    renders['gt_box_pred_mask'] = model(objs, triples, obj_to_img, boxes_gt=boxes)[0]
    renders['pred_box_pred_mask'] = model(objs, triples, obj_to_img)[0]
    samples = {k: deprocess(v) for k, v in renders.items()}

    # one read-back for every scalar
    names = list(per_loss)
    means = torch.stack([torch.stack(per_loss[k]).double().mean() for k in names]).cpu().tolist()
    mean_losses = dict(zip(names, means))
    avg_iou = (iou_sum / n_boxes).cpu()

  def host(x):
    return None if x is None else x.detach().cpu().clone()

  batch_data = {
    'objs': host(objs), 'boxes_gt': host(boxes), 'masks_gt': host(masks),
    'triples': host(triples), 'obj_to_img': host(obj_to_img),
    'triple_to_img': host(triple_to_img), 'boxes_pred': host(boxes_pred),
    'masks_pred': host(masks_pred),
  }
  return mean_losses, samples, batch_data, avg_iou
