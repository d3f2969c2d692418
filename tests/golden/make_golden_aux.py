"""Golden fixtures for the rows next to the training step:
de-normalisation to uint8, box IoU, and the validation pass ``check_model`` —
all produced by the UNMODIFIED reference imported from /root/reference
(build container only).  Kept apart from make_golden.py so the training-step
fixtures are not rewritten.

  python tests/golden/make_golden_aux.py
"""
import argparse
import contextlib
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from refimport import import_reference, REF  # noqa: E402
from sg2im_b200.synth import make_vocab, synth_batch  # noqa: E402

SMALL_G = dict(embedding_dim=16, gconv_dim=16, gconv_hidden_dim=32, gconv_num_layers=2,
               refinement_dims=(32, 16), mask_size=8, layout_noise_dim=0,
               normalization='batch', activation='leakyrelu-0.2')


def main():
  assert import_reference() is not None, 'reference tree not found'
  sys.path.insert(0, os.path.join(REF, 'scripts'))
  from sg2im.data.utils import imagenet_deprocess_batch
  from sg2im.metrics import jaccard
  from sg2im.model import Sg2ImModel
  import train as ref_train

  out = {}
  # ---- 1. imagenet_deprocess_batch (sg2im/data/utils.py:48-67)
  g = torch.Generator().manual_seed(21)
  imgs = torch.randn(5, 3, 12, 10, generator=g) * 1.3
  imgs[1] *= 4.0                                      # saturates without rescale
  imgs[2, :, :, :] = imgs[2, :, :1, :1] + 1e-3 * imgs[2]   # tiny dynamic range
  out['deprocess'] = dict(imgs=imgs, rescaled=imagenet_deprocess_batch(imgs),
                          plain=imagenet_deprocess_batch(imgs, rescale=False))

  # ---- 2. jaccard (sg2im/metrics.py:27-35)
  a = torch.rand(9, 2, generator=g) * 0.6
  pred = torch.cat([a, a + torch.rand(9, 2, generator=g) * 0.4], dim=1)
  b = torch.rand(9, 2, generator=g) * 0.6
  gt = torch.cat([b, b + torch.rand(9, 2, generator=g) * 0.4], dim=1)
  gt[0] = pred[0]                                      # IoU 1
  gt[1] = torch.tensor([0.9, 0.9, 1.0, 1.0]); pred[1] = torch.tensor([0.0, 0.0, 0.1, 0.1])  # IoU 0
  out['jaccard'] = dict(pred=pred, gt=gt, value=jaccard(pred, gt))

  # ---- 3. check_model (scripts/train.py:309-384), VG-style and COCO-style batches
  vocab = make_vocab(9, 5)
  args = argparse.Namespace(l1_pixel_loss_weight=1.0, bbox_pred_loss_weight=10.0,
                            predicate_pred_loss_weight=0.5, mask_loss_weight=0.1,
                            num_val_samples=6)
  for name, with_masks in (('check_vg', False), ('check_coco', True)):
    torch.manual_seed(8)
    with contextlib.redirect_stdout(io.StringIO()):
      model = Sg2ImModel(vocab=vocab, image_size=(32, 32), **SMALL_G)
    with torch.no_grad():
      model.box_net[2].bias.copy_(torch.tensor([0.1, 0.15, 0.6, 0.7]))   # finite predicted boxes
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    loader = [synth_batch(N=4, objs_per_img=3, rels_per_img=2, image_size=(32, 32), num_objs=9,
                          num_preds=5, masks=with_masks, mask_size=8, seed=30 + i)
              for i in range(3)]                       # 4 + 4 >= 6: the third batch is never read
    model.train()                                      # train.py:511 then :613
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self     # check_model hard-codes .cuda()
    try:
      mean_losses, samples, batch_data, avg_iou = ref_train.check_model(args, 0, loader, model)
    finally:
      torch.Tensor.cuda = saved
    sd1 = {k: v.detach().clone() for k, v in model.state_dict().items() if 'running' in k or
           'num_batches' in k}
    out[name] = dict(kwargs=dict(SMALL_G, image_size=(32, 32)), vocab=vocab, args=vars(args),
                     sd=sd0, loader=loader, mean_losses={k: float(v) for k, v in mean_losses.items()},
                     samples=samples, batch_data=batch_data, avg_iou=avg_iou.detach().clone(),
                     bn_after=sd1)
  # ---- 4. layout / crop under the torch-0.4 sampling convention (align_corners=True, what
  # the published checkpoints were trained with): the reference's own functions with
  # F.grid_sample's default flipped, on the inputs of layout.pt / crop.pt
  import functools
  import torch.nn.functional as F
  from sg2im.layout import masks_to_layout, boxes_to_layout
  from sg2im.bilinear import crop_bbox_batch
  lay = torch.load(os.path.join(HERE, 'layout.pt'))
  crp = torch.load(os.path.join(HERE, 'crop.pt'))
  stock = F.grid_sample
  F.grid_sample = functools.partial(stock, align_corners=True)
  try:
    with contextlib.redirect_stdout(io.StringIO()):
      out['align_corners'] = dict(
          masks=masks_to_layout(lay['rvecs'], lay['rboxes'], lay['rmasks'], lay['robj_to_img'], 24, 40),
          boxes=boxes_to_layout(lay['vecs'], lay['boxes'], lay['obj_to_img'], 24, 20),
          crops=crop_bbox_batch(crp['feats'], crp['boxes'], crp['bbox_to_feats'], 6, 7))
  finally:
    F.grid_sample = stock
  path = os.path.join(HERE, 'aux.pt')
  torch.save(out, path)
  print('aux.pt %.1f KB' % (os.path.getsize(path) / 1024.0))
  for k in ('check_vg', 'check_coco'):
    print(k, out[k]['mean_losses'], float(out[k]['avg_iou']))


if __name__ == '__main__':
  main()
