"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference (imported from /root/reference; build container only).

  python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these are
"outputs of the reference itself run here".  Every fixture stores the inputs,
the (small) state_dict and the reference outputs; model sizes are shrunk so the
fixtures stay small.  The oracle (oracle/sg2im_oracle.py) is replayed against
them by tests/test_oracle_golden.py.
"""
import copy
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from refimport import import_reference, REF  # noqa: E402
from sg2im_b200.synth import make_vocab, synth_batch  # noqa: E402

SMALL_G = dict(embedding_dim=16, gconv_dim=16, gconv_hidden_dim=32,
               gconv_num_layers=3, refinement_dims=(32, 16), mask_size=8,
               layout_noise_dim=4, normalization='batch',
               activation='leakyrelu-0.2')
SMALL_D_ARCH = 'C4-8-2,C4-16-2'


def clone_sd(m):
  return {k: v.detach().clone() for k, v in m.state_dict().items()}


def save(name, obj):
  path = os.path.join(HERE, name)
  torch.save(obj, path)
  print('%-22s %8.1f KB' % (name, os.path.getsize(path) / 1024.0))


def main():
  ref = import_reference()
  assert ref is not None, 'reference tree not found'
  from sg2im.graph import GraphTripleConv
  from sg2im.layout import masks_to_layout, boxes_to_layout
  from sg2im.bilinear import crop_bbox_batch
  from sg2im.model import Sg2ImModel
  from sg2im.discriminators import PatchDiscriminator, AcCropDiscriminator
  from sg2im.losses import gan_g_loss, gan_d_loss
  import contextlib, io

  # ---- 1. one scene-graph convolution layer (bit-exact pin for the pooling)
  torch.manual_seed(1)
  gc = GraphTripleConv(input_dim=8, output_dim=8, hidden_dim=16, pooling='avg')
  O, T = 7, 10
  obj_vecs, pred_vecs = torch.randn(O, 8), torch.randn(T, 8)
  edges = torch.tensor([[0, 1], [1, 2], [0, 2], [3, 0], [2, 2], [4, 0], [0, 4],
                        [1, 3], [3, 1], [4, 2]])          # objects 5, 6 unused
  new_obj, new_p = gc(obj_vecs, pred_vecs, edges)
  gc_sum = GraphTripleConv(input_dim=8, output_dim=8, hidden_dim=16, pooling='sum')
  gc_sum.load_state_dict(gc.state_dict())
  new_obj_sum, _ = gc_sum(obj_vecs, pred_vecs, edges)
  save('gconv.pt', dict(sd=clone_sd(gc), obj_vecs=obj_vecs, pred_vecs=pred_vecs,
                        edges=edges, new_obj=new_obj.detach(), new_p=new_p.detach(),
                        new_obj_sum=new_obj_sum.detach()))

  # ---- 2. layout: the literal inputs of the reference's own demo
  # (sg2im/layout.py:166-232), at 32x32 instead of 256x256
  vecs = torch.tensor([[1., 0, 0], [0, 1, 0], [0, 0, 1]] * 2)
  boxes = torch.tensor([[0.25, 0.125, 0.5, 0.875], [0, 0, 1, 0.25],
                        [0.6125, 0, 0.875, 1], [0, 0.8, 1, 1.0],
                        [0.25, 0.125, 0.5, 0.875], [0.6125, 0, 0.875, 1]])
  obj_to_img = torch.tensor([0, 0, 0, 1, 1, 1])
  diamond = torch.tensor([[0., 0, 1, 0, 0], [0, 1, 1, 1, 0], [1, 1, 1, 1, 1],
                          [0, 1, 1, 1, 0], [0, 0, 1, 0, 0]])
  ring = torch.tensor([[0., 0, 1, 0, 0], [0, 1, 0, 1, 0], [1, 0, 0, 0, 1],
                       [0, 1, 0, 1, 0], [0, 0, 1, 0, 0]])
  masks = torch.stack([diamond, ring, diamond, diamond, diamond, diamond])
  out_b = boxes_to_layout(vecs, boxes, obj_to_img, 32)
  out_m = masks_to_layout(vecs, boxes, masks, obj_to_img, 32)
  torch.manual_seed(2)
  rv, rb = torch.randn(9, 5), torch.rand(9, 2) * 0.6
  rb = torch.cat([rb, rb + torch.rand(9, 2) * 0.4 + 0.05], dim=1)
  rm = torch.rand(9, 16, 16)
  r2i = torch.tensor([0, 0, 1, 1, 1, 2, 2, 2, 2])
  out_r = masks_to_layout(rv, rb, rm, r2i, 24, 40)     # H != W
  save('layout.pt', dict(vecs=vecs, boxes=boxes, obj_to_img=obj_to_img, masks=masks,
                         out_boxes=out_b, out_masks=out_m, rvecs=rv, rboxes=rb,
                         rmasks=rm, robj_to_img=r2i, out_rand=out_r))

  # ---- 3. crops, objects NOT grouped by image (exercises the inverse permute)
  torch.manual_seed(3)
  feats = torch.randn(3, 3, 16, 20)
  cb = torch.tensor([[0.1, 0.2, 0.6, 0.9], [0., 0., 1., 1.], [0.3, 0.3, 0.5, 0.4],
                     [-0.1, 0.5, 0.7, 1.2], [0.45, 0.05, 0.95, 0.55]])
  b2f = torch.tensor([2, 0, 1, 0, 2])
  crops = crop_bbox_batch(feats, cb, b2f, 8)
  save('crop.pt', dict(feats=feats, boxes=cb, bbox_to_feats=b2f, crops=crops))

  # ---- 4. small generator, train-mode forward, VG mode and COCO mode
  vocab = make_vocab(9, 5)
  torch.manual_seed(4)
  with contextlib.redirect_stdout(io.StringIO()):
    model = Sg2ImModel(vocab=vocab, image_size=(32, 32), **SMALL_G)
  # random init gives degenerate predicted boxes (ReLU -> x1 == x0 -> 0/0 in
  # _boxes_to_grid, NaN in the reference too); shift the box head's bias so the
  # predicted-box path is exercised with finite values.
  with torch.no_grad():
    model.box_net[2].bias.copy_(torch.tensor([0.1, 0.15, 0.6, 0.7]))
  sd0 = clone_sd(model)
  batch = synth_batch(N=4, objs_per_img=3, rels_per_img=2, image_size=(32, 32),
                      num_objs=9, num_preds=5, seed=11)
  imgs, objs, boxes, triples, obj_to_img, _ = batch
  model.train()
  torch.manual_seed(1234)
  out_vg = model(objs, triples, obj_to_img, boxes_gt=boxes)
  sd_after_vg = clone_sd(model)
  model.load_state_dict(sd0)
  gt_masks = torch.randint(0, 2, (objs.size(0), 8, 8))
  torch.manual_seed(1234)
  out_coco = model(objs, triples, obj_to_img, boxes_gt=boxes, masks_gt=gt_masks)
  model.load_state_dict(sd0)
  model.eval()
  torch.manual_seed(1234)
  with torch.no_grad():
    out_eval = model(objs, triples, obj_to_img)            # predicted boxes+masks
  save('generator.pt', dict(
      kwargs=dict(SMALL_G, image_size=(32, 32)), vocab=vocab, sd=sd0, batch=batch,
      gt_masks=gt_masks, noise_seed=1234,
      out_vg=[t.detach() for t in out_vg], out_coco=[t.detach() for t in out_coco],
      out_eval=[t.detach() for t in out_eval],
      running_after_vg={k: v for k, v in sd_after_vg.items() if 'running' in k}))

  # ---- 4b. config 1: the sheep scene graphs through forward_json, 64x64, eval
  sheep = json.load(open(os.path.join(REF, 'scene_graphs', 'figure_6_sheep.json')))
  names = sorted({o for sg in sheep for o in sg['objects']})
  preds = sorted({r[1] for sg in sheep for r in sg['relationships']})
  obj_names = ['__image__'] + names
  pred_names = ['__in_image__'] + preds
  svocab = {'object_idx_to_name': obj_names,
            'object_name_to_idx': {n: i for i, n in enumerate(obj_names)},
            'pred_idx_to_name': pred_names,
            'pred_name_to_idx': {n: i for i, n in enumerate(pred_names)}}
  torch.manual_seed(5)
  kw = dict(SMALL_G, refinement_dims=(32, 16, 8), mask_size=16)
  with contextlib.redirect_stdout(io.StringIO()):
    smodel = Sg2ImModel(vocab=svocab, image_size=(64, 64), **kw)
  with torch.no_grad():
    smodel.box_net[2].bias.copy_(torch.tensor([0.1, 0.15, 0.6, 0.7]))
  smodel.eval()
  enc = smodel.encode_scene_graphs(copy.deepcopy(sheep))
  torch.manual_seed(77)
  with torch.no_grad():
    s_out = smodel.forward_json(copy.deepcopy(sheep))
  save('sheep.pt', dict(kwargs=dict(kw, image_size=(64, 64)), vocab=svocab,
                        scene_graphs=sheep, sd=clone_sd(smodel), encoded=enc,
                        noise_seed=77, out=[t.detach() for t in s_out]))

  # ---- 5. discriminators + GAN losses
  torch.manual_seed(6)
  with contextlib.redirect_stdout(io.StringIO()):
    d_img = PatchDiscriminator(arch=SMALL_D_ARCH, normalization='batch',
                               activation='leakyrelu-0.2', padding='valid')
    d_obj = AcCropDiscriminator(vocab=vocab, arch=SMALL_D_ARCH, normalization='batch',
                                activation='leakyrelu-0.2', padding='valid',
                                object_size=16)
  sd_img, sd_obj = clone_sd(d_img), clone_sd(d_obj)
  fake = torch.randn(4, 3, 32, 32)
  s_img_real, s_img_fake = d_img(imgs), d_img(fake)
  s_obj, ac = d_obj(imgs, objs, boxes, obj_to_img)
  save('disc.pt', dict(arch=SMALL_D_ARCH, crop=16, vocab=vocab, sd_img=sd_img,
                       sd_obj=sd_obj, batch=batch, fake=fake,
                       img_scores_real=s_img_real.detach(),
                       img_scores_fake=s_img_fake.detach(),
                       obj_scores=s_obj.detach(), ac_loss=ac.detach(),
                       g_loss=gan_g_loss(s_img_fake).detach(),
                       d_loss=gan_d_loss(s_img_real, s_img_fake).detach()))

  # ---- 6. two full training iterations, scripts/train.py:508-592 verbatim flow
  import importlib.util
  spec = importlib.util.spec_from_file_location('ref_train',
                                                os.path.join(REF, 'scripts', 'train.py'))
  ref_train = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(ref_train)
  args = ref_train.parser.parse_args([])
  from sg2im.utils import LossManager
  model.load_state_dict(sd0)
  model.train()
  d_img.load_state_dict(sd_img)
  d_obj.load_state_dict(sd_obj)
  opt = torch.optim.Adam(model.parameters(), lr=args.learning_rate)
  opt_o = torch.optim.Adam(d_obj.parameters(), lr=args.learning_rate)
  opt_i = torch.optim.Adam(d_img.parameters(), lr=args.learning_rate)
  history = []
  for it in range(2):
    torch.manual_seed(500 + it)
    imgs_pred, boxes_pred, masks_pred, pscores = model(objs, triples, obj_to_img,
                                                       boxes_gt=boxes, masks_gt=None)
    total_loss, losses = ref_train.calculate_model_losses(
        args, False, model, imgs, imgs_pred, boxes, boxes_pred, None, masks_pred,
        triples[:, 1], pscores)
    scores_fake, ac_loss = d_obj(imgs_pred, objs, boxes, obj_to_img)
    total_loss = ref_train.add_loss(total_loss, ac_loss, losses, 'ac_loss',
                                    args.ac_loss_weight)
    w = args.discriminator_loss_weight * args.d_obj_weight
    total_loss = ref_train.add_loss(total_loss, gan_g_loss(scores_fake), losses,
                                    'g_gan_obj_loss', w)
    scores_fake = d_img(imgs_pred)
    w = args.discriminator_loss_weight * args.d_img_weight
    total_loss = ref_train.add_loss(total_loss, gan_g_loss(scores_fake), losses,
                                    'g_gan_img_loss', w)
    losses['total_loss'] = total_loss.item()
    opt.zero_grad()
    total_loss.backward()
    opt.step()
    imgs_fake = imgs_pred.detach()
    dl = LossManager()
    sf, acf = d_obj(imgs_fake, objs, boxes, obj_to_img)
    sr, acr = d_obj(imgs, objs, boxes, obj_to_img)
    dl.add_loss(gan_d_loss(sr, sf), 'd_obj_gan_loss')
    dl.add_loss(acr, 'd_ac_loss_real')
    dl.add_loss(acf, 'd_ac_loss_fake')
    opt_o.zero_grad()
    dl.total_loss.backward()
    opt_o.step()
    losses.update(dl.all_losses)
    dl = LossManager()
    sf = d_img(imgs_fake)                      # fake first, then real (train.py:584-585)
    sr = d_img(imgs)
    dl.add_loss(gan_d_loss(sr, sf), 'd_img_gan_loss')
    opt_i.zero_grad()
    dl.total_loss.backward()
    opt_i.step()
    losses.update(dl.all_losses)
    history.append(losses)
  save('train_step.pt', dict(
      kwargs=dict(SMALL_G, image_size=(32, 32)), arch=SMALL_D_ARCH, crop=16,
      vocab=vocab, batch=batch, sd_g=sd0, sd_img=sd_img, sd_obj=sd_obj,
      noise_seeds=[500, 501], losses=history,
      sd_g_after=clone_sd(model), sd_img_after=clone_sd(d_img),
      sd_obj_after=clone_sd(d_obj)))


if __name__ == '__main__':
  main()
