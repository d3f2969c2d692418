"""GPU parity at the object / triple counts of the other BASELINE.json configurations, and the
reference's own loop body driven with the plain modules (no TrainStep).

* C2 (COCO-64: 6 + 1 objects and 12 triples per image, ground-truth masks, 64 x 64) and C5 (dense
  graphs: 32 + 1 objects and 64 triples per image) with the DEFAULT architecture of scripts/train.py
  (embedding 128, gconv 128 / 512 / 5 layers, CRN 1024-512-256-128-64, mask 16, noise 32): generator
  forward and full backward against the CPU oracle.  The batch is 8 images and the dense case runs at
  64 x 64 so that the reference algorithm's (O, D, H, W) layout temporary (sg2im/layout.py:86-90)
  stays under ~1 GB on the host; every kernel still sees the configuration's per-image counts.
* scripts/train.py:508-592 with our modules as plain nn.Modules — three torch.optim.Adam
  optimisers, autograd, .backward(), .step(), exactly the reference's statements — against the
  golden losses of the unmodified reference.  (tests/test_dropin_train_py.py drives the modules from
  the reference's own train.py on the CPU shim; /root/reference does not exist on the GPU box.)
"""
import contextlib
import io

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err, load_golden

pytestmark = pytest.mark.gpu


def dev():
  return torch.device('cuda:0')


CASES = {
    # name: (N, objs/img, rels/img, image size, gt masks, num_objs, num_preds)
    'C2-coco64': dict(N=8, objs_per_img=6, rels_per_img=6, image_size=(64, 64), masks=True, num_objs=184,
                      num_preds=7),
    'C5-dense': dict(N=8, objs_per_img=32, rels_per_img=32, image_size=(64, 64), masks=False, num_objs=179,
                     num_preds=46),
}


@pytest.mark.parametrize('math', ['fp32', 'bf16x3'])
@pytest.mark.parametrize('name', list(CASES))
def test_default_architecture_at_config_counts_vs_oracle(name, math):
  from oracle import sg2im_oracle as orc
  from sg2im_b200 import ops
  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.synth import make_vocab, synth_batch
  cfg = CASES[name]
  H, W = cfg['image_size']
  torch.manual_seed(0)
  vocab = make_vocab(cfg['num_objs'], cfg['num_preds'])
  with contextlib.redirect_stdout(io.StringIO()):
    m = Sg2ImModel(vocab, image_size=(H, W), embedding_dim=128, gconv_dim=128, gconv_hidden_dim=512,
                   gconv_num_layers=5, mask_size=16, layout_noise_dim=32,
                   refinement_dims=(1024, 512, 256, 128, 64))
  batch = synth_batch(seed=3, **cfg)
  masks = None
  if len(batch) == 7:
    imgs, objs, boxes, masks, triples, o2i, _ = batch
  else:
    imgs, objs, boxes, triples, o2i, _ = batch
  N = imgs.size(0)
  noise = torch.randn(N, 32, H, W, generator=torch.Generator().manual_seed(4))
  wimg = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(5))
  sd = {k: (v.detach().clone().requires_grad_(True) if v.dtype.is_floating_point and 'running' not in k
            else v.detach().clone()) for k, v in m.state_dict().items()}
  torch.set_num_threads(16)
  ref = orc.generator_forward(sd, (H, W), objs, triples, o2i, boxes_gt=boxes, masks_gt=masks, noise=noise,
                              training=True, num_imgs=N)
  (ref[0] * wimg).sum().add(ref[1].pow(2).sum()).backward()
  ops.set_conv_math(math)
  try:
    d = dev()
    m = m.to(d)
    m.train()
    out = m(objs.to(d), triples.to(d), o2i.to(d), boxes_gt=boxes.to(d),
            masks_gt=None if masks is None else masks.to(d), noise=noise.to(d), num_imgs=N)
    errs = [rel_err(a, b) for a, b in zip(out, ref) if a is not None and b is not None]
    print(name, math, 'forward rel err', errs)
    assert max(errs) < 1e-3
    ((out[0] * wimg.to(d)).sum() + out[1].pow(2).sum()).backward()
    worst, worst_cos = 0.0, 1.0
    for k, p in m.named_parameters():
      rg = sd[k].grad
      if rg is None or p.grad is None or '.net.0.bias' in k or '.net.3.bias' in k:
        continue
      if float(rg.abs().max()) == 0.0:
        continue
      a, b = p.grad.detach().double().cpu().flatten(), rg.double().flatten()
      worst = max(worst, rel_err(p.grad, rg))
      worst_cos = min(worst_cos, float((a @ b) / (a.norm() * b.norm()).clamp(min=1e-300)))
    print(name, math, 'param-grad worst rel err %.2e, worst cosine %.6f' % (worst, worst_cos))
    # whole-network gradients are bounded by activation-kink flips, not by the arithmetic
    # (tests/test_gpu_bf16x3.py; measured here on the exact-fp32 kernels, forward error 7e-6: worst
    # max-norm 7.6e-2 on one small tensor at cosine 0.99997): the DIRECTION of every parameter
    # gradient must agree (cosine); the max-norm of the difference is dominated by single flipped
    # elements on small-gradient tensors (measured up to 0.44 at cosine 0.9997 in bf16x3) and is
    # only bounded loosely
    assert worst_cos > (0.9995 if math == 'fp32' else 0.999) and worst < (0.25 if math == 'fp32' else 1.0)
  finally:
    ops.set_conv_math('fp32')


@pytest.mark.parametrize('math', ['fp32', 'bf16x3'])
def test_reference_loop_body_with_plain_modules(math):
  """scripts/train.py:508-592, statement for statement, on our modules."""
  import test_gpu_model as G
  from sg2im_b200 import ops
  from sg2im_b200.losses import get_gan_losses
  g = load_golden('train_step.pt')
  ops.set_conv_math(math)
  try:
    model, obj_discriminator, img_discriminator = G._build_all(g)
    gan_g_loss, gan_d_loss = get_gan_losses('gan')
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-4)
    optimizer_d_obj = torch.optim.Adam(obj_discriminator.parameters(), lr=1e-4)
    optimizer_d_img = torch.optim.Adam(img_discriminator.parameters(), lr=1e-4)
    imgs, objs, boxes, triples, obj_to_img, triple_to_img = [t.to(dev()) for t in g['batch']]
    kw = g['kwargs']
    model.train(); obj_discriminator.train(); img_discriminator.train()
    for it, seed in enumerate(g['noise_seeds']):
      noise = G._noise(seed, imgs.size(0), kw['layout_noise_dim'], kw['image_size']).to(dev())
      predicates = triples[:, 1]
      imgs_pred, boxes_pred, masks_pred, predicate_scores = model(objs, triples, obj_to_img, boxes_gt=boxes,
                                                                  noise=noise)
      # calculate_model_losses (train.py:387-412) with the default weights
      losses = {}
      total_loss = torch.zeros(1).to(imgs)
      def add_loss(total, cur, name, weight):
        cur = cur * weight
        losses[name] = cur.item()
        return total + cur
      total_loss = add_loss(total_loss, F.l1_loss(imgs_pred, imgs), 'L1_pixel_loss', 1.0)
      total_loss = add_loss(total_loss, F.mse_loss(boxes_pred, boxes), 'bbox_pred', 10.0)
      scores_fake, ac_loss = obj_discriminator(imgs_pred, objs, boxes, obj_to_img)
      total_loss = add_loss(total_loss, ac_loss, 'ac_loss', 0.1)
      total_loss = add_loss(total_loss, gan_g_loss(scores_fake), 'g_gan_obj_loss', 0.01 * 1.0)
      scores_fake = img_discriminator(imgs_pred)
      total_loss = add_loss(total_loss, gan_g_loss(scores_fake), 'g_gan_img_loss', 0.01 * 1.0)
      losses['total_loss'] = total_loss.item()
      optimizer.zero_grad()
      total_loss.backward()
      optimizer.step()

      d_obj_losses = {}
      imgs_fake = imgs_pred.detach()
      scores_fake, ac_loss_fake = obj_discriminator(imgs_fake, objs, boxes, obj_to_img)
      scores_real, ac_loss_real = obj_discriminator(imgs, objs, boxes, obj_to_img)
      d_obj_gan_loss = gan_d_loss(scores_real, scores_fake)
      d_obj_total = d_obj_gan_loss + ac_loss_real + ac_loss_fake
      losses['d_obj_gan_loss'] = d_obj_gan_loss.item()
      losses['d_ac_loss_real'] = ac_loss_real.item()
      losses['d_ac_loss_fake'] = ac_loss_fake.item()
      optimizer_d_obj.zero_grad()
      d_obj_total.backward()
      optimizer_d_obj.step()

      scores_fake = img_discriminator(imgs_fake)
      scores_real = img_discriminator(imgs)
      d_img_gan_loss = gan_d_loss(scores_real, scores_fake)
      losses['d_img_gan_loss'] = d_img_gan_loss.item()
      optimizer_d_img.zero_grad()
      d_img_gan_loss.backward()
      optimizer_d_img.step()

      for k, v in g['losses'][it].items():
        assert abs(losses[k] - v) <= 1e-3 * max(1.0, abs(v)), (it, k, losses[k], v)
  finally:
    ops.set_conv_math('fp32')
