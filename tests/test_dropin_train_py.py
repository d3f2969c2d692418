"""Drop-in check (build container only): the reference's own scripts/train.py,
UNMODIFIED, imported with sg2im.{model,discriminators,losses,...} aliased to
sg2im_b200 (the recipe of INTEGRATION.md §1).  Its builders
(build_model / build_obj_discriminator / build_img_discriminator) construct our
modules from its argparse defaults, and its loss assembly + one optimiser step
run on them (CUDA op boundary swapped for the CPU shim)."""
import contextlib
import importlib.util
import io
import os
import sys

import pytest
import torch

from refimport import have_reference, import_reference, REF

pytestmark = pytest.mark.skipif(not have_reference(), reason='reference tree not mounted')

ALIASED = ('model', 'discriminators', 'losses', 'layers', 'graph', 'crn', 'layout', 'bilinear')


@contextlib.contextmanager
def aliased_reference():
  import_reference()
  import sg2im
  import sg2im_b200.model, sg2im_b200.discriminators, sg2im_b200.losses, sg2im_b200.layers  # noqa
  import sg2im_b200.graph, sg2im_b200.crn, sg2im_b200.layout, sg2im_b200.bilinear  # noqa
  import sg2im_b200
  saved = {}
  for name in ALIASED:
    key = 'sg2im.' + name
    saved[key] = sys.modules.get(key)
    sys.modules[key] = getattr(sg2im_b200, name)
    saved['attr.' + name] = getattr(sg2im, name, None)
    setattr(sg2im, name, getattr(sg2im_b200, name))
  try:
    yield
  finally:
    for name in ALIASED:
      key = 'sg2im.' + name
      if saved[key] is None:
        sys.modules.pop(key, None)
      else:
        sys.modules[key] = saved[key]
      if saved['attr.' + name] is not None:
        setattr(sg2im, name, saved['attr.' + name])


def test_reference_train_py_builds_and_steps_our_modules():
  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator
  from sg2im_b200.synth import make_vocab, synth_batch
  from cpu_shim import cpu_ops
  with aliased_reference():
    spec = importlib.util.spec_from_file_location('ref_train_dropin',
                                                  os.path.join(REF, 'scripts', 'train.py'))
    train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(train)
    args = train.parser.parse_args([
        '--image_size', '32,32', '--refinement_network_dims', '16,8', '--embedding_dim', '8',
        '--gconv_dim', '8', '--gconv_hidden_dim', '16', '--gconv_num_layers', '2',
        '--mask_size', '8', '--layout_noise_dim', '4', '--crop_size', '16',
        '--d_obj_arch', 'C4-8-2,C4-16-2', '--d_img_arch', 'C4-8-2,C4-16-2'])
    train.check_args(args)
    vocab = make_vocab(7, 4)
    with contextlib.redirect_stdout(io.StringIO()):
      model, model_kwargs = train.build_model(args, vocab)
      d_obj, _ = train.build_obj_discriminator(args, vocab)
      d_img, _ = train.build_img_discriminator(args, vocab)
    assert type(model) is Sg2ImModel
    assert type(d_obj) is AcCropDiscriminator and type(d_img) is PatchDiscriminator
    assert model_kwargs['image_size'] == (32, 32)
    gan_g_loss, gan_d_loss = train.get_gan_losses(args.gan_loss_type)

    imgs, objs, boxes, triples, o2i, _ = synth_batch(N=3, objs_per_img=3, rels_per_img=2,
                                                     image_size=(32, 32), num_objs=7, num_preds=4,
                                                     seed=1)
    opt = torch.optim.Adam(model.parameters(), lr=args.learning_rate)
    with cpu_ops():
      # the forward call and loss assembly of train.py:524-550, its own functions
      imgs_pred, boxes_pred, masks_pred, pscores = model(objs, triples, o2i, boxes_gt=boxes,
                                                         masks_gt=None)
      total, losses = train.calculate_model_losses(args, False, model, imgs, imgs_pred, boxes,
                                                   boxes_pred, None, masks_pred, triples[:, 1],
                                                   pscores)
      scores_fake, ac_loss = d_obj(imgs_pred, objs, boxes, o2i)
      total = train.add_loss(total, ac_loss, losses, 'ac_loss', args.ac_loss_weight)
      total = train.add_loss(total, gan_g_loss(scores_fake), losses, 'g_gan_obj_loss',
                             args.discriminator_loss_weight * args.d_obj_weight)
      total = train.add_loss(total, gan_g_loss(d_img(imgs_pred)), losses, 'g_gan_img_loss',
                             args.discriminator_loss_weight * args.d_img_weight)
      before = model.refinement_net.output_conv[2].weight.detach().clone()
      opt.zero_grad()
      total.backward()
      opt.step()
    assert imgs_pred.shape == (3, 3, 32, 32) and masks_pred.shape == (objs.numel(), 8, 8)
    assert set(losses) >= {'L1_pixel_loss', 'bbox_pred', 'ac_loss', 'g_gan_obj_loss', 'g_gan_img_loss'}
    assert all(v == v for v in losses.values())
    assert not torch.equal(before, model.refinement_net.output_conv[2].weight)
    # checkpoint hand-off in the reference's own format (train.py:473-500, :161-172)
    ckpt = {'model_kwargs': model_kwargs, 'model_state': model.state_dict()}
    with contextlib.redirect_stdout(io.StringIO()):
      clone = Sg2ImModel(**ckpt['model_kwargs'])
    clone.load_state_dict(ckpt['model_state'])
