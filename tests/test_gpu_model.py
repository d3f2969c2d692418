"""GPU parity tests, whole modules: generator, discriminators and the full
training iteration against the golden fixtures (reference outputs) and the
CPU oracle.  Tolerance 1e-3 relative (north_star), measured values are ~1e-5
on the exact-fp32 path."""
import contextlib
import io

import pytest
import torch

from conftest import rel_err, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-3


def dev():
  return torch.device('cuda:0')


def _noise(seed, n, nd, hw):
  torch.manual_seed(seed)
  return torch.randn(n, nd, hw[0], hw[1])


def _quiet():
  return contextlib.redirect_stdout(io.StringIO())


def _build_generator(g):
  from sg2im_b200.model import Sg2ImModel
  with _quiet():
    m = Sg2ImModel(vocab=g['vocab'], **g['kwargs'])
  m.load_state_dict(g['sd'])           # strict: keys/shapes equal the reference's
  return m.to(dev())


def test_generator_forward_vg_coco_eval():
  g = load_golden('generator.pt')
  imgs, objs, boxes, triples, o2i, _ = [t.to(dev()) for t in g['batch']]
  kw = g['kwargs']
  noise = _noise(g['noise_seed'], imgs.size(0), kw['layout_noise_dim'], kw['image_size']).to(dev())
  m = _build_generator(g)
  m.train()
  out = m(objs, triples, o2i, boxes_gt=boxes, noise=noise)
  for a, b, name in zip(out, g['out_vg'], ('img', 'boxes', 'masks', 'rel')):
    assert rel_err(a, b) < TOL, name
  sd = m.state_dict()
  for k, v in g['running_after_vg'].items():
    assert rel_err(sd[k], v) < TOL, k
  m = _build_generator(g)
  m.train()
  out = m(objs, triples, o2i, boxes_gt=boxes, masks_gt=g['gt_masks'].to(dev()), noise=noise,
          num_imgs=imgs.size(0))
  for a, b in zip(out, g['out_coco']):
    assert rel_err(a, b) < TOL
  m = _build_generator(g)
  m.eval()
  with torch.no_grad():
    out = m(objs, triples, o2i, noise=noise)
  for a, b in zip(out, g['out_eval']):
    assert rel_err(a, b) < TOL


def test_config1_sheep_forward_json():
  import copy
  g = load_golden('sheep.pt')
  m = _build_generator(g)
  m.eval()
  objs, triples, o2i = m.encode_scene_graphs(copy.deepcopy(g['scene_graphs']))
  for a, b in zip((objs, triples, o2i), g['encoded']):
    assert torch.equal(a.cpu(), b)
  kw = g['kwargs']
  noise = _noise(g['noise_seed'], 7, kw['layout_noise_dim'], kw['image_size']).to(dev())
  with torch.no_grad():
    out = m(objs, triples, o2i, noise=noise)
  for a, b in zip(out, g['out']):
    assert rel_err(a, b) < TOL


def test_discriminators_forward():
  from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator
  from sg2im_b200.losses import gan_g_loss, gan_d_loss
  g = load_golden('disc.pt')
  imgs, objs, boxes, triples, o2i, _ = [t.to(dev()) for t in g['batch']]
  with _quiet():
    d_img = PatchDiscriminator(arch=g['arch'], normalization='batch',
                               activation='leakyrelu-0.2', padding='valid')
    d_obj = AcCropDiscriminator(vocab=g['vocab'], arch=g['arch'], normalization='batch',
                                activation='leakyrelu-0.2', padding='valid',
                                object_size=g['crop'])
  d_img.load_state_dict(g['sd_img'])
  d_obj.load_state_dict(g['sd_obj'])
  d_img, d_obj = d_img.to(dev()), d_obj.to(dev())
  s_real = d_img(imgs)
  s_fake = d_img(g['fake'].to(dev()))
  assert rel_err(s_real, g['img_scores_real']) < TOL
  assert rel_err(s_fake, g['img_scores_fake']) < TOL
  assert rel_err(gan_g_loss(s_fake), g['g_loss']) < TOL
  assert rel_err(gan_d_loss(s_real, s_fake), g['d_loss']) < TOL
  s_obj, ac = d_obj(imgs, objs, boxes, o2i)
  assert rel_err(s_obj, g['obj_scores']) < TOL
  assert rel_err(ac, g['ac_loss']) < TOL


def _build_all(g):
  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator
  with _quiet():
    m = Sg2ImModel(vocab=g['vocab'], **g['kwargs'])
    d_img = PatchDiscriminator(arch=g['arch'], normalization='batch',
                               activation='leakyrelu-0.2', padding='valid')
    d_obj = AcCropDiscriminator(vocab=g['vocab'], arch=g['arch'], normalization='batch',
                                activation='leakyrelu-0.2', padding='valid',
                                object_size=g['crop'])
  m.load_state_dict(g['sd_g'])
  d_img.load_state_dict(g['sd_img'])
  d_obj.load_state_dict(g['sd_obj'])
  return m.to(dev()), d_obj.to(dev()), d_img.to(dev())


def test_two_training_iterations_match_reference():
  """scripts/train.py:508-592 reproduced: per-iteration losses of the
  unmodified reference (golden) vs the CUDA path, 2 iterations incl. Adam."""
  from sg2im_b200.train_step import TrainStep
  g = load_golden('train_step.pt')
  m, d_obj, d_img = _build_all(g)
  step = TrainStep(m, d_obj, d_img)
  batch = [t.to(dev()) for t in g['batch']]
  kw = g['kwargs']
  for it, seed in enumerate(g['noise_seeds']):
    noise = _noise(seed, batch[0].size(0), kw['layout_noise_dim'], kw['image_size']).to(dev())
    losses, _ = step.step(batch, noise=noise)
    for k, v in g['losses'][it].items():
      assert abs(losses[k] - v) <= 1e-3 * max(1.0, abs(v)), (it, k, losses[k], v)
  # parameters after 2 Adam steps: moved by at most 2*lr; must track the reference
  for net, after in ((m, g['sd_g_after']), (d_obj, g['sd_obj_after']), (d_img, g['sd_img_after'])):
    sd = net.state_dict()
    for k, v in after.items():
      if v.dtype.is_floating_point:
        assert (sd[k].cpu() - v).abs().max() < 2.5e-4, k


def test_generator_gradients_vs_oracle():
  """Full backward through CRN, layout, mask head and graph convolution:
  parameter gradients of a scalar loss vs the CPU oracle's autograd."""
  from oracle import sg2im_oracle as orc
  g = load_golden('generator.pt')
  batch = g['batch']
  imgs, objs, boxes, triples, o2i, _ = batch
  kw = g['kwargs']
  N = imgs.size(0)
  noise = _noise(5, N, kw['layout_noise_dim'], kw['image_size'])
  sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and 'running' not in k
            else v.clone()) for k, v in g['sd'].items()}
  ref = orc.generator_forward(sd, kw['image_size'], objs, triples, o2i, boxes_gt=boxes,
                              noise=noise, training=True, num_imgs=N)
  wimg = torch.randn(ref[0].shape, generator=torch.Generator().manual_seed(9))
  (ref[0] * wimg).sum().add(ref[1].pow(2).sum()).backward()
  m = _build_generator(g)
  m.train()
  d = dev()
  out = m(objs.to(d), triples.to(d), o2i.to(d), boxes_gt=boxes.to(d), noise=noise.to(d), num_imgs=N)
  ((out[0] * wimg.to(d)).sum() + out[1].pow(2).sum()).backward()
  worst = 0.0
  for k, p in m.named_parameters():
    rg = sd[k].grad
    if rg is None:
      assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
      continue
    if '.net.0.bias' in k or '.net.3.bias' in k:
      continue      # conv bias feeding a train-mode BN: the true gradient is 0, both sides hold rounding noise
    e = rel_err(p.grad, rg)
    worst = max(worst, e)
    assert e < TOL, (k, e)
  print('worst param-grad rel err', worst)


def test_training_iteration_tf32_tensor_core_path():
  """Same two reference iterations with the convolutions on the tcgen05 TF32
  kernels (forward, dgrad, wgrad).  TF32 operand truncation (2^-10 relative,
  what cuDNN's default allow_tf32 path also does for the reference on GPU)
  moves the losses by ~1e-3 relative; tolerance stated: 1e-2."""
  from sg2im_b200 import ops
  from sg2im_b200.train_step import TrainStep
  g = load_golden('train_step.pt')
  m, d_obj, d_img = _build_all(g)
  ops.set_conv_math('tf32')
  try:
    step = TrainStep(m, d_obj, d_img)
    batch = [t.to(dev()) for t in g['batch']]
    kw = g['kwargs']
    worst = 0.0
    for it, seed in enumerate(g['noise_seeds']):
      noise = _noise(seed, batch[0].size(0), kw['layout_noise_dim'], kw['image_size']).to(dev())
      losses, _ = step.step(batch, noise=noise)
      for k, v in g['losses'][it].items():
        e = abs(losses[k] - v) / max(1.0, abs(v))
        worst = max(worst, e)
        assert e <= 1e-2, (it, k, losses[k], v)
    print('worst loss deviation under tf32', worst)
  finally:
    ops.set_conv_math('fp32')


def test_cuda_graph_step_matches_eager_step():
  """TrainStep(cuda_graph=True): replayed iterations give the same losses and
  parameters as eager iterations (same batches, same injected noise)."""
  from sg2im_b200.train_step import TrainStep
  g = load_golden('train_step.pt')
  kw = g['kwargs']
  batch = [t.to(dev()) for t in g['batch']]
  N = batch[0].size(0)
  runs = {}
  for mode in (False, True):
    m, d_obj, d_img = _build_all(g)
    step = TrainStep(m, d_obj, d_img, cuda_graph=mode, graph_warmup=2)
    hist = []
    for it in range(6):
      noise = _noise(900 + it, N, kw['layout_noise_dim'], kw['image_size']).to(dev())
      losses, imgs = step.step(batch, noise=noise)
      hist.append(losses)
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    # the discriminators too (parameters and BatchNorm buffers): the graph-mode step runs their
    # iteration on a second stream beside the generator backward
    state.update({'d_obj.' + k: v.detach().clone() for k, v in d_obj.state_dict().items()})
    state.update({'d_img.' + k: v.detach().clone() for k, v in d_img.state_dict().items()})
    runs[mode] = (hist, state)
    if mode:
      assert step.replays == 4 and step.launches_per_replay > 100     # calls 2..5 replay
  for it in range(6):
    for k, v in runs[False][0][it].items():
      assert abs(runs[True][0][it][k] - v) <= 2e-4 * max(1.0, abs(v)), (it, k)
  assert any('running_mean' in k and k.startswith('d_') for k in runs[False][1])
  for k, v in runs[False][1].items():
    if v.dtype.is_floating_point:
      assert (runs[True][1][k] - v).abs().max() < 1e-3, k
    else:
      assert torch.equal(runs[True][1][k], v), k                       # num_batches_tracked


def test_generator_forward_tf32_error_vs_fp32_reference():
  """End-to-end image error of the tensor-core (TF32) path against the fp32
  reference output (golden), same weights / inputs / noise.  TF32 keeps 10
  mantissa bits per operand (the hardware drops the low 13 bits of the fp32
  words it is fed), so per-conv error is ~1e-3 and BatchNorm re-normalises
  between layers; measured end to end ~2e-3, asserted < 1e-2.  The exact-fp32
  path (set_conv_math('fp32')) meets the 1e-3 bar with margin (see
  test_generator_forward_vg_coco_eval)."""
  from sg2im_b200 import ops
  g = load_golden('generator.pt')
  imgs, objs, boxes, triples, o2i, _ = [t.to(dev()) for t in g['batch']]
  kw = g['kwargs']
  noise = _noise(g['noise_seed'], imgs.size(0), kw['layout_noise_dim'], kw['image_size']).to(dev())
  ops.set_conv_math('tf32')
  try:
    m = _build_generator(g)
    m.train()
    out = m(objs, triples, o2i, boxes_gt=boxes, noise=noise)
    errs = [rel_err(a, b) for a, b in zip(out, g['out_vg'])]
    print('tf32 end-to-end rel err (img, boxes, masks, rel):', errs)
    assert max(errs) < 1e-2
  finally:
    ops.set_conv_math('fp32')
