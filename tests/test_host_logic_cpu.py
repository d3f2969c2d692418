"""Host-logic tests on CPU: the nn.Module mirror and the training iteration are
replayed against the golden fixtures (outputs of the unmodified reference) with
the CUDA op boundary swapped for its mathematical definition
(tests/cpu_shim.py).  This pins the WIRING — slices, orders, buffers, which
BatchNorm sees which tensor, loss assembly, optimiser plumbing — without a GPU;
the kernels themselves are pinned by the `-m gpu` tests."""
import contextlib
import copy
import io

import torch

from conftest import load_golden, rel_err
from cpu_shim import cpu_ops

TOL = 1e-5


def _quiet():
  return contextlib.redirect_stdout(io.StringIO())


def _noise(seed, n, nd, hw):
  torch.manual_seed(seed)
  return torch.randn(n, nd, hw[0], hw[1])


def _generator(g):
  from sg2im_b200.model import Sg2ImModel
  with _quiet():
    m = Sg2ImModel(vocab=g['vocab'], **g['kwargs'])
  m.load_state_dict(g['sd'] if 'sd' in g else g['sd_g'])
  return m


def test_generator_wiring_vg_coco_eval():
  g = load_golden('generator.pt')
  imgs, objs, boxes, triples, o2i, _ = g['batch']
  kw = g['kwargs']
  noise = _noise(g['noise_seed'], imgs.size(0), kw['layout_noise_dim'], kw['image_size'])
  with cpu_ops():
    m = _generator(g)
    m.train()
    out = m(objs, triples, o2i, boxes_gt=boxes, noise=noise)
    for a, b, name in zip(out, g['out_vg'], ('img', 'boxes', 'masks', 'rel')):
      assert rel_err(a, b) < TOL, name
    sd = m.state_dict()
    for k, v in g['running_after_vg'].items():
      assert rel_err(sd[k], v) < TOL, k
    assert int(sd['mask_net.1.num_batches_tracked']) == 1
    m = _generator(g)
    m.train()
    out = m(objs, triples, o2i, boxes_gt=boxes, masks_gt=g['gt_masks'], noise=noise,
            num_imgs=imgs.size(0))
    for a, b in zip(out, g['out_coco']):
      assert rel_err(a, b) < TOL
    m = _generator(g)
    m.eval()
    with torch.no_grad():
      out = m(objs, triples, o2i, noise=noise)
    for a, b in zip(out, g['out_eval']):
      assert rel_err(a, b) < TOL


def test_forward_json_wiring_config1():
  g = load_golden('sheep.pt')
  kw = g['kwargs']
  with cpu_ops():
    m = _generator(g)
    m.eval()
    torch.manual_seed(g['noise_seed'])                  # forward_json draws the noise itself
    with torch.no_grad():
      out = m.forward_json(copy.deepcopy(g['scene_graphs']))
  for a, b in zip(out, g['out']):
    assert rel_err(a, b) < TOL
  assert out[0].shape == (7, 3, 64, 64)


def _discriminators(g):
  from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator
  with _quiet():
    d_img = PatchDiscriminator(arch=g['arch'], normalization='batch',
                               activation='leakyrelu-0.2', padding='valid')
    d_obj = AcCropDiscriminator(vocab=g['vocab'], arch=g['arch'], normalization='batch',
                                activation='leakyrelu-0.2', padding='valid',
                                object_size=g['crop'])
  d_img.load_state_dict(g['sd_img'])
  d_obj.load_state_dict(g['sd_obj'])
  return d_obj, d_img


def test_discriminator_wiring():
  from sg2im_b200.losses import gan_g_loss, gan_d_loss
  g = load_golden('disc.pt')
  imgs, objs, boxes, triples, o2i, _ = g['batch']
  with cpu_ops():
    d_obj, d_img = _discriminators(g)
    s_real = d_img(imgs)
    s_fake = d_img(g['fake'])
    assert rel_err(s_real, g['img_scores_real']) < TOL
    assert rel_err(s_fake, g['img_scores_fake']) < TOL
    assert rel_err(gan_g_loss(s_fake), g['g_loss']) < TOL
    assert rel_err(gan_d_loss(s_real, s_fake), g['d_loss']) < TOL
    s_obj, ac = d_obj(imgs, objs, boxes, o2i)
    assert rel_err(s_obj, g['obj_scores']) < TOL
    assert rel_err(ac, g['ac_loss']) < TOL
    # the never-applied classifier stays in the state_dict (discriminators.py:40-45)
    assert 'classifier.weight' in d_img.state_dict()


def test_training_iteration_wiring():
  """scripts/train.py:508-592 flow of TrainStep (flat gradient buckets, frozen
  discriminators during the generator step, loss weights, three Adam steps) vs
  two iterations of the unmodified reference."""
  from sg2im_b200.train_step import TrainStep
  g = load_golden('train_step.pt')
  kw = g['kwargs']
  with cpu_ops():
    m = _generator(g)
    d_obj, d_img = _discriminators(g)
    step = TrainStep(m, d_obj, d_img)
    N = g['batch'][0].size(0)
    for it, seed in enumerate(g['noise_seeds']):
      noise = _noise(seed, N, kw['layout_noise_dim'], kw['image_size'])
      losses, imgs_fake = step.step(g['batch'], noise=noise)
      for k, v in g['losses'][it].items():
        assert abs(losses[k] - v) <= 1e-5 * max(1.0, abs(v)), (it, k, losses[k], v)
      assert imgs_fake.shape == (N, 3) + tuple(kw['image_size']) and not imgs_fake.requires_grad
    # every parameter moved like the reference's (bias-before-BN entries are Adam noise there
    # and exactly still here: allow 2 * lr)
    for net, after in ((m, g['sd_g_after']), (d_obj, g['sd_obj_after']), (d_img, g['sd_img_after'])):
      sd = net.state_dict()
      for k, v in after.items():
        if v.dtype.is_floating_point:
          assert (sd[k] - v).abs().max() < 2.5e-4, k
        else:
          assert torch.equal(sd[k], v), k
    # gradients live in one flat bucket per network
    for name, bucket in step.buckets.items():
      for p in bucket.params:
        assert p.grad is not None and p.grad.data_ptr() >= bucket.flat.data_ptr()
    # discriminators are trainable again after the step
    assert all(p.requires_grad for p in d_obj.parameters())


def test_eval_bn_folding_wiring():
  """Inference path: eval-mode BatchNorm folded into the preceding convolution
  (crn.FOLD_EVAL_BN) reproduces the reference's eval outputs; train mode is
  untouched by the switch."""
  from sg2im_b200 import crn
  g = load_golden('generator.pt')
  imgs, objs, boxes, triples, o2i, _ = g['batch']
  kw = g['kwargs']
  noise = _noise(g['noise_seed'], imgs.size(0), kw['layout_noise_dim'], kw['image_size'])
  s = load_golden('sheep.pt')
  crn.FOLD_EVAL_BN = True
  try:
    with cpu_ops():
      m = _generator(g)
      # make the running statistics non-trivial so the fold is actually exercised
      for name, buf in m.named_buffers():
        if name.endswith('running_mean'):
          buf.copy_(torch.linspace(-0.3, 0.4, buf.numel()))
        elif name.endswith('running_var'):
          buf.copy_(torch.linspace(0.5, 1.7, buf.numel()))
      m.eval()
      with torch.no_grad():
        folded = m(objs, triples, o2i, noise=noise)
        crn.FOLD_EVAL_BN = False
        plain = m(objs, triples, o2i, noise=noise)
        crn.FOLD_EVAL_BN = True
      for a, b in zip(folded, plain):
        assert rel_err(a, b) < TOL
      assert not torch.equal(folded[0], plain[0])           # a different op order did run
      # train mode ignores the switch (batch statistics cannot be folded)
      m.train()
      out = m(objs, triples, o2i, boxes_gt=boxes, noise=noise)
      m2 = _generator(g)
      for name, buf in m2.named_buffers():
        if name.endswith('running_mean'):
          buf.copy_(torch.linspace(-0.3, 0.4, buf.numel()))
        elif name.endswith('running_var'):
          buf.copy_(torch.linspace(0.5, 1.7, buf.numel()))
      m2.train()
      crn.FOLD_EVAL_BN = False
      ref = m2(objs, triples, o2i, boxes_gt=boxes, noise=noise)
      assert torch.equal(out[0], ref[0])
      # config 1 (figure_6_sheep.json through forward_json) with folding on
      crn.FOLD_EVAL_BN = True
      ms = _generator(s)
      ms.eval()
      torch.manual_seed(s['noise_seed'])
      with torch.no_grad():
        out = ms.forward_json(copy.deepcopy(s['scene_graphs']))
      for a, b in zip(out, s['out']):
        assert rel_err(a, b) < TOL
  finally:
    crn.FOLD_EVAL_BN = False


def test_flat_adam_matches_torch_adam():
  """FlatAdam (one flat bucket, sg2im_adam_flat arithmetic) tracks
  torch.optim.Adam over several steps, keeps state_dict keys/shapes, pads
  odd-sized parameters to 16 bytes, and skips on found_inf."""
  from sg2im_b200.train_step import FlatGrads, FlatAdam
  torch.manual_seed(3)
  def net():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 1))
  a, b = net(), net()
  keys = list(a.state_dict().keys())
  opt_ref = torch.optim.Adam(a.parameters(), lr=1e-2)
  with cpu_ops():
    bucket = FlatGrads(b.parameters(), align=4)
    opt = FlatAdam(bucket, lr=1e-2)
    assert all(o % 4 == 0 for o in bucket.offsets)
    assert bucket.flat.numel() == sum(-(-p.numel() // 4) * 4 for p in b.parameters())
    assert list(b.state_dict().keys()) == keys
    for it in range(6):
      x = torch.randn(11, 7)
      opt_ref.zero_grad()
      a(x).pow(2).mean().backward()
      opt_ref.step()
      bucket.zero()
      b(x).pow(2).mean().backward()
      opt.step()
      for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-7), it
    assert float(opt.step_count) == 6
    # parameters are views of the flat bucket; the padding never moves
    for p, o in zip(bucket.params, bucket.offsets):
      assert p.data_ptr() == opt.flat_params.data_ptr() + 4 * o
    used = torch.zeros_like(opt.flat_params, dtype=torch.bool)
    for p, o in zip(bucket.params, bucket.offsets):
      used[o:o + p.numel()] = True
    assert bool((opt.flat_params[~used] == 0).all())
    # found_inf: nothing moves, the step count stays
    before = opt.flat_params.clone()
    opt.found_inf = torch.ones(())
    opt.step()
    assert torch.equal(opt.flat_params, before) and float(opt.step_count) == 6
    # state round trip still works on the re-pointed parameters
    b.load_state_dict(a.state_dict())
    for pa, pb in zip(a.parameters(), b.parameters()):
      assert torch.equal(pa, pb)
    assert bucket.params[0].data_ptr() == opt.flat_params.data_ptr()


def test_flat_adam_checkpoint_in_torch_adam_format():
  """FlatAdam.state_dict() is torch.optim.Adam's (per-parameter step / exp_avg / exp_avg_sq, one
  param group): a run checkpointed with either optimiser resumes with the other, like the
  reference's `optimizer.state_dict()` round trip (scripts/train.py:633-641, :454-463); and a
  parameter re-allocated behind the optimiser's back is detected instead of silently detached."""
  import pytest
  from sg2im_b200.train_step import FlatGrads, FlatAdam
  def net():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
  def batches(n, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(11, 7, generator=g) for _ in range(n)]
  with cpu_ops():
    # 3 steps with torch Adam -> checkpoint -> 3 more with FlatAdam  ==  6 steps of torch Adam
    ref = net(); opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-2)
    for x in batches(6, 1):
      opt_ref.zero_grad(); ref(x).pow(2).mean().backward(); opt_ref.step()
    a = net(); opt_a = torch.optim.Adam(a.parameters(), lr=1e-2)
    xs = batches(6, 1)
    for x in xs[:3]:
      opt_a.zero_grad(); a(x).pow(2).mean().backward(); opt_a.step()
    b = net(); b.load_state_dict(a.state_dict())
    bucket = FlatGrads(b.parameters(), align=4)
    opt_b = FlatAdam(bucket, lr=1e-2)
    opt_b.load_state_dict(opt_a.state_dict())
    for x in xs[3:]:
      bucket.zero(); b(x).pow(2).mean().backward(); opt_b.step()
    for pr, pb in zip(ref.parameters(), b.parameters()):
      assert torch.allclose(pr, pb, rtol=1e-5, atol=1e-7)
    # ... and back: FlatAdam's checkpoint loads into torch.optim.Adam
    sd = opt_b.state_dict()
    assert set(sd) == {'state', 'param_groups'} and len(sd['state']) == 4
    c = net(); c.load_state_dict(b.state_dict())
    opt_c = torch.optim.Adam(c.parameters(), lr=1e-2)
    opt_c.load_state_dict(sd)
    x = batches(1, 9)[0]
    opt_c.zero_grad(); c(x).pow(2).mean().backward(); opt_c.step()
    bucket.zero(); b(x).pow(2).mean().backward(); opt_b.step()
    for pc, pb in zip(c.parameters(), b.parameters()):
      assert torch.allclose(pc, pb, rtol=1e-5, atol=1e-7)
    # a parameter moved out of the flat bucket: loud failure
    next(b.parameters()).data = next(b.parameters()).data.clone()
    with pytest.raises(RuntimeError, match='re-allocated'):
      opt_b.step()


def test_training_iteration_with_flat_adam():
  """TrainStep(fused_adam='flat') reproduces the reference's two iterations like
  the torch.optim.Adam configuration does."""
  from sg2im_b200.train_step import TrainStep, FlatAdam
  g = load_golden('train_step.pt')
  kw = g['kwargs']
  with cpu_ops():
    m = _generator(g)
    d_obj, d_img = _discriminators(g)
    step = TrainStep(m, d_obj, d_img, fused_adam='flat')
    assert all(isinstance(o, FlatAdam) for o in step.opts.values())
    N = g['batch'][0].size(0)
    for it, seed in enumerate(g['noise_seeds']):
      noise = _noise(seed, N, kw['layout_noise_dim'], kw['image_size'])
      losses, _ = step.step(g['batch'], noise=noise)
      for k, v in g['losses'][it].items():
        assert abs(losses[k] - v) <= 1e-5 * max(1.0, abs(v)), (it, k, losses[k], v)
    for net, after in ((m, g['sd_g_after']), (d_obj, g['sd_obj_after']), (d_img, g['sd_img_after'])):
      sd = net.state_dict()
      assert list(sd.keys()) == list(after.keys())
      for k, v in after.items():
        if v.dtype.is_floating_point:
          assert (sd[k] - v).abs().max() < 2.5e-4, k
