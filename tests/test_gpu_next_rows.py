"""GPU parity tests for the rows NEXT to the training step and for the
second-generation kernels of the step itself: de-normalisation kernel, validation pass, staged
batch upload, COCO relation synthesis, the align_corners=True (torch-0.4 checkpoint) sampling
convention, eval-mode BatchNorm folding; BatchNorm-backward / layout / normalise-activate v2
kernels, flat Adam, the weight-gradient ("kcc") weight layout with direct gradient accumulation,
the fused activation-backward + bias-gradient pass; the rest of the factory surface (general
pooling, instance normalisation, residual blocks) and the layout's box gradient.

All of these ran green on a B200 in round 2 (profiles/r02_first_call.txt) and are part of the
default `-m gpu` suite.
"""
import contextlib
import io
import os

import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


def dev():
  return torch.device('cuda:0')


def test_deprocess_bytes_identical_to_reference():
  from sg2im_b200.images import imagenet_deprocess_batch
  g = load_golden('aux.pt')['deprocess']
  x = g['imgs'].to(dev())
  out = imagenet_deprocess_batch(x)
  assert out.device.type == 'cpu' and out.dtype == torch.uint8
  assert torch.equal(out, g['rescaled'])
  assert torch.equal(imagenet_deprocess_batch(x, rescale=False), g['plain'])
  # the generator hands out an NCHW view of an NHWC buffer: read in place
  view = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
  assert torch.equal(imagenet_deprocess_batch(view), g['rescaled'])
  nhwc = imagenet_deprocess_batch(view, device_out=True, channels_last=True)
  assert nhwc.device == x.device and torch.equal(nhwc.cpu().permute(0, 3, 1, 2), g['rescaled'])


def test_deprocess_full_size_properties():
  from sg2im_b200.images import imagenet_deprocess_batch
  from oracle import validation_oracle as vorc
  x = torch.randn(32, 3, 128, 128, generator=torch.Generator().manual_seed(3))
  out = imagenet_deprocess_batch(x.to(dev()))
  assert torch.equal(out, vorc.imagenet_deprocess_batch(x))
  flat = out.view(32, -1)
  assert bool((flat.min(dim=1).values == 0).all()) and bool((flat.max(dim=1).values == 255).all())
  with pytest.raises(RuntimeError):
    imagenet_deprocess_batch(x)                        # CPU tensor: refused


@pytest.mark.parametrize('name', ['check_vg', 'check_coco'])
def test_check_model_matches_reference(name):
  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.validate import check_model
  from sg2im_b200 import ops
  ops.set_conv_math('fp32')
  g = load_golden('aux.pt')[name]
  with contextlib.redirect_stdout(io.StringIO()):
    model = Sg2ImModel(vocab=g['vocab'], **g['kwargs'])
  model.load_state_dict(g['sd'])
  model.to(dev()).train()
  mean_losses, samples, batch_data, avg_iou = check_model(g['args'], 0, g['loader'], model)
  for k, v in g['mean_losses'].items():
    assert abs(mean_losses[k] - v) <= 1e-4 * max(1.0, abs(v)), k
  assert abs(float(avg_iou) - float(g['avg_iou'])) < 1e-5
  assert torch.equal(samples['gt_img'], g['samples']['gt_img'])
  for k in ('gt_box_gt_mask', 'gt_box_pred_mask', 'pred_box_pred_mask'):
    d = (samples[k].int() - g['samples'][k].int()).abs()
    assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 0.01, k
  assert rel_err(batch_data['boxes_pred'], g['batch_data']['boxes_pred']) < TOL
  sd = model.state_dict()
  for k, v in g['bn_after'].items():
    assert rel_err(sd[k].float(), v.float()) < TOL, k


def test_staged_batch_upload():
  from sg2im_b200 import batching
  from sg2im_b200.synth import synth_batch
  imgs, objs, boxes, triples, o2i, t2i = synth_batch(N=3, objs_per_img=4, rels_per_img=2,
                                                     image_size=(16, 16), num_objs=9, num_preds=5)
  samples = batching.uncollate((imgs, objs, boxes, triples, o2i, t2i))
  staged = batching.collate(samples, pin=True)
  assert staged.floats.is_pinned() and staged.ints.is_pinned()
  on_dev = staged.to(dev())
  torch.cuda.synchronize()
  for a, b in zip(on_dev, (imgs, objs, boxes, triples, o2i, t2i)):
    assert a.is_cuda and torch.equal(a.cpu(), b)


def test_align_corners_true_layout_and_crop():
  """torch-0.4 sampling convention for the published checkpoints."""
  from sg2im_b200 import layout as L
  from sg2im_b200.bilinear import crop_bbox_batch
  from oracle import sg2im_oracle as orc
  g = load_golden('aux.pt')['align_corners']
  lay, crp = load_golden('layout.pt'), load_golden('crop.pt')
  d = dev()
  L.ALIGN_CORNERS = True
  try:
    m = L.masks_to_layout(lay['rvecs'].to(d), lay['rboxes'].to(d), lay['rmasks'].to(d),
                          lay['robj_to_img'].to(d), 24, 40)
    b = L.boxes_to_layout(lay['vecs'].to(d), lay['boxes'].to(d), lay['obj_to_img'].to(d), 24, 20)
    c = crop_bbox_batch(crp['feats'].to(d), crp['boxes'].to(d), crp['bbox_to_feats'].to(d), 6, 7)
    assert rel_err(m, g['masks']) < TOL and rel_err(b, g['boxes']) < TOL
    assert rel_err(c, g['crops']) < TOL
    # gradients against the oracle under the same convention (the backward kernel needs D % 4 == 0)
    gen = torch.Generator().manual_seed(2)
    O_, D_, M_ = 9, 8, 6
    vecs = torch.randn(O_, D_, generator=gen).requires_grad_(True)
    masks = torch.rand(O_, M_, M_, generator=gen).requires_grad_(True)
    xy = torch.rand(O_, 2, generator=gen) * 0.6
    bx = torch.cat([xy, xy + torch.rand(O_, 2, generator=gen) * 0.35 + 0.1], 1)
    o2i = torch.sort(torch.randint(0, 3, (O_,), generator=gen)).values
    ref = orc.masks_to_layout(vecs, bx, masks, o2i, 24, 40, 3, align_corners=True)
    gy = torch.randn(ref.shape, generator=gen)
    ref.backward(gy)
    vd = vecs.detach().to(d).clone().requires_grad_(True)
    md = masks.detach().to(d).clone().requires_grad_(True)
    out = L.masks_to_layout(vd, bx.to(d), md, o2i.to(d), 24, 40, num_imgs=3)
    assert rel_err(out, ref) < TOL
    out.backward(gy.to(d))
    assert rel_err(vd.grad, vecs.grad) < TOL and rel_err(md.grad, masks.grad) < TOL
  finally:
    L.ALIGN_CORNERS = False


@pytest.mark.parametrize('math', ['fp32', 'bf16x3', 'tf32'])
def test_eval_bn_folding_sheep(math):
  """Inference with eval-mode BatchNorm folded into the convolutions
  (crn.FOLD_EVAL_BN) against the reference's config-1 outputs."""
  import copy
  from sg2im_b200 import crn, ops
  from sg2im_b200.model import Sg2ImModel
  s = load_golden('sheep.pt')
  ops.set_conv_math(math)
  crn.FOLD_EVAL_BN = True
  try:
    with contextlib.redirect_stdout(io.StringIO()):
      m = Sg2ImModel(vocab=s['vocab'], **s['kwargs'])
    m.load_state_dict(s['sd'] if 'sd' in s else s['sd_g'])
    m.to(dev()).eval()
    kw = s['kwargs']
    torch.manual_seed(s['noise_seed'])
    noise = torch.randn(7, kw['layout_noise_dim'], 64, 64)
    objs, triples, o2i = m.encode_scene_graphs(copy.deepcopy(s['scene_graphs']))
    with torch.no_grad():
      out = m(objs, triples, o2i, noise=noise.to(dev()))
    tol = 1e-2 if math == 'tf32' else (TOL if math == 'fp32' else 1e-3)   # tf32: the labelled fast mode
    for a, b in zip(out, s['out']):
      assert rel_err(a, b) < tol
  finally:
    crn.FOLD_EVAL_BN = False
    ops.set_conv_math('fp32')


@pytest.mark.parametrize('N,H,W,C,up,extra', [
    (4, 16, 16, 64, 1, 0), (4, 16, 16, 64, 2, 40), (2, 8, 8, 1024, 2, 160), (3, 5, 7, 12, 1, 0),
    (32, 64, 64, 128, 2, 160), (32, 128, 128, 64, 1, 0)])
def test_bn_backward_v2_matches_v1_and_torch(N, H, W, C, up, extra):
  """Second-generation BatchNorm+LeakyReLU(+x2 upsample) backward kernels
  (csrc/norm_act_v2.cu, SG2IM_BNBWD_V2=1) against the first generation and, for
  the small cases, against torch autograd on the CPU."""
  import torch.nn as nn
  import torch.nn.functional as F
  from sg2im_b200 import ops
  g = torch.Generator().manual_seed(N * 1000 + C)
  x = torch.randn(N, H, W, C, generator=g)
  gy = torch.randn(N, H * up, W * up, C + extra, generator=g)
  gy[..., :extra] = 0                                  # only the written slice carries gradient

  def run(v2):
    if v2:
      os.environ['SG2IM_BNBWD_V2'] = '1'
    else:
      os.environ['SG2IM_BNBWD_V2'] = '0'                # the first-generation kernel
    try:
      bn = nn.BatchNorm2d(C).to(dev())
      with torch.no_grad():
        bn.weight.copy_(torch.linspace(0.5, 1.5, C)); bn.bias.copy_(torch.linspace(-0.2, 0.3, C))
      xd = x.to(dev()).clone().requires_grad_(True)
      out = torch.zeros(N, H * up, W * up, C + extra, device=dev()) if extra else None
      y = ops.bn_act(xd, bn, 0.2, up=up, out=out, out_coff=extra)
      y.backward(gy.to(dev()))
      return xd.grad.cpu(), bn.weight.grad.cpu(), bn.bias.grad.cpu()
    finally:
      os.environ.pop('SG2IM_BNBWD_V2', None)

  a, b = run(False), run(True)
  for u, v in zip(a, b):
    assert rel_err(v, u) < 1e-5
  if N * H * W * C <= 1 << 20:
    bn = nn.BatchNorm2d(C)
    with torch.no_grad():
      bn.weight.copy_(torch.linspace(0.5, 1.5, C)); bn.bias.copy_(torch.linspace(-0.2, 0.3, C))
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.leaky_relu(bn(xr), 0.2)
    if up > 1:
      yr = F.interpolate(yr, scale_factor=up, mode='nearest')
    yr.backward(gy[..., extra:].permute(0, 3, 1, 2))
    assert rel_err(b[0], xr.grad.permute(0, 2, 3, 1)) < TOL
    assert rel_err(b[1], bn.weight.grad) < TOL and rel_err(b[2], bn.bias.grad) < TOL


@pytest.mark.parametrize('O,N,D,M,H,W,with_masks', [
    (12, 3, 16, 8, 24, 20, True), (40, 4, 128, 16, 64, 64, True), (40, 4, 128, 16, 64, 64, False),
    (320, 32, 128, 16, 128, 128, True), (7, 2, 256, 5, 9, 37, True)])
def test_layout_backward_v2(O, N, D, M, H, W, with_masks):
  """Tiled layout backward (SG2IM_LAYOUT_V2=1) against the first generation and,
  for the small cases, the oracle's autograd."""
  from sg2im_b200 import layout as L
  from oracle import sg2im_oracle as orc
  g = torch.Generator().manual_seed(O + D)
  vecs = torch.randn(O, D, generator=g)
  xy = torch.rand(O, 2, generator=g) * 0.6
  boxes = torch.cat([xy, xy + torch.rand(O, 2, generator=g) * 0.35 + 0.1], 1)
  o2i = torch.sort(torch.randint(0, N, (O,), generator=g)).values
  boxes[-1] = torch.tensor([0., 0., 1., 1.])
  if O > 64:
    boxes[0] = torch.tensor([0.2, 0.3, 0.2, 0.5])         # degenerate width: contributes nothing
  masks = torch.rand(O, M, M, generator=g) if with_masks else None
  gy = torch.randn(N, D, H, W, generator=g)
  d = dev()

  def run(v2):
    if v2:
      os.environ['SG2IM_LAYOUT_V2'] = '1'
    else:
      os.environ['SG2IM_LAYOUT_V2'] = '0'                # the first-generation kernel
    try:
      vd = vecs.to(d).clone().requires_grad_(True)
      if with_masks:
        md = masks.to(d).clone().requires_grad_(True)
        out = L.masks_to_layout(vd, boxes.to(d), md, o2i.to(d), H, W, num_imgs=N)
      else:
        md = None
        out = L.boxes_to_layout(vd, boxes.to(d), o2i.to(d), H, W, num_imgs=N)
      out.backward(gy.to(d))
      return vd.grad.cpu(), (md.grad.cpu() if md is not None else None)
    finally:
      os.environ.pop('SG2IM_LAYOUT_V2', None)

  a, b = run(False), run(True)
  assert rel_err(b[0], a[0]) < 1e-5
  if with_masks:
    assert rel_err(b[1], a[1]) < 1e-5
  if O <= 64:
    vr = vecs.clone().requires_grad_(True)
    if with_masks:
      mr = masks.clone().requires_grad_(True)
      ref = orc.masks_to_layout(vr, boxes, mr, o2i, H, W, N)
    else:
      ref = orc.boxes_to_layout(vr, boxes, o2i, H, W, N)
    ref.backward(gy)
    assert rel_err(b[0], vr.grad) < TOL
    if with_masks:
      assert rel_err(b[1], mr.grad) < TOL


def test_flat_adam_kernel_matches_torch_adam():
  """sg2im_adam_flat (csrc/adam.cu) vs torch.optim.Adam on the device, incl. a
  bucket whose length is not a multiple of 4 and the found_inf skip."""
  from sg2im_b200 import ops
  for n in (1027, 4096, 28135695 // 7):
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g)
    ref = p0.clone().to(dev()).requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    p = p0.clone().to(dev())
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros((), device=dev())
    for it in range(4):
      grad = torch.randn(n, generator=g).to(dev())
      ref.grad = grad.clone()
      opt.step()
      ops.adam_flat(p, grad, m, v, step, 1e-3, 0.9, 0.999, 1e-8)
      assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-7), (n, it)
    assert float(step) == 4
    before = p.clone()
    ops.adam_flat(p, grad, m, v, step, 1e-3, 0.9, 0.999, 1e-8, found_inf=torch.ones((), device=dev()))
    torch.cuda.synchronize()
    assert torch.equal(p, before) and float(step) == 4


@pytest.mark.parametrize('graph', [False, True])
def test_train_step_with_flat_adam(graph):
  from sg2im_b200 import ops
  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator
  from sg2im_b200.train_step import TrainStep
  ops.set_conv_math('fp32')
  g = load_golden('train_step.pt')
  kw = g['kwargs']
  with contextlib.redirect_stdout(io.StringIO()):
    m = Sg2ImModel(vocab=g['vocab'], **kw)
    d_img = PatchDiscriminator(arch=g['arch'], normalization='batch', activation='leakyrelu-0.2',
                               padding='valid')
    d_obj = AcCropDiscriminator(vocab=g['vocab'], arch=g['arch'], normalization='batch',
                                activation='leakyrelu-0.2', padding='valid', object_size=g['crop'])
  m.load_state_dict(g['sd_g']); d_img.load_state_dict(g['sd_img']); d_obj.load_state_dict(g['sd_obj'])
  for net in (m, d_img, d_obj):
    net.to(dev())
  step = TrainStep(m, d_obj, d_img, fused_adam='flat', cuda_graph=graph, graph_warmup=1)
  batch = [t.to(dev()) for t in g['batch']]
  N = batch[0].size(0)
  for it, seed in enumerate(g['noise_seeds']):
    torch.manual_seed(seed)
    noise = torch.randn(N, kw['layout_noise_dim'], *kw['image_size']).to(dev())
    losses, _ = step.step(batch, noise=noise)
    for k, v in g['losses'][it].items():
      assert abs(losses[k] - v) <= 1e-3 * max(1.0, abs(v)), (it, k, losses[k], v)
  sd = m.state_dict()
  for k, v in g['sd_g_after'].items():
    if v.dtype.is_floating_point:
      assert (sd[k].cpu() - v).abs().max() < 2.5e-4, k


@pytest.mark.parametrize('O,N,D,M,H,W,nc', [
    (12, 3, 16, 8, 24, 20, 0), (40, 4, 128, 16, 64, 64, 32), (320, 32, 128, 16, 128, 128, 32),
    (45, 2, 128, 16, 32, 32, 40), (9, 3, 64, 0, 16, 48, 5), (7, 2, 256, 20, 9, 37, 0)])
def test_layout_forward_v2_bit_identical(O, N, D, M, H, W, nc):
  """Band-resident layout forward (SG2IM_LAYOUT_V2=1): same summation order as the
  first generation, so the outputs must be bit-identical — incl. images with more
  than 16 objects (several passes), no masks (M=0), masks too large for shared
  memory (M=20) and noise channels."""
  from sg2im_b200 import ops
  g = torch.Generator().manual_seed(O * 7 + D)
  vecs = torch.randn(O, D, generator=g).to(dev())
  xy = torch.rand(O, 2, generator=g) * 0.6
  boxes = torch.cat([xy, xy + torch.rand(O, 2, generator=g) * 0.35 + 0.1], 1).to(dev())
  o2i = torch.sort(torch.randint(0, N, (O,), generator=g)).values
  if O == 45:
    o2i = torch.cat([torch.zeros(40, dtype=torch.int64), torch.ones(5, dtype=torch.int64)])
  o2i = o2i.to(dev())
  masks = torch.rand(O, M, M, generator=g).to(dev()) if M else None
  noise = torch.randn(N, nc, H, W, generator=g).to(dev()) if nc else None
  outs = []
  for v2 in (False, True):
    if v2:
      os.environ['SG2IM_LAYOUT_V2'] = '1'
    else:
      os.environ['SG2IM_LAYOUT_V2'] = '0'                # the first-generation kernel
    try:
      for math in ('fp32', 'tf32'):
        ops.set_conv_math(math)                            # tf32: the stack variant rounds its output
        outs.append(ops.Layout.apply(vecs, boxes, masks, o2i, N, H, W, noise, False))
        extras = [0, 8]
        if H % 2 == 0 and W % 2 == 0:
          outs.append(ops.LayoutStack.apply(vecs, boxes, masks, o2i, N, H, W, noise, False, extras)[1]
                      [..., :D + nc].contiguous())
    finally:
      os.environ.pop('SG2IM_LAYOUT_V2', None)
      ops.set_conv_math('fp32')
  half = len(outs) // 2
  for a, b in zip(outs[:half], outs[half:]):
    assert torch.equal(a, b)


@pytest.mark.parametrize('N,H,W,C,up,extra,use_bn', [
    (4, 16, 16, 64, 1, 0, True), (4, 16, 16, 64, 2, 40, True), (2, 8, 8, 1024, 2, 160, True),
    (3, 5, 7, 12, 1, 0, False), (3, 5, 7, 12, 2, 4, False), (32, 64, 64, 128, 2, 160, True)])
def test_scale_act_forward_v2_bit_identical(N, H, W, C, up, extra, use_bn):
  """Input-stationary BN-apply + LeakyReLU + x2 upsample forward (SG2IM_BNFWD_V2=1):
  same per-element operations as the first generation => identical bits."""
  import torch.nn as nn
  from sg2im_b200 import ops
  x = torch.randn(N, H, W, C, generator=torch.Generator().manual_seed(C + up)).to(dev())
  outs = []
  for v2 in (False, True):
    if v2:
      os.environ['SG2IM_BNFWD_V2'] = '1'
    else:
      os.environ['SG2IM_BNFWD_V2'] = '0'                # the first-generation kernel
    try:
      for math in ('fp32', 'tf32'):
        ops.set_conv_math(math)
        bn = None
        if use_bn:
          bn = nn.BatchNorm2d(C).to(dev())
          with torch.no_grad():
            bn.weight.copy_(torch.linspace(0.5, 1.5, C)); bn.bias.copy_(torch.linspace(-0.2, 0.3, C))
        out = torch.full((N, H * up, W * up, C + extra), 7.0, device=dev()) if extra else None
        with torch.no_grad():
          outs.append(ops.bn_act(x, bn, 0.2, up=up, out=out, out_coff=extra).clone())
    finally:
      os.environ.pop('SG2IM_BNFWD_V2', None)
      ops.set_conv_math('fp32')
  for a, b in zip(outs[:2], outs[2:]):
    assert torch.equal(a, b)
  if extra:
    assert bool((outs[2][..., :extra] == 7.0).all())       # channels in front of the slice untouched


@pytest.mark.parametrize('M,C', [(448, 1152), (320, 128), (1, 5), (8192, 64), (37, 179), (9000, 32)])
def test_colsum_small_kernel(M, C):
  """Single-launch column sum for the small GEMMs' bias gradients (SG2IM_COLSUM_V2=1)."""
  from sg2im_b200 import ops
  x = torch.randn(M, C, generator=torch.Generator().manual_seed(M + C))
  ref = x.double().sum(dim=0).float()
  got = ops.colsum(x.to(dev())).cpu()
  os.environ['SG2IM_COLSUM_V2'] = '0'                     # the three-launch first generation
  try:
    base = ops.colsum(x.to(dev())).cpu()
  finally:
    os.environ.pop('SG2IM_COLSUM_V2', None)
  assert torch.allclose(got, ref, rtol=1e-6, atol=1e-5)
  assert torch.allclose(base, ref, rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize('adam,graph', [(None, False), ('flat', False), (None, True)])
def test_train_step_with_weights_in_the_gradient_layout(adam, graph):
  """TrainStep(weights='kcc') on the tensor-core path: forward reads [KH][KW][Cin][Cout] weights
  MN-major, dgrad K-major with flipped taps, wgrad lands in place — no pack / unpack kernels."""
  from sg2im_b200 import ops
  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator
  from sg2im_b200.train_step import TrainStep
  g = load_golden('train_step.pt')
  kw = g['kwargs']
  with contextlib.redirect_stdout(io.StringIO()):
    m = Sg2ImModel(vocab=g['vocab'], **kw)
    d_img = PatchDiscriminator(arch=g['arch'], normalization='batch', activation='leakyrelu-0.2',
                               padding='valid')
    d_obj = AcCropDiscriminator(vocab=g['vocab'], arch=g['arch'], normalization='batch',
                                activation='leakyrelu-0.2', padding='valid', object_size=g['crop'])
  m.load_state_dict(g['sd_g']); d_img.load_state_dict(g['sd_img']); d_obj.load_state_dict(g['sd_obj'])
  for net in (m, d_img, d_obj):
    net.to(dev())
  ops.set_conv_math('tf32')
  try:
    step = TrainStep(m, d_obj, d_img, weights='kcc', fused_adam=adam, cuda_graph=graph, graph_warmup=1)
    batch = [t.to(dev()) for t in g['batch']]
    N = batch[0].size(0)
    for it, seed in enumerate(g['noise_seeds']):
      torch.manual_seed(seed)
      noise = torch.randn(N, kw['layout_noise_dim'], *kw['image_size']).to(dev())
      losses, _ = step.step(batch, noise=noise)
      for k, v in g['losses'][it].items():
        assert abs(losses[k] - v) / max(1.0, abs(v)) <= 1e-2, (it, k, losses[k], v)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g['sd_g_after'].keys())
    for k, v in g['sd_g_after'].items():
      if v.dtype.is_floating_point:
        assert (sd[k].cpu() - v).abs().max() < 2e-3, k
  finally:
    ops.set_conv_math('fp32')


def test_fused_activation_backward_bias_gradient():
  """sg2im_act_bwd_colsum (SG2IM_ACTBWD_FUSED=1) vs act_bwd + colsum."""
  from sg2im_b200 import ops
  g = torch.Generator().manual_seed(5)
  for M, C in ((1000, 64), (37, 132), (32 * 128 * 128, 64)):
    dy, y = torch.randn(M, C, generator=g).to(dev()), torch.randn(M, C, generator=g).to(dev())
    ref_dx = ops.act_bwd(dy, y, 0.2)
    ref_db = ops.colsum(ref_dx)
    bias = torch.zeros(C, device=dev(), requires_grad=True)
    ops.FUSE_ACT_BWD = True
    try:
      dx, db = ops.act_bwd_bias(dy, y, 0.2, bias)
    finally:
      ops.FUSE_ACT_BWD = False
    assert torch.equal(dx, ref_dx)
    assert torch.allclose(db, ref_db, rtol=1e-4, atol=1e-3)


def test_coco_relations_match_the_reference_samples_and_scale():
  """sg2im_coco_relations (COCO scene-graph synthesis): the reference's own
  samples (tests/golden/coco_rel.pt) collated into one batch, against the pinned oracle; then a
  training-size batch (32 images x 7 objects) checked through properties."""
  import random
  from oracle import relations_oracle as RO
  from sg2im_b200 import batching
  d = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'coco_rel.pt'))
  names = d['pred_names']
  idx = {n: i for i, n in enumerate(names)}
  vocab = {'object_name_to_idx': {'__image__': 0}, 'pred_name_to_idx': idx}
  for mask_size in (16, 5):
    samples = [s for s in d['samples'] if s['masks'].size(1) == mask_size]
    rng = random.Random(4)
    want, base = [], 0
    for s in samples:
      t, _ = RO.sample_triples(s['objs'], s['boxes'], s['masks'].long(), idx, rng=rng)
      want += [[a + base, p, b + base] for a, p, b in t]
      base += s['objs'].numel()
    boxes = torch.cat([s['boxes'] for s in samples])
    masks = torch.cat([s['masks'].long() for s in samples])
    counts = [s['objs'].numel() for s in samples]
    triples, t2i, o2i = batching.coco_relations(boxes, masks, counts, vocab, rng=random.Random(4), device=dev())
    assert triples.device == dev() or triples.device.type == dev().type
    assert torch.equal(triples.cpu(), torch.tensor(want))   # golden margins are >= 3e-2
    assert torch.equal(t2i.cpu(), torch.repeat_interleave(
        torch.arange(len(counts)), torch.tensor([2 * (c - 1) if c > 2 else c - 1 for c in counts])))
  # training-size batch: structure of the table
  g = torch.Generator().manual_seed(0)
  N, c = 32, 7
  xy = torch.rand(N * c, 2, generator=g) * 0.6
  boxes = torch.cat([xy, xy + torch.rand(N * c, 2, generator=g) * 0.4], 1)
  masks = (torch.rand(N * c, 16, 16, generator=g) < 0.5).long()
  triples, t2i, o2i = batching.coco_relations(boxes.to(dev()), masks.to(dev()), [c] * N, vocab,
                                              rng=random.Random(1), device=dev())
  triples, t2i, o2i = triples.cpu(), t2i.cpu(), o2i.cpu()
  assert triples.shape == (N * 2 * (c - 1), 3)
  assert torch.equal(o2i[triples[:, 0]], t2i) and torch.equal(o2i[triples[:, 2]], t2i)   # never across images
  in_img = triples[:, 1] == idx['__in_image__']
  assert int(in_img.sum()) == N * (c - 1) and bool((triples[in_img, 2] % c == c - 1).all())
  assert bool((triples[~in_img, 0] != triples[~in_img, 2]).all())


@pytest.mark.parametrize('N,H,W,Ci,Cf,Co,K,P', [
    (32, 8, 8, 160, 1184, 1024, 3, 1),      # CRN stage 0 conv 1: channel prefix of a wider weight, per-tap kernel
    (8, 32, 32, 256, 256, 256, 3, 1),       # BN=128/256 per-tap tiles
    (4, 128, 128, 288, 288, 64, 3, 1),      # halo kernel (weight-stationary)
    (4, 64, 64, 96, 96, 128, 2, 0),         # 2x2 taps (the space-to-depth discriminator convs)
    (448, 1, 1, 384, 384, 512, 1, 0),       # linear layer as a 1x1 convolution
])
def test_conv_from_weight_gradient_layout_matches_packed(N, H, W, Ci, Cf, Co, K, P):
  """sg2im_conv_tc_kcc (WMODE 1 forward: MN-major B operand; WMODE 2 data gradient: tap flip) vs
  the validated kernels on packed weights.  Same A operand, same products, same accumulation
  order => identical bits; the weights are RN-TF32 in both."""
  from sg2im_b200 import ops
  ops.set_conv_math('tf32')
  try:
    g = torch.Generator().manual_seed(Ci + Co + K)
    x = torch.randn(N, H, W, Ci, generator=g).to(dev())
    w_full = (torch.randn(Co, Cf, K, K, generator=g) * 0.05).to(dev())
    b = torch.randn(Co, generator=g).to(dev())
    Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
    gy = torch.randn(N, Ho, Wo, Co, generator=g).to(dev())
    kcc = w_full.permute(2, 3, 1, 0).reshape(K * K, Cf, Co).contiguous()
    ops.round_tf32(kcc, kcc)
    y_ref = ops.conv_tc(x, ops.pack_tc_fwd(w_full, Ci), b, K, K, P, Co, 1, 0.2)
    y = ops.conv_tc_kcc(x, kcc, Cf, False, b, K, K, P, Co, 1, 0.2)
    assert torch.equal(y, y_ref)
    if K - 1 - P >= 0:
      dx_ref = ops.conv_tc(gy, ops.pack_tc_dgrad(w_full, Ci), None, K, K, K - 1 - P, Ci, out_hw=(H, W))
      dx = ops.conv_tc_kcc(gy, kcc, Cf, True, None, K, K, K - 1 - P, Ci, out_hw=(H, W))
      assert torch.equal(dx, dx_ref)
    torch.cuda.synchronize()
  finally:
    ops.set_conv_math('fp32')


# ---- factory-surface rows a9 / a10: normalization='instance', build_cnn's 'R'
# and 'PX' tokens.  Instance normalisation and the residual block reuse hardware-validated kernels
# (per-image BatchNorm statistics / apply, convolutions); csrc/pool.cu is new device code.

@pytest.mark.parametrize('N,H,W,C,f', [(2, 8, 8, 8, 2), (3, 13, 9, 6, 3), (1, 4, 4, 4, 4),
                                       (2, 7, 10, 5, 2), (2, 6, 6, 12, 1)])
@pytest.mark.parametrize('mode', [0, 1])
def test_pool2d_forward_backward_vs_torch(N, H, W, C, f, mode):
  """sg2im_pool2d_fwd / _bwd vs F.avg_pool2d / F.max_pool2d (kernel = stride = f, floor mode):
  ragged sizes, channel counts off the float4 path, and — for max — exact ties (first element in
  row-major window order wins, like ATen)."""
  import torch.nn.functional as F
  from sg2im_b200 import ops
  g = torch.Generator().manual_seed(N * 100 + H * 10 + f + mode)
  x = torch.randn(N, C, H, W, generator=g)
  if mode == 1:
    x = (x * 2).round() / 2                           # plenty of exact ties inside windows
  xr = x.clone().requires_grad_(True)
  want = (F.avg_pool2d if mode == 0 else F.max_pool2d)(xr, f, f)
  wgt = torch.randn(want.shape, generator=g)
  (want * wgt).sum().backward()
  h = x.permute(0, 2, 3, 1).contiguous().to(dev()).requires_grad_(True)
  got = ops.Pool2d.apply(h, f, mode)
  (got * wgt.permute(0, 2, 3, 1).to(dev())).sum().backward()
  assert got.shape == (N, H // f, W // f, C)
  if mode == 1:
    assert torch.equal(got.detach().cpu().permute(0, 3, 1, 2), want.detach())
    assert torch.equal(h.grad.cpu().permute(0, 3, 1, 2), xr.grad)
  else:
    assert rel_err(got.detach().cpu().permute(0, 3, 1, 2), want.detach()) < 1e-6
    assert rel_err(h.grad.cpu().permute(0, 3, 1, 2), xr.grad) < 1e-6


def test_pool2d_refuses_empty_output():
  from sg2im_b200 import ops
  with pytest.raises(RuntimeError):
    ops.Pool2d.apply(torch.zeros(1, 2, 2, 4, device=dev()), 3, 1)


@pytest.mark.parametrize('up,slope,sliced', [(1, 1.0, False), (1, 0.2, False), (2, 0.2, True)])
def test_instance_norm_vs_oracle(up, slope, sliced):
  """layers.InstanceNorm2d / layers.norm_act(+ LeakyReLU, + nearest x2 upsample into a channel
  slice, as the CRN uses it) vs oracle.instancenorm2d; forward and gradient."""
  import torch.nn.functional as F
  from oracle import sg2im_oracle as orc
  from sg2im_b200.layers import InstanceNorm2d, norm_act
  g = torch.Generator().manual_seed(7)
  N, H, W, C, coff = 3, 6, 5, 8, 4
  x = torch.randn(N, C, H, W, generator=g) * 2 + 0.5
  xr = x.clone().requires_grad_(True)
  want = F.leaky_relu(orc.instancenorm2d(xr), slope) if slope != 1.0 else orc.instancenorm2d(xr)
  if up > 1:
    want = F.interpolate(want, scale_factor=up, mode='nearest')
  wgt = torch.randn(want.shape, generator=g)
  (want * wgt).sum().backward()
  norm = InstanceNorm2d(C)
  assert len(norm.state_dict()) == 0
  for training in (True, False):                      # instance statistics in both modes
    norm.train(training)
    h = x.permute(0, 2, 3, 1).contiguous().to(dev()).requires_grad_(True)
    if sliced:
      buf = torch.full((N, H * up, W * up, coff + C + 4), 3.0, device=dev())
      out = norm_act(h, norm, slope, up=up, out=buf, out_coff=coff)
      got = out[..., coff:coff + C]
      assert bool((out[..., :coff] == 3.0).all()) and bool((out[..., coff + C:] == 3.0).all())
    else:
      got = norm_act(h, norm, slope, up=up)
    (got * wgt.permute(0, 2, 3, 1).to(dev())).sum().backward()
    assert rel_err(got.detach().cpu().permute(0, 3, 1, 2), want.detach()) < TOL
    assert rel_err(h.grad.cpu().permute(0, 3, 1, 2), xr.grad) < TOL


@pytest.mark.parametrize('arch,norm,pool,size', [('R,C3-8,R,P2,R', 'batch', 'max', 8),
                                                 ('I4,C3-4,R,P3', 'instance', 'avg', 9)])
def test_build_cnn_residual_pool_instance_vs_torch(arch, norm, pool, size):
  """build_cnn with 'R' / 'PX' tokens and instance normalisation against the same network spelled
  with torch.nn.functional on the CPU (sg2im/layers.py:89-117,129-213), forward, input gradient,
  running statistics (the residual block evaluates its body twice per forward in train mode)."""
  import torch.nn.functional as F
  from oracle import sg2im_oracle as orc
  from sg2im_b200.layers import build_cnn, ResidualBlock
  torch.manual_seed(3)
  with contextlib.redirect_stdout(io.StringIO()):
    net, _ = build_cnn(arch, normalization=norm, activation='leakyrelu-0.2', padding='same',
                       pooling=pool)
  sd0 = {k: v.clone() for k, v in net.state_dict().items()}
  net = net.to(dev()).train()
  cin = int(arch[1]) if arch.startswith('I') else 3
  x = torch.randn(2, cin, size, size)

  def run_ref(mods, sd, prefix, t, twice=False):
    for i, m in enumerate(mods):
      key = '%s%d' % (prefix, i)
      if isinstance(m, ResidualBlock):
        if any(isinstance(c, torch.nn.BatchNorm2d) for c in m.net):
          run_ref(list(m.net), sd, key + '.net.', t.detach())          # discarded first evaluation
        t = t + run_ref(list(m.net), sd, key + '.net.', t)
      elif isinstance(m, torch.nn.Conv2d):
        t = F.conv2d(t, sd[key + '.weight'], sd[key + '.bias'], stride=m.stride, padding=m.padding)
      elif isinstance(m, torch.nn.BatchNorm2d):
        t = orc.batchnorm2d(sd, key, t, True)
      elif isinstance(m, torch.nn.InstanceNorm2d):
        t = orc.instancenorm2d(t)
      elif isinstance(m, torch.nn.LeakyReLU):
        t = F.leaky_relu(t, m.negative_slope)
      elif isinstance(m, torch.nn.MaxPool2d):
        t = F.max_pool2d(t, m.kernel_size, m.stride)
      elif isinstance(m, torch.nn.AvgPool2d):
        t = F.avg_pool2d(t, m.kernel_size, m.stride)
      else:
        raise AssertionError(type(m))
    return t

  xr = x.clone().requires_grad_(True)
  sd = {k: v.clone() for k, v in sd0.items()}
  want = run_ref(list(net), sd, '', xr)
  wgt = torch.randn(want.shape)
  (want * wgt).sum().backward()
  xg = x.to(dev()).requires_grad_(True)
  got = net(xg)
  (got * wgt.to(dev())).sum().backward()
  assert rel_err(got.detach().cpu(), want.detach()) < TOL
  assert rel_err(xg.grad.cpu(), xr.grad) < TOL
  for k, v in net.state_dict().items():
    if 'running' in k or 'num_batches' in k:
      assert rel_err(v.cpu().float(), sd[k].float()) < TOL, k


def _with_math(mode, fn):
  from sg2im_b200 import ops
  ops.set_conv_math(mode)
  try:
    return fn()
  finally:
    ops.set_conv_math('fp32')


# ---- gradient w.r.t. the layout boxes (csrc/layout_boxes.cu): the generator trained on its own
# predicted boxes (Sg2ImModel.forward without boxes_gt, sg2im/model.py:151-160)

@pytest.mark.parametrize('with_masks,align,O,N,D,M,H,W,nc', [
    (True, False, 9, 3, 16, 8, 24, 20, 0), (False, False, 7, 2, 12, 0, 16, 16, 4),
    (True, True, 6, 2, 8, 5, 12, 18, 0), (True, False, 5, 2, 132, 16, 32, 32, 0)])
def test_layout_gradient_wrt_boxes(with_masks, align, O, N, D, M, H, W, nc):
  from oracle import sg2im_oracle as orc
  from sg2im_b200 import ops
  g = torch.Generator().manual_seed(O * 7 + D)
  vecs = torch.randn(O, D, generator=g)
  xy = torch.rand(O, 2, generator=g) * 0.5
  boxes = torch.cat([xy, xy + 0.15 + 0.35 * torch.rand(O, 2, generator=g)], 1)
  boxes[0] = torch.tensor([-0.1, 0.2, 0.7, 1.3])              # partly outside the image
  masks = torch.rand(O, M, M, generator=g) if with_masks else None
  o2i = torch.sort(torch.randint(0, N, (O,), generator=g)).values
  noise = torch.randn(N, nc, H, W, generator=g) if nc else None
  br = boxes.clone().requires_grad_(True)
  vr = vecs.clone().requires_grad_(True)
  if with_masks:
    want = orc.masks_to_layout(vr, br, masks, o2i, H, W, N, align_corners=align)
  else:
    want = orc.boxes_to_layout(vr, br, o2i, H, W, N, align_corners=align)
  wgt = torch.randn(want.shape, generator=g)
  (want * wgt).sum().backward()
  bm = boxes.clone().to(dev()).requires_grad_(True)
  vm = vecs.clone().to(dev()).requires_grad_(True)
  got = ops.Layout.apply(vm, bm, None if masks is None else masks.to(dev()), o2i.to(dev()), N, H, W,
                         None if noise is None else noise.to(dev()), align)
  (got[..., :D] * wgt.permute(0, 2, 3, 1).to(dev())).sum().backward()
  assert rel_err(got[..., :D].detach().cpu().permute(0, 3, 1, 2), want.detach()) < TOL
  assert rel_err(vm.grad.cpu(), vr.grad) < TOL
  assert rel_err(bm.grad.cpu(), br.grad) < 5 * TOL, (bm.grad.cpu(), br.grad)


def test_generator_trains_on_predicted_boxes():
  """No boxes_gt: the layout is built from boxes_pred and the image loss reaches box_net and the
  graph convolution through the sampling grid (model.py:151-160).  Parameter gradients vs the
  oracle's autograd."""
  import test_gpu_model as G
  from oracle import sg2im_oracle as orc
  g = load_golden('generator.pt')
  imgs, objs, boxes, triples, o2i, _ = g['batch']
  kw = g['kwargs']
  N = imgs.size(0)
  noise = G._noise(5, N, kw['layout_noise_dim'], kw['image_size'])
  sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and 'running' not in k
            else v.clone()) for k, v in g['sd'].items()}
  ref = orc.generator_forward(sd, kw['image_size'], objs, triples, o2i, boxes_gt=None, noise=noise,
                              training=True, num_imgs=N)
  wimg = torch.randn(ref[0].shape, generator=torch.Generator().manual_seed(9))
  (ref[0] * wimg).sum().backward()
  m = G._build_generator(g)
  if m.box_net[0].weight.device != dev():
    m = m.to(dev())
  m.train()
  d = dev()
  out = m(objs.to(d), triples.to(d), o2i.to(d), boxes_gt=None, noise=noise.to(d), num_imgs=N)
  (out[0] * wimg.to(d)).sum().backward()
  assert rel_err(out[0].detach().cpu(), ref[0].detach()) < TOL
  seen_box = False
  for k, p in m.named_parameters():
    rg = sd[k].grad
    if rg is None or '.net.0.bias' in k or '.net.3.bias' in k:
      continue
    if k.startswith('box_net'):
      seen_box = seen_box or float(rg.abs().max()) > 0
    assert rel_err(p.grad.cpu(), rg) < 10 * TOL, (k, rel_err(p.grad.cpu(), rg))
  assert seen_box                                  # box_net is reached ONLY through the layout here


def test_cuda_graph_cache_is_bounded_for_variable_batch_shapes():
  """Real VG / COCO loaders change the object / triple counts almost every iteration.  The graph
  cache of TrainStep is LRU-bounded and a signature is captured only after it was seen twice:
  one-off shapes run eagerly, recurring ones replay, and evicted graphs release their memory."""
  import test_gpu_model as G
  from sg2im_b200.synth import synth_batch
  from sg2im_b200.train_step import TrainStep
  g = load_golden('train_step.pt')
  m, d_obj, d_img = G._build_all(g)
  kw = g['kwargs']
  H, W = kw['image_size']
  N = g['batch'][0].size(0)
  step = TrainStep(m, d_obj, d_img, cuda_graph=True, graph_warmup=1, max_graphs=2, graph_min_seen=2,
                   capture_cooldown=4)
  def batch(objs, rels, seed):
    return [t.to(dev()) for t in synth_batch(N=N, objs_per_img=objs, rels_per_img=rels, image_size=(H, W),
                                             num_objs=9, num_preds=5, seed=seed)]
  shapes = [(3, 2), (2, 1), (4, 3), (3, 1)]
  torch.cuda.synchronize()
  hist = []
  for it in range(24):
    o, r = shapes[it % 3] if it < 18 else shapes[3]            # three recurring signatures, then a fourth
    noise = G._noise(100 + it, N, kw['layout_noise_dim'], (H, W)).to(dev())
    losses, imgs = step.step(batch(o, r, it), noise=noise)
    assert all(v == v for v in losses.values())
    hist.append((len(step._graphs), step.replays))
  assert max(h[0] for h in hist) <= 2                          # never more than max_graphs captured graphs
  assert step.graph_evictions >= 1 and step.replays >= 4       # recurring shapes replay; LRU evicts
  # a shape seen once runs eagerly and is not captured
  before = len(step._graphs), step.replays
  step.step(batch(1, 1, 999), noise=G._noise(5, N, kw['layout_noise_dim'], (H, W)).to(dev()))
  assert (len(step._graphs), step.replays) == before
  # a loader cycling through more signatures than the cache holds: evicting captures are rate-limited
  # (capture_cooldown), the rest of the steps run eagerly instead of re-capturing every time
  ev0 = step.graph_evictions
  for it in range(12):
    o, r = shapes[it % 4]
    step.step(batch(o, r, 500 + it), noise=G._noise(7, N, kw['layout_noise_dim'], (H, W)).to(dev()))
  assert step.graph_evictions - ev0 <= 12 // 4 + 1


@pytest.mark.parametrize('shape,target', [((320, 1), 1.0), ((32, 256, 14, 14), 0.0), ((5,), 1.0)])
def test_fused_gan_bce_loss_matches_the_reference_composition(shape, target):
  """ops.BCELogitsMean (csrc/loss.cu) — what gan_g_loss / gan_d_loss launch on CUDA tensors — vs
  bce_loss(scores, ones / zeros) composed from torch ops as in sg2im/losses.py:39-57, value and
  gradient (through a non-trivial upstream factor)."""
  from sg2im_b200 import ops
  from sg2im_b200.losses import bce_loss, gan_d_loss
  g = torch.Generator().manual_seed(len(shape))
  x = (torch.randn(*shape, generator=g) * 3).to(dev())
  xr = x.clone().requires_grad_(True)
  ref = bce_loss(xr.reshape(-1), torch.full((x.numel(),), target, device=dev()))
  (ref * 0.3).backward()
  xd = x.clone().requires_grad_(True)
  out = ops.BCELogitsMean.apply(xd.reshape(-1), target)
  (out * 0.3).backward()
  assert abs(float(out) - float(ref)) <= 2e-6 * max(1.0, abs(float(ref)))
  assert rel_err(xd.grad, xr.grad) < 1e-5
  # the public loss functions route to it
  a, b = x.clone().requires_grad_(True), (x * 0.5).clone().requires_grad_(True)
  d = gan_d_loss(a, b)
  want = bce_loss(a.detach().reshape(-1), torch.ones(x.numel(), device=dev())) + \
      bce_loss(b.detach().reshape(-1), torch.zeros(x.numel(), device=dev()))
  assert abs(float(d) - float(want)) <= 4e-6 * max(1.0, abs(float(want)))


@pytest.mark.parametrize('target', [0.0, 1.0])
def test_fused_gan_bce_loss_on_saturated_and_zero_scores(target):
  """Saturated logits stay finite, and a score of exactly 0 gets the reference composition's
  subgradient (clamp(x, min=0) differentiates as 1, |x| as 0 there: 1 - t, not 0.5 - t)."""
  from sg2im_b200 import ops
  from sg2im_b200.losses import bce_loss
  x = torch.tensor([-1e4, -88.0, -30.0, -1e-8, 0.0, 1e-8, 30.0, 88.0, 1e4, 3.0e38, -3.0e38], device=dev())
  xr = x.clone().requires_grad_(True)
  ref = bce_loss(xr, torch.full_like(xr, target))
  ref.backward()
  xd = x.clone().requires_grad_(True)
  out = ops.BCELogitsMean.apply(xd, target)
  out.backward()
  assert bool(torch.isfinite(out)) and abs(float(out) - float(ref)) <= 1e-6 * abs(float(ref))
  assert bool(torch.isfinite(xd.grad).all())
  assert float((xd.grad - xr.grad).abs().max()) <= 1e-7
