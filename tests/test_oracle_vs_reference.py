"""Live cross-checks against the UNMODIFIED reference imported from
/root/reference (present only in the build container; skipped elsewhere):

  * the oracle (oracle/sg2im_oracle.py) vs the reference on configurations the
    golden fixtures do not cover (layer counts, mask sizes, 'none'
    normalisation, activation strings, no masks, no noise);
  * the nn.Module mirror (sg2im_b200.*) under the CPU op shim vs the reference,
    same configurations — same state_dict loaded into both.
"""
import contextlib
import io

import pytest
import torch

from conftest import rel_err
from refimport import have_reference, import_reference

pytestmark = pytest.mark.skipif(not have_reference(), reason='reference tree not mounted')

TOL = 1e-5

CONFIGS = [
    dict(embedding_dim=8, gconv_dim=8, gconv_hidden_dim=16, gconv_num_layers=1,
         refinement_dims=(16,), mask_size=4, layout_noise_dim=0, image_size=(16, 16)),
    dict(embedding_dim=8, gconv_dim=16, gconv_hidden_dim=16, gconv_num_layers=2,
         refinement_dims=(16, 8, 8), mask_size=None, layout_noise_dim=4, image_size=(32, 32)),
    dict(embedding_dim=12, gconv_dim=12, gconv_hidden_dim=24, gconv_num_layers=4,
         refinement_dims=(24, 16, 8, 8), mask_size=16, layout_noise_dim=8, image_size=(32, 64),
         normalization='none', activation='relu'),
    dict(embedding_dim=8, gconv_dim=8, gconv_hidden_dim=16, gconv_num_layers=0,
         refinement_dims=(8, 8), mask_size=8, layout_noise_dim=4, image_size=(16, 16),
         gconv_pooling='sum', activation='leakyrelu'),
    dict(embedding_dim=8, gconv_dim=8, gconv_hidden_dim=16, gconv_num_layers=1,
         refinement_dims=(16, 8), mask_size=8, layout_noise_dim=4, image_size=(16, 16),
         normalization='instance'),
    # SURVEY §8d C4 shape in miniature: a SIX-stage refinement network (VG-256 has one stage more
    # than the 128x128 default; check_args only needs H // 2^stages >= 1, train.py:153-158)
    dict(embedding_dim=8, gconv_dim=8, gconv_hidden_dim=16, gconv_num_layers=2,
         refinement_dims=(16, 16, 8, 8, 8, 4), mask_size=16, layout_noise_dim=4, image_size=(64, 64)),
]


def _quiet():
  return contextlib.redirect_stdout(io.StringIO())


def _batch(H, W, seed):
  from sg2im_b200.synth import synth_batch
  return synth_batch(N=3, objs_per_img=4, rels_per_img=3, image_size=(H, W), num_objs=7,
                     num_preds=4, seed=seed)


@pytest.mark.parametrize('idx', range(len(CONFIGS)))
def test_generator_oracle_and_mirror_vs_live_reference(idx):
  import_reference()
  from sg2im.model import Sg2ImModel as RefModel
  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.synth import make_vocab
  from oracle import sg2im_oracle as orc
  from cpu_shim import cpu_ops
  kw = dict(CONFIGS[idx])
  H, W = kw['image_size']
  vocab = make_vocab(7, 4)
  torch.manual_seed(100 + idx)
  with _quiet():
    ref = RefModel(vocab=vocab, **kw)
    mine = Sg2ImModel(vocab=vocab, **kw)
  with torch.no_grad():
    ref.box_net[2].bias.copy_(torch.tensor([0.1, 0.15, 0.6, 0.7]))      # finite predicted boxes
  sd = {k: v.clone() for k, v in ref.state_dict().items()}
  assert list(mine.state_dict().keys()) == list(sd.keys())
  mine.load_state_dict(sd)
  imgs, objs, boxes, triples, o2i, _ = _batch(H, W, seed=idx)
  nd = kw['layout_noise_dim']
  for training, use_gt in ((True, True), (False, False)):
    ref.train(training)
    mine.train(training)
    ref.load_state_dict(sd)
    mine.load_state_dict(sd)
    torch.manual_seed(7)
    noise = torch.randn(3, nd, H, W) if nd > 0 else None
    torch.manual_seed(7)                                 # the reference draws the same noise itself
    out_ref = ref(objs, triples, o2i, boxes_gt=boxes if use_gt else None)
    out_orc = orc.generator_forward(
        {k: v.clone() for k, v in sd.items()}, (H, W), objs, triples, o2i,
        boxes_gt=boxes if use_gt else None, noise=noise, training=training,
        activation=kw.get('activation', 'leakyrelu-0.2'),
        normalization=kw.get('normalization', 'batch'),
        gconv_pooling=kw.get('gconv_pooling', 'avg'), num_imgs=3)
    with cpu_ops():
      out_mine = mine(objs, triples, o2i, boxes_gt=boxes if use_gt else None, noise=noise,
                      num_imgs=3)
    for r, o, m in zip(out_ref, out_orc, out_mine):
      if r is None:
        assert o is None and m is None
        continue
      assert rel_err(o, r) < TOL
      assert rel_err(m, r) < TOL
    if training:
      for k, v in ref.state_dict().items():
        if 'running' in k or 'num_batches' in k:
          assert rel_err(mine.state_dict()[k].float(), v.float()) < TOL, k


@pytest.mark.parametrize('arch,norm,pad', [('C4-8-2,C4-16-2,C4-16-2', 'batch', 'valid'),
                                          ('C3-8,C3-8-2', 'none', 'same'),
                                          ('C4-8-2,C4-16-2', 'instance', 'valid'),
                                          ('C4-8-2', 'batch', 'valid')])
def test_discriminators_mirror_vs_live_reference(arch, norm, pad):
  import_reference()
  from sg2im.discriminators import PatchDiscriminator as RefP, AcCropDiscriminator as RefA
  from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator
  from sg2im_b200.synth import make_vocab
  from oracle import sg2im_oracle as orc
  from cpu_shim import cpu_ops
  vocab = make_vocab(7, 4)
  torch.manual_seed(5)
  with _quiet():
    rp = RefP(arch=arch, normalization=norm, activation='leakyrelu-0.2', padding=pad)
    ra = RefA(vocab=vocab, arch=arch, normalization=norm, activation='leakyrelu-0.2',
              padding=pad, object_size=32)
    mp = PatchDiscriminator(arch=arch, normalization=norm, activation='leakyrelu-0.2', padding=pad)
    ma = AcCropDiscriminator(vocab=vocab, arch=arch, normalization=norm,
                             activation='leakyrelu-0.2', padding=pad, object_size=32)
  assert list(mp.state_dict().keys()) == list(rp.state_dict().keys())
  assert list(ma.state_dict().keys()) == list(ra.state_dict().keys())
  mp.load_state_dict(rp.state_dict())
  ma.load_state_dict(ra.state_dict())
  imgs, objs, boxes, triples, o2i, _ = _batch(32, 32, seed=3)
  with cpu_ops():
    assert rel_err(mp(imgs), rp(imgs)) < TOL
    s_m, ac_m = ma(imgs, objs, boxes, o2i)
  s_r, ac_r = ra(imgs, objs, boxes, o2i)
  assert rel_err(s_m, s_r) < TOL and rel_err(ac_m, ac_r) < TOL
  out = orc.patch_discriminator({k: v.clone() for k, v in mp.state_dict().items()}, imgs, arch,
                                normalization=norm, padding=pad, training=True)
  # the oracle sees the running stats AFTER the two forwards above; outputs in train mode do not depend on them
  assert rel_err(out, rp(imgs)) < TOL


@pytest.mark.parametrize('dims,act,bn,final,drop', [
    ((12, 16, 8), 'relu', 'none', True, 0), ((12, 16, 8), 'leakyrelu', 'batch', True, 0),
    ((6, 10, 10, 4), 'relu', 'batch', False, 0), ((6, 10), 'relu', 'none', True, 0.5)])
def test_build_mlp_mirror_vs_live_reference(dims, act, bn, final, drop):
  """sg2im/layers.py:216-232 incl. BatchNorm1d, LeakyReLU, no final non-linearity, Dropout."""
  import_reference()
  from sg2im.layers import build_mlp as ref_build
  from sg2im_b200.layers import build_mlp
  from cpu_shim import cpu_ops
  torch.manual_seed(11)
  ref = ref_build(list(dims), activation=act, batch_norm=bn, dropout=drop, final_nonlinearity=final)
  mine = build_mlp(list(dims), activation=act, batch_norm=bn, dropout=drop, final_nonlinearity=final)
  assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
  assert [type(m).__name__.replace('Linear', 'Linear') for m in mine] is not None
  mine.load_state_dict(ref.state_dict())
  x = torch.randn(9, dims[0])
  for training in (True, False):
    if drop > 0 and training:
      continue                                         # dropout masks are RNG-dependent
    ref.train(training); mine.train(training)
    with cpu_ops():
      got = mine(x)
    assert rel_err(got, ref(x)) < TOL
  for k, v in ref.state_dict().items():
    assert rel_err(mine.state_dict()[k].float(), v.float()) < TOL, k


@pytest.mark.parametrize('arch,norm,pad,pool,size', [
    ('I5,C3-8,U2,C3-6', 'batch', 'same', 'avg', 8),
    ('C3-8,P2,C3-8-2,FC-32-10,FC-10-3', 'none', 'same', 'avg', 8),
    ('C4-8-2,C4-8-2', 'batch', 'valid', 'avg', 16),
    ('C1-4,C3-4', 'none', 'same', 'avg', 6),
    ('C3-8,U2,C3-6,C3-4-2', 'instance', 'same', 'avg', 8),
    ('C3-8,P2,C3-8,P3', 'batch', 'same', 'max', 13),          # max pooling, ragged 13 -> 6 -> 2
    ('C3-4,P4,C1-4', 'none', 'same', 'avg', 9),               # average pooling by 4, ragged
    ('R,C3-8,R,P2,R', 'batch', 'same', 'max', 8),             # residual blocks (first one un-normalised)
    ('I4,C3-4,R', 'instance', 'same', 'avg', 6)])
def test_build_cnn_mirror_vs_live_reference(arch, norm, pad, pool, size):
  """sg2im/layers.py:129-213: input-channel spec, upsample, average pooling,
  fully-connected tail, stride / padding variants."""
  import_reference()
  from sg2im.layers import build_cnn as ref_build
  from sg2im_b200.layers import build_cnn
  from cpu_shim import cpu_ops
  torch.manual_seed(13)
  with _quiet():
    ref, c_ref = ref_build(arch, normalization=norm, activation='leakyrelu-0.2', padding=pad,
                           pooling=pool)
    mine, c_mine = build_cnn(arch, normalization=norm, activation='leakyrelu-0.2', padding=pad,
                             pooling=pool)
  assert c_ref == c_mine
  assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
  mine.load_state_dict(ref.state_dict())
  cin = int(arch[1]) if arch.startswith('I') else 3
  x = torch.randn(2, cin, size, size)
  for training in (True, False):
    ref.train(training); mine.train(training)
    with cpu_ops():
      got = mine(x)
    want = ref(x)
    assert got.shape == want.shape and rel_err(got, want) < TOL
  # running statistics incl. the residual block's double evaluation (layers.py:115-116)
  for k, v in ref.state_dict().items():
    assert rel_err(mine.state_dict()[k].float(), v.float()) < TOL, k


def test_generator_with_batchnorm_mlps_vs_live_reference():
  """mlp_normalization='batch' (BatchNorm1d inside the scene-graph MLPs, model.py:57-66)."""
  import_reference()
  from sg2im.model import Sg2ImModel as RefModel
  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.synth import make_vocab
  from cpu_shim import cpu_ops
  kw = dict(embedding_dim=8, gconv_dim=8, gconv_hidden_dim=16, gconv_num_layers=2,
            refinement_dims=(16, 8), mask_size=8, layout_noise_dim=0, image_size=(16, 16),
            mlp_normalization='batch')
  vocab = make_vocab(7, 4)
  torch.manual_seed(21)
  with _quiet():
    ref = RefModel(vocab=vocab, **kw)
    mine = Sg2ImModel(vocab=vocab, **kw)
  with torch.no_grad():
    ref.box_net[3].bias.copy_(torch.tensor([0.1, 0.15, 0.6, 0.7]))
  assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
  mine.load_state_dict(ref.state_dict())
  imgs, objs, boxes, triples, o2i, _ = _batch(16, 16, seed=9)
  ref.train(); mine.train()
  out_ref = ref(objs, triples, o2i, boxes_gt=boxes)
  with cpu_ops():
    out_mine = mine(objs, triples, o2i, boxes_gt=boxes, num_imgs=3)
  for r, m in zip(out_ref, out_mine):
    assert rel_err(m, r) < TOL
  for k, v in ref.state_dict().items():
    assert rel_err(mine.state_dict()[k].float(), v.float()) < TOL, k


def test_refinement_network_standalone_forward_backward_vs_live_reference():
  """sg2im/crn.py:68-111 called on its own (layout tensor in, image out): exercises the
  stage-buffer construction from an existing layout and its hand-written backward
  (crn._StackFromLayout: average-pool cascade folded back into the layout gradient)."""
  import_reference()
  from sg2im.crn import RefinementNetwork as RefNet
  from sg2im_b200.crn import RefinementNetwork
  from cpu_shim import cpu_ops
  dims = (6, 16, 8, 8)
  torch.manual_seed(31)
  ref = RefNet(dims=dims, normalization='batch', activation='leakyrelu-0.2')
  mine = RefinementNetwork(dims=dims, normalization='batch', activation='leakyrelu-0.2')
  assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
  mine.load_state_dict(ref.state_dict())
  layout = torch.randn(2, 6, 16, 16)
  gy = torch.randn(2, 3, 16, 16)
  lr = layout.clone().requires_grad_(True)
  out_ref = ref(lr)
  out_ref.backward(gy)
  lm = layout.clone().requires_grad_(True)
  with cpu_ops():
    out = mine(lm)
    out.backward(gy)
  assert rel_err(out, out_ref) < TOL
  assert rel_err(lm.grad, lr.grad) < 1e-4
  gr = dict(ref.named_parameters())
  for k, p in mine.named_parameters():
    if k.endswith('net.0.bias') or k.endswith('net.3.bias'):
      continue                                           # conv bias in front of BatchNorm: rounding noise only
    assert rel_err(p.grad, gr[k].grad) < 1e-3, k


@pytest.mark.parametrize('pooling', ['sum', 'avg'])
def test_layout_functions_mirror_vs_live_reference(pooling):
  """sg2im/layout.py:30-91,131-162: boxes_to_layout / masks_to_layout as free functions, both
  pooling modes ('avg' divides by the per-image object count, clamped at 1 — image 1 below has no
  objects — and prints the counts like the reference), forward and gradient w.r.t. the vectors."""
  import_reference()
  from sg2im import layout as ref
  from sg2im_b200 import layout as mine
  from cpu_shim import cpu_ops
  g = torch.Generator().manual_seed(17)
  O, D, M, H, W = 7, 6, 4, 12, 10
  vecs = torch.randn(O, D, generator=g)
  xy = torch.rand(O, 2, generator=g) * 0.5
  boxes = torch.cat([xy, xy + 0.2 + 0.3 * torch.rand(O, 2, generator=g)], 1)
  masks = torch.rand(O, M, M, generator=g)
  o2i = torch.tensor([0, 0, 0, 2, 2, 3, 3])                 # image 1 is empty
  for fn, extra in (('boxes_to_layout', ()), ('masks_to_layout', (masks,))):
    vr = vecs.clone().requires_grad_(True)
    vm = vecs.clone().requires_grad_(True)
    with _quiet():
      want = getattr(ref, fn)(vr, boxes, *extra, o2i, H, W, pooling=pooling)
      with cpu_ops():
        got = getattr(mine, fn)(vm, boxes, *extra, o2i, H, W, pooling=pooling)
    assert got.shape == want.shape == (4, D, H, W)
    assert rel_err(got, want) < TOL
    wgt = torch.randn(want.shape, generator=g)
    (want * wgt).sum().backward()
    with cpu_ops():
      (got * wgt).sum().backward()
    assert rel_err(vm.grad, vr.grad) < TOL
  with pytest.raises(ValueError):
    mine.boxes_to_layout(vecs, boxes, o2i, H, W, pooling='max')


def test_bilinear_helpers_mirror_vs_live_reference():
  """sg2im/bilinear.py: tensor_linspace (bit for bit) and the crop entry points under their three
  names (crop_bbox_batch, crop_bbox_batch_cudnn, crop_bbox), objects not grouped by image."""
  import_reference()
  from sg2im import bilinear as ref
  from sg2im_b200 import bilinear as mine
  from cpu_shim import cpu_ops
  g = torch.Generator().manual_seed(23)
  a, b = torch.randn(5, 3, generator=g), torch.randn(5, 3, generator=g)
  assert torch.equal(mine.tensor_linspace(a, b, steps=7), ref.tensor_linspace(a, b, steps=7))
  feats = torch.randn(3, 4, 12, 10, generator=g)
  xy = torch.rand(6, 2, generator=g) * 0.5
  boxes = torch.cat([xy, xy + 0.2 + 0.3 * torch.rand(6, 2, generator=g)], 1)
  idx = torch.tensor([2, 0, 1, 0, 2, 1])
  want = ref.crop_bbox_batch(feats, boxes, idx, 6, 5)
  with cpu_ops():
    assert rel_err(mine.crop_bbox_batch(feats, boxes, idx, 6, 5), want) < TOL
    assert rel_err(mine.crop_bbox_batch_cudnn(feats, boxes, idx, 6, 5), want) < TOL
    assert rel_err(mine.crop_bbox(feats, boxes[:3], 6), ref.crop_bbox(feats, boxes[:3], 6)) < TOL
