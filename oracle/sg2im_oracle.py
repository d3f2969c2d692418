"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.

A CPU restatement (torch fp32 on the host, functional style over a flat
``state_dict``) of the one hot path of google/sg2im that this repo accelerates:
the generator + discriminator training step.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.

Why torch-CPU and not C/numpy: the reference itself is 100 % Python over
PyTorch; the arithmetic of the path lives in the third-party
dependency ``torch`` (requirements.txt:21 pins torch==0.4.0; installed here:
2.11.0).  The oracle therefore restates the reference's *composition* of those
primitives (which op, in which order, on which slices) and anchors the
primitives on the installed torch CPU kernels, i.e. "the reference under the
installed torch" (grid_sample => align_corners=False,
F.upsample => nearest).

PARITY PINNING: the reference has no tests and no golden vectors.
The oracle is pinned against outputs of the *unmodified reference itself*,
imported in the build container from /root/reference by
``tests/golden/make_golden.py`` (committed) and stored as small fixtures under
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` replays them on every
CPU run, and ``tests/test_oracle_vs_reference.py`` re-imports the reference
live when /root/reference exists.

Each function cites the reference file:line it follows (paths relative to
/root/reference).
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------

def activation_slope(name):
  """sg2im/layers.py:33-46.  The reference overwrites ``name`` with
  'leakyrelu' unconditionally (line 39), so *every* activation string yields a
  LeakyReLU; the slope is parsed only from 'leakyrelu-<s>' and otherwise is
  nn.LeakyReLU's default 0.01."""
  if name.lower().startswith('leakyrelu') and '-' in name:
    return float(name.split('-')[1])
  return 0.01


def mlp(sd, prefix, x, num_linear=2):
  """sg2im/layers.py:216-232 with the defaults every caller on the path uses
  (activation='relu', batch_norm='none', final_nonlinearity=True): Linear then
  ReLU for every layer, including the last.  Linear i lives at index 2*i."""
  for i in range(num_linear):
    w = sd['%s.%d.weight' % (prefix, 2 * i)]
    b = sd['%s.%d.bias' % (prefix, 2 * i)]
    x = torch.relu(F.linear(x, w, b))
  return x


def batchnorm2d(sd, prefix, x, training, momentum=0.1, eps=1e-5):
  """nn.BatchNorm2d as built by sg2im/layers.py:22-30 / sg2im/model.py:99.
  Train mode: batch statistics (biased var for normalisation, unbiased for
  the running update), running stats updated in place in ``sd``."""
  rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
  if training and (prefix + '.num_batches_tracked') in sd:
    sd[prefix + '.num_batches_tracked'] += 1
  return F.batch_norm(x, rm, rv, sd[prefix + '.weight'], sd[prefix + '.bias'],
                      training, momentum, eps)


# --------------------------------------------------------------------------
# a4: scene-graph convolution
# --------------------------------------------------------------------------

def graph_pool(new_t_vecs, edges, num_objs, hidden, dout, pooling='avg'):
  """sg2im/graph.py:85-114.  Sum the subject slice [0:H] of every triple into
  its subject row, then the object slice [H+Dout:2H+Dout] into its object row
  (two scatter_adds in that order => per destination: subject contributions in
  ascending t, then object contributions in ascending t — the summation order
  the CUDA kernel reproduces bit-exactly), then divide by clamp(count, 1)."""
  s_idx = edges[:, 0].contiguous()
  o_idx = edges[:, 1].contiguous()
  new_s = new_t_vecs[:, :hidden]
  new_o = new_t_vecs[:, hidden + dout:2 * hidden + dout]
  pooled = torch.zeros(num_objs, hidden, dtype=new_t_vecs.dtype)
  pooled = pooled.scatter_add(0, s_idx.view(-1, 1).expand_as(new_s), new_s)
  pooled = pooled.scatter_add(0, o_idx.view(-1, 1).expand_as(new_o), new_o)
  if pooling == 'avg':
    ones = torch.ones(edges.size(0), dtype=new_t_vecs.dtype)
    counts = torch.zeros(num_objs, dtype=new_t_vecs.dtype)
    counts = counts.scatter_add(0, s_idx, ones).scatter_add(0, o_idx, ones)
    pooled = pooled / counts.clamp(min=1).view(-1, 1)
  return pooled


def graph_pool_sequential(new_t_vecs, edges, num_objs, hidden, dout, pooling='avg'):
  """Pure-Python loop statement of graph_pool (small cases only): defines the
  summation order contract without relying on scatter_add's implementation."""
  T = edges.size(0)
  pooled = torch.zeros(num_objs, hidden, dtype=new_t_vecs.dtype)
  counts = [0] * num_objs
  for t in range(T):
    s = int(edges[t, 0])
    pooled[s] = pooled[s] + new_t_vecs[t, :hidden]
    counts[s] += 1
  for t in range(T):
    o = int(edges[t, 1])
    pooled[o] = pooled[o] + new_t_vecs[t, hidden + dout:2 * hidden + dout]
    counts[o] += 1
  if pooling == 'avg':
    c = torch.tensor([max(x, 1) for x in counts], dtype=new_t_vecs.dtype)
    pooled = pooled / c.view(-1, 1)
  return pooled


def graph_triple_conv(sd, prefix, obj_vecs, pred_vecs, edges, pooling='avg'):
  """sg2im/graph.py:56-120."""
  w1a = sd[prefix + '.net1.0.weight']   # (H, 3*Din)
  w1b = sd[prefix + '.net1.2.weight']   # (2H+Dout, H)
  hidden = w1a.size(0)
  dout = w1b.size(0) - 2 * hidden
  s_idx, o_idx = edges[:, 0], edges[:, 1]
  cur_t = torch.cat([obj_vecs[s_idx], pred_vecs, obj_vecs[o_idx]], dim=1)
  new_t = mlp(sd, prefix + '.net1', cur_t)
  new_p = new_t[:, hidden:hidden + dout]
  pooled = graph_pool(new_t, edges, obj_vecs.size(0), hidden, dout, pooling)
  new_obj = mlp(sd, prefix + '.net2', pooled)
  return new_obj, new_p


# --------------------------------------------------------------------------
# a6: layout
# --------------------------------------------------------------------------

def boxes_to_grid(boxes, H, W):
  """sg2im/layout.py:94-128: sampling grid in [-1,1] that maps the output
  canvas onto each box's local [0,1]^2 frame."""
  O = boxes.size(0)
  x0, y0, x1, y1 = [boxes[:, i].view(O, 1, 1) for i in range(4)]
  X = torch.linspace(0, 1, steps=W).view(1, 1, W).to(boxes)
  Y = torch.linspace(0, 1, steps=H).view(1, H, 1).to(boxes)
  X = ((X - x0) / (x1 - x0)).expand(O, H, W)
  Y = ((Y - y0) / (y1 - y0)).expand(O, H, W)
  return torch.stack([X, Y], dim=3).mul(2).sub(1)


def pool_samples(samples, obj_to_img, num_imgs):
  """sg2im/layout.py:131-162 with pooling='sum' (the only mode the model
  uses, model.py:157-162); N is passed in instead of the .item() host sync."""
  O, D, H, W = samples.size()
  out = torch.zeros(num_imgs, D, H, W, dtype=samples.dtype)
  idx = obj_to_img.view(O, 1, 1, 1).expand(O, D, H, W)
  return out.scatter_add(0, idx, samples)


def masks_to_layout(vecs, boxes, masks, obj_to_img, H, W, num_imgs, align_corners=False):
  """sg2im/layout.py:66-91 (grid_sample defaults under torch>=1.3:
  bilinear, zeros padding, align_corners=False — SURVEY.md §0.5;
  align_corners=True is the torch-0.4 convention the published checkpoints
  were trained under)."""
  O, D = vecs.size()
  M = masks.size(1)
  grid = boxes_to_grid(boxes, H, W)
  img_in = vecs.view(O, D, 1, 1) * masks.float().view(O, 1, M, M)
  sampled = F.grid_sample(img_in, grid, mode='bilinear', padding_mode='zeros',
                          align_corners=align_corners)
  return pool_samples(sampled, obj_to_img, num_imgs)


def boxes_to_layout(vecs, boxes, obj_to_img, H, W, num_imgs, align_corners=False):
  """sg2im/layout.py:30-63: same warp with a constant 8x8 "mask" of ones."""
  O, D = vecs.size()
  grid = boxes_to_grid(boxes, H, W)
  img_in = vecs.view(O, D, 1, 1).expand(O, D, 8, 8)
  sampled = F.grid_sample(img_in, grid, mode='bilinear', padding_mode='zeros',
                          align_corners=align_corners)
  return pool_samples(sampled, obj_to_img, num_imgs)


# --------------------------------------------------------------------------
# a5: mask head, a8: cascaded refinement network
# --------------------------------------------------------------------------

def mask_net(sd, prefix, obj_vecs, training):
  """sg2im/model.py:94-106 (+ the call at :145-147): per object
  [up x2 nearest -> BN -> conv3x3 -> ReLU] until mask_size, then conv1x1 -> 1
  channel, squeeze, sigmoid."""
  x = obj_vecs.view(obj_vecs.size(0), -1, 1, 1)
  i = 0
  while (prefix + '.%d.running_mean' % (i + 1)) in sd:
    x = F.interpolate(x, scale_factor=2, mode='nearest')
    x = batchnorm2d(sd, prefix + '.%d' % (i + 1), x, training)
    x = F.conv2d(x, sd[prefix + '.%d.weight' % (i + 2)],
                 sd[prefix + '.%d.bias' % (i + 2)], padding=1)
    x = torch.relu(x)
    i += 4
  x = F.conv2d(x, sd[prefix + '.%d.weight' % i], sd[prefix + '.%d.bias' % i])
  return x.squeeze(1).sigmoid()


def instancenorm2d(x, eps=1e-5):
  """nn.InstanceNorm2d with torch's defaults, what get_normalization_2d('instance') builds
  (sg2im/layers.py:24): per (image, channel) mean and BIASED variance over H*W, no affine
  parameters, no running statistics (so the same in train and eval mode)."""
  mean = x.mean(dim=(2, 3), keepdim=True)
  var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
  return (x - mean) / torch.sqrt(var + eps)


def refinement_module(sd, prefix, layout, feats, slope, normalization, training):
  """sg2im/crn.py:54-65 (+ layer list :40-52)."""
  HH, H = layout.size(2), feats.size(2)
  if HH > H:
    factor = HH // H
    layout = F.avg_pool2d(layout, kernel_size=factor, stride=factor)
  x = torch.cat([layout, feats], dim=1)            # layout channels FIRST
  if normalization in ('batch', 'instance'):         # the norm module occupies a Sequential slot
    conv_idx, bn_idx = (0, 3), (1, 4)
  elif normalization == 'none':
    conv_idx, bn_idx = (0, 2), (None, None)
  else:
    raise ValueError('oracle covers normalization in {batch, instance, none}')
  for ci, bi in zip(conv_idx, bn_idx):
    x = F.conv2d(x, sd['%s.net.%d.weight' % (prefix, ci)],
                 sd['%s.net.%d.bias' % (prefix, ci)], padding=1)
    if normalization == 'instance':
      x = instancenorm2d(x)
    elif bi is not None:
      x = batchnorm2d(sd, '%s.net.%d' % (prefix, bi), x, training)
    x = F.leaky_relu(x, slope)
  return x


def refinement_network(sd, prefix, layout, slope, normalization, training):
  """sg2im/crn.py:88-111."""
  N, _, H, W = layout.size()
  n_mod = 0
  while ('%s.refinement_modules.%d.net.0.weight' % (prefix, n_mod)) in sd:
    n_mod += 1
  feats = torch.zeros(N, 1, H >> n_mod, W >> n_mod, dtype=layout.dtype)
  for i in range(n_mod):
    feats = F.interpolate(feats, scale_factor=2, mode='nearest')
    feats = refinement_module(sd, '%s.refinement_modules.%d' % (prefix, i),
                              layout, feats, slope, normalization, training)
  x = F.conv2d(feats, sd[prefix + '.output_conv.0.weight'],
               sd[prefix + '.output_conv.0.bias'], padding=1)
  x = F.leaky_relu(x, slope)
  return F.conv2d(x, sd[prefix + '.output_conv.2.weight'],
                  sd[prefix + '.output_conv.2.bias'])


# --------------------------------------------------------------------------
# a1: generator forward
# --------------------------------------------------------------------------

def generator_forward(sd, image_size, objs, triples, obj_to_img=None,
                      boxes_gt=None, masks_gt=None, noise=None, training=True,
                      activation='leakyrelu-0.2', normalization='batch',
                      gconv_pooling='avg', num_imgs=None):
  """sg2im/model.py:108-171.  ``noise`` (N, noise_dim, H, W) replaces the
  torch.randn draw at :164-169 so CPU and GPU runs see the same values;
  pass None for layout_noise_dim == 0."""
  O = objs.size(0)
  s, p, o = triples[:, 0], triples[:, 1], triples[:, 2]
  edges = torch.stack([s, o], dim=1)
  if obj_to_img is None:
    obj_to_img = torch.zeros(O, dtype=objs.dtype)
  if num_imgs is None:
    num_imgs = int(obj_to_img.max()) + 1

  obj_vecs = sd['obj_embeddings.weight'][objs]
  obj_vecs_orig = obj_vecs
  pred_vecs = sd['pred_embeddings.weight'][p]

  if 'gconv.weight' in sd:                         # gconv_num_layers == 0
    obj_vecs = F.linear(obj_vecs, sd['gconv.weight'], sd['gconv.bias'])
  else:
    obj_vecs, pred_vecs = graph_triple_conv(sd, 'gconv', obj_vecs, pred_vecs,
                                            edges, gconv_pooling)
  layer = 0
  while ('gconv_net.gconvs.%d.net1.0.weight' % layer) in sd:
    obj_vecs, pred_vecs = graph_triple_conv(
        sd, 'gconv_net.gconvs.%d' % layer, obj_vecs, pred_vecs, edges,
        gconv_pooling)
    layer += 1

  boxes_pred = mlp(sd, 'box_net', obj_vecs)

  masks_pred = None
  if 'mask_net.1.running_mean' in sd:
    masks_pred = mask_net(sd, 'mask_net', obj_vecs, training)

  rel_in = torch.cat([boxes_pred[s], boxes_pred[o],
                      obj_vecs_orig[s], obj_vecs_orig[o]], dim=1)
  rel_scores = mlp(sd, 'rel_aux_net', rel_in)

  H, W = image_size
  layout_boxes = boxes_pred if boxes_gt is None else boxes_gt
  if masks_pred is None:
    layout = boxes_to_layout(obj_vecs, layout_boxes, obj_to_img, H, W, num_imgs)
  else:
    layout_masks = masks_pred if masks_gt is None else masks_gt
    layout = masks_to_layout(obj_vecs, layout_boxes, layout_masks, obj_to_img,
                             H, W, num_imgs)
  if noise is not None:
    layout = torch.cat([layout, noise], dim=1)
  img = refinement_network(sd, 'refinement_net', layout,
                           activation_slope(activation), normalization, training)
  return img, boxes_pred, masks_pred, rel_scores


# --------------------------------------------------------------------------
# a10-a13: discriminators
# --------------------------------------------------------------------------

def parse_conv_arch(arch):
  """The 'CK-X[-S]' subset of sg2im/layers.py:129-213 (what the default
  --d_obj_arch/--d_img_arch use, scripts/train.py:122-130).  Returns
  [(K, Cout, stride), ...]."""
  if isinstance(arch, str):
    arch = arch.split(',')
  out = []
  for s in arch:
    if s[0] == 'I':
      continue
    if s[0] != 'C':
      raise ValueError('oracle covers conv-only architectures, got "%s"' % s)
    vals = [int(v) for v in s[1:].split('-')]
    out.append((vals[0], vals[1], vals[2] if len(vals) == 3 else 1))
  return out


def disc_cnn(sd, prefix, x, arch, normalization, activation, padding, training):
  """build_cnn's Sequential (layers.py:164-181): conv; then for every later
  conv: norm, activation, conv.  Module indices advance by 1 per present layer."""
  slope = activation_slope(activation)
  idx = 0
  for li, (K, _, stride) in enumerate(parse_conv_arch(arch)):
    if li > 0:
      if normalization == 'batch':
        x = batchnorm2d(sd, '%s.%d' % (prefix, idx), x, training)
        idx += 1
      elif normalization == 'instance':
        x = instancenorm2d(x)
        idx += 1
      elif normalization != 'none':
        raise ValueError('oracle covers normalization in {batch, instance, none}')
      x = F.leaky_relu(x, slope)
      idx += 1
    P = 0 if padding == 'valid' else (K - 1) // 2
    x = F.conv2d(x, sd['%s.%d.weight' % (prefix, idx)],
                 sd['%s.%d.bias' % (prefix, idx)], stride=stride, padding=P)
    idx += 1
  return x


def patch_discriminator(sd, x, arch, normalization='batch',
                        activation='leakyrelu-0.2', padding='valid',
                        training=True):
  """sg2im/discriminators.py:42-45: returns cnn(x); ``classifier`` is never
  applied."""
  return disc_cnn(sd, 'cnn', x, arch, normalization, activation, padding, training)


def crop_bbox_batch(feats, bbox, bbox_to_feats, HH, WW=None, align_corners=False):
  """sg2im/bilinear.py:28-43 -> :69-100 -> :103-132 -> :249-278.  For box b:
  X = (1-a)*(2*x0-1) + a*(2*x1-1), a = linspace(0,1,WW) (tensor_linspace's
  start_w*start + end_w*end form), Y likewise; bilinear grid_sample of
  feats[bbox_to_feats[b]] (zeros padding, align_corners=False).  The
  reference's group-by-image / inverse-permute dance is an implementation
  detail: the result is crops[b] for every b in input order."""
  if WW is None:
    WW = HH
  B = bbox.size(0)
  bb = 2 * bbox - 1
  x0, y0, x1, y1 = bb[:, 0], bb[:, 1], bb[:, 2], bb[:, 3]
  wx1 = torch.linspace(0, 1, steps=WW).view(1, WW)
  wx0 = torch.linspace(1, 0, steps=WW).view(1, WW)
  wy1 = torch.linspace(0, 1, steps=HH).view(1, HH)
  wy0 = torch.linspace(1, 0, steps=HH).view(1, HH)
  X = (wx0 * x0.view(B, 1) + wx1 * x1.view(B, 1)).view(B, 1, WW).expand(B, HH, WW)
  Y = (wy0 * y0.view(B, 1) + wy1 * y1.view(B, 1)).view(B, HH, 1).expand(B, HH, WW)
  grid = torch.stack([X, Y], dim=3)
  return F.grid_sample(feats[bbox_to_feats], grid, mode='bilinear',
                       padding_mode='zeros', align_corners=align_corners)


def ac_crop_discriminator(sd, imgs, objs, boxes, obj_to_img, arch, object_size,
                          normalization='batch', activation='leakyrelu-0.2',
                          padding='valid', training=True):
  """sg2im/discriminators.py:87-90 + :68-75: crop -> cnn -> GAP -> Linear(D,1024)
  -> {real_classifier, obj_classifier}; ac_loss = cross_entropy(obj_scores, objs)."""
  crops = crop_bbox_batch(imgs, boxes, obj_to_img, object_size)
  x = disc_cnn(sd, 'discriminator.cnn.0', crops, arch, normalization,
               activation, padding, training)
  x = x.view(x.size(0), x.size(1), -1).mean(dim=2)            # layers.py:83-86
  vecs = F.linear(x, sd['discriminator.cnn.2.weight'], sd['discriminator.cnn.2.bias'])
  real_scores = F.linear(vecs, sd['discriminator.real_classifier.weight'],
                         sd['discriminator.real_classifier.bias'])
  obj_scores = F.linear(vecs, sd['discriminator.obj_classifier.weight'],
                        sd['discriminator.obj_classifier.bias'])
  return real_scores, F.cross_entropy(obj_scores, objs)


# --------------------------------------------------------------------------
# a14: losses, a15: the training step
# --------------------------------------------------------------------------

def bce_loss(x, target):
  """sg2im/losses.py:39-57."""
  return (x.clamp(min=0) - x * target + (1 + (-x.abs()).exp()).log()).mean()


def gan_g_loss(scores_fake):
  """sg2im/losses.py:72-83."""
  s = scores_fake.reshape(-1)
  return bce_loss(s, torch.ones_like(s))


def gan_d_loss(scores_real, scores_fake):
  """sg2im/losses.py:86-103."""
  r, f = scores_real.reshape(-1), scores_fake.reshape(-1)
  return bce_loss(r, torch.ones_like(r)) + bce_loss(f, torch.zeros_like(f))


DEFAULT_ARGS = dict(                      # scripts/train.py:94-131 defaults
    l1_pixel_loss_weight=1.0, bbox_pred_loss_weight=10.0,
    predicate_pred_loss_weight=0.0, mask_loss_weight=0.0,
    discriminator_loss_weight=0.01, d_obj_weight=1.0, d_img_weight=1.0,
    ac_loss_weight=0.1, d_arch='C4-64-2,C4-128-2,C4-256-2',
    d_normalization='batch', d_activation='leakyrelu-0.2', d_padding='valid',
    crop_size=32, activation='leakyrelu-0.2', normalization='batch',
    learning_rate=1e-4)


def generator_losses(args, img, img_pred, bbox, bbox_pred, masks, masks_pred,
                     predicates, predicate_scores):
  """scripts/train.py:387-412 (boxes_gt is always given => pixel loss kept)."""
  losses = {}
  total = torch.zeros(1, dtype=img.dtype)
  l1 = F.l1_loss(img_pred, img) * args['l1_pixel_loss_weight']
  losses['L1_pixel_loss'] = l1
  total = total + l1
  lb = F.mse_loss(bbox_pred, bbox) * args['bbox_pred_loss_weight']
  losses['bbox_pred'] = lb
  total = total + lb
  if args['predicate_pred_loss_weight'] > 0:
    lp = F.cross_entropy(predicate_scores, predicates) * args['predicate_pred_loss_weight']
    losses['predicate_pred'] = lp
    total = total + lp
  if args['mask_loss_weight'] > 0 and masks is not None and masks_pred is not None:
    lm = F.binary_cross_entropy(masks_pred, masks.float()) * args['mask_loss_weight']
    losses['mask_loss'] = lm
    total = total + lm
  return total, losses


class OracleTrainer(object):
  """One G + D_obj + D_img iteration, scripts/train.py:508-592, over three flat
  state dicts.  Parameters (floating tensors that the reference registers as
  nn.Parameter) become autograd leaves updated by torch.optim.Adam(lr=1e-4),
  exactly as train.py:426,436,443; buffers (running stats) stay plain tensors."""

  BUFFER_SUFFIXES = ('running_mean', 'running_var', 'num_batches_tracked')

  def __init__(self, g_sd, dobj_sd, dimg_sd, image_size, args=None):
    self.args = dict(DEFAULT_ARGS)
    if args:
      self.args.update(args)
    self.image_size = tuple(image_size)
    self.sd = {}
    self.opt = {}
    for name, src in (('g', g_sd), ('d_obj', dobj_sd), ('d_img', dimg_sd)):
      sd = {}
      params = []
      for k, v in src.items():
        v = v.detach().clone()
        if not k.endswith(self.BUFFER_SUFFIXES):
          v.requires_grad_(True)
          params.append(v)
        sd[k] = v
      self.sd[name] = sd
      self.opt[name] = torch.optim.Adam(params, lr=self.args['learning_rate'])

  def step(self, batch, noise):
    a = self.args
    if len(batch) == 6:
      imgs, objs, boxes, triples, obj_to_img, _ = batch
      masks = None
    else:
      imgs, objs, boxes, masks, triples, obj_to_img, _ = batch
    N = imgs.size(0)
    g, dobj, dimg = self.sd['g'], self.sd['d_obj'], self.sd['d_img']
    d_kw = dict(arch=a['d_arch'], normalization=a['d_normalization'],
                activation=a['d_activation'], padding=a['d_padding'])

    # ---- generator step: train.py:524-560
    imgs_pred, boxes_pred, masks_pred, rel_scores = generator_forward(
        g, self.image_size, objs, triples, obj_to_img, boxes_gt=boxes,
        masks_gt=masks, noise=noise, training=True, activation=a['activation'],
        normalization=a['normalization'], num_imgs=N)
    total, losses = generator_losses(a, imgs, imgs_pred, boxes, boxes_pred,
                                     masks, masks_pred, triples[:, 1], rel_scores)
    scores_fake, ac_loss = ac_crop_discriminator(
        dobj, imgs_pred, objs, boxes, obj_to_img, object_size=a['crop_size'], **d_kw)
    losses['ac_loss'] = ac_loss * a['ac_loss_weight']
    total = total + losses['ac_loss']
    losses['g_gan_obj_loss'] = gan_g_loss(scores_fake) * (
        a['discriminator_loss_weight'] * a['d_obj_weight'])
    total = total + losses['g_gan_obj_loss']
    scores_fake = patch_discriminator(dimg, imgs_pred, **d_kw)
    losses['g_gan_img_loss'] = gan_g_loss(scores_fake) * (
        a['discriminator_loss_weight'] * a['d_img_weight'])
    total = total + losses['g_gan_img_loss']
    losses['total_loss'] = total
    out = {k: float(v) for k, v in losses.items()}
    if not math.isfinite(out['total_loss']):
      return out, imgs_pred.detach()
    self.opt['g'].zero_grad()
    total.backward()
    self.opt['g'].step()

    # ---- object discriminator step: train.py:566-579
    imgs_fake = imgs_pred.detach()
    s_fake, ac_fake = ac_crop_discriminator(
        dobj, imgs_fake, objs, boxes, obj_to_img, object_size=a['crop_size'], **d_kw)
    s_real, ac_real = ac_crop_discriminator(
        dobj, imgs, objs, boxes, obj_to_img, object_size=a['crop_size'], **d_kw)
    d_obj_gan = gan_d_loss(s_real, s_fake)
    d_obj_total = d_obj_gan + ac_real + ac_fake
    out['d_obj_gan_loss'] = float(d_obj_gan)
    out['d_ac_loss_real'] = float(ac_real)
    out['d_ac_loss_fake'] = float(ac_fake)
    self.opt['d_obj'].zero_grad()
    d_obj_total.backward()
    self.opt['d_obj'].step()

    # ---- image discriminator step: train.py:581-592
    s_fake = patch_discriminator(dimg, imgs_fake, **d_kw)
    s_real = patch_discriminator(dimg, imgs, **d_kw)
    d_img_gan = gan_d_loss(s_real, s_fake)
    out['d_img_gan_loss'] = float(d_img_gan)
    self.opt['d_img'].zero_grad()
    d_img_gan.backward()
    self.opt['d_img'].step()
    return out, imgs_fake
