"""One generator + object-discriminator + image-discriminator training
iteration — the loop body of the reference's scripts/train.py:508-592 as a
reusable object, plus the data-parallel plumbing the reference never had.

Multi-GPU: one process per GPU, the batch sharded by image, no
exchange in the forward pass (objects and triples never cross images),
BatchNorm statistics rank-local, and exactly one collective per optimiser per
iteration: an NCCL all-reduce (sum, then / world) of that network's flat fp32
gradient bucket.  The non-finite-loss skip of train.py:552-555 is made
collective (all-reduce MIN of the finite flag) so ranks cannot diverge.
"""
import collections
import os
import math

import torch
import torch.nn.functional as F

from .losses import get_gan_losses

DEFAULT_ARGS = dict(                      # scripts/train.py:94-131
    l1_pixel_loss_weight=1.0, bbox_pred_loss_weight=10.0,
    predicate_pred_loss_weight=0.0, mask_loss_weight=0.0,
    discriminator_loss_weight=0.01, d_obj_weight=1.0, d_img_weight=1.0,
    ac_loss_weight=0.1, gan_loss_type='gan', learning_rate=1e-4)


class FlatGrads(object):
  """All gradients of one network in a single fp32 bucket: every parameter's
  ``.grad`` is a view into ``flat``, so zeroing is one memset and the
  data-parallel exchange is one all-reduce with no packing copies."""

  def __init__(self, params, align=1):
    """align: every parameter's slice starts at a multiple of `align` elements
    (FlatAdam lays the parameters out the same way and needs 16-byte aligned
    weights/biases for the kernels: align=4); padding stays zero."""
    self.params = [p for p in params if p.requires_grad]
    self.offsets, off = [], 0
    for p in self.params:
      self.offsets.append(off)
      off += -(-p.numel() // align) * align
    dev = self.params[0].device
    self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
    for p, o in zip(self.params, self.offsets):
      p.grad = self._slot(p, o)

  def _slot(self, p, o):
    """The parameter's slice of the bucket with the parameter's own strides (parameters stored
    in a permuted layout, layers.to_kcc_, get gradients in the same layout)."""
    return self.flat[o:o + p.numel()].as_strided(p.shape, p.stride())

  def zero(self):
    self.flat.zero_()
    # re-attach in case something replaced .grad (e.g. set_to_none)
    for p, o in zip(self.params, self.offsets):
      if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * o:
        p.grad = self._slot(p, o)

  def all_reduce_mean(self, group=None, opt=None):
    """Gradient mean over the ranks: one SUM all-reduce of the flat bucket.  The division by the
    world size is folded into the optimiser's gradient scale when it has one (FlatAdam: no extra
    pass over the bucket), else done in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
      world = dist.get_world_size(group)
      if world > 1:
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if isinstance(opt, FlatAdam):
          opt.grad_scale = 1.0 / world
        else:
          self.flat.div_(world)


class FlatAdam(object):
  """Adam over one flat bucket per network: the parameters are
  re-pointed into a flat fp32 buffer laid out like the FlatGrads bucket, the two
  moments are flat too, and one update is one streaming kernel
  (``sg2im_adam_flat``) instead of torch.optim.Adam's multi-tensor launches.
  Same arithmetic and defaults as torch.optim.Adam (amsgrad=False); the step
  count lives on the device and ``found_inf`` (0-dim device float, nonzero =
  skip) makes it usable inside a CUDA graph.  ``state_dict`` keys and values of
  the network are untouched (parameters stay ``nn.Parameter``s of the same
  shape; only their storage moves)."""

  def __init__(self, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, shadow=False):
    """shadow: keep a round-to-nearest-TF32 copy of the parameters (written by the same kernel)
    and hang each parameter's view of it on the parameter as `_tc_shadow` — what the tensor-core
    kernels read when the weights live in the weight-gradient layout (ops.ConvKCC); call
    refresh_shadow() after loading a state_dict."""
    self.bucket = bucket
    self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
    flat = bucket.flat
    self.flat_params = torch.zeros_like(flat)
    with torch.no_grad():
      for p, o in zip(bucket.params, bucket.offsets):
        if o % 4:
          raise ValueError('FlatAdam needs FlatGrads(align=4)')
        dst = self.flat_params[o:o + p.numel()].as_strided(p.shape, p.stride())
        dst.copy_(p)
        p.data = dst
    self.exp_avg = torch.zeros_like(flat)
    self.exp_avg_sq = torch.zeros_like(flat)
    self.step_count = torch.zeros((), dtype=torch.float32, device=flat.device)
    self.found_inf = None            # set by the graph path, like torch's capturable Adam
    self.grad_scale = 1.0            # multiplies the gradients (FlatGrads.all_reduce_mean: 1 / world)
    self.shadow = None
    if shadow:
      self.shadow = torch.zeros_like(flat)
      for p, o in zip(bucket.params, bucket.offsets):
        p._tc_shadow = self.shadow[o:o + p.numel()].as_strided(p.shape, p.stride())
      self.refresh_shadow()

  def refresh_shadow(self):
    if self.shadow is not None:
      from . import ops
      ops.round_tf32(self.flat_params, self.shadow)

  def _check_attached(self):
    for p, o in zip(self.bucket.params, self.bucket.offsets):
      if p.data_ptr() != self.flat_params.data_ptr() + 4 * o:
        raise RuntimeError('sg2im_b200: a parameter was re-allocated (model.to() / .float() / '
                           'assignment to .data) after FlatAdam took it over')

  def state_dict(self):
    """torch.optim.Adam's format (per-parameter `step`, `exp_avg`, `exp_avg_sq`; one param group),
    so checkpoints written by scripts/train.py:633-641 style code resume with either optimiser."""
    state = {}
    for i, (p, o) in enumerate(zip(self.bucket.params, self.bucket.offsets)):
      sl = slice(o, o + p.numel())
      state[i] = {'step': self.step_count.detach().clone(),
                  'exp_avg': self.exp_avg[sl].as_strided(p.shape, p.stride()).clone(),
                  'exp_avg_sq': self.exp_avg_sq[sl].as_strided(p.shape, p.stride()).clone()}
    group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay,
                 amsgrad=False, params=list(range(len(self.bucket.params))))
    return {'state': state, 'param_groups': [group]}

  def load_state_dict(self, sd):
    g = sd['param_groups'][0]
    self.lr, self.betas, self.eps = g['lr'], tuple(g['betas']), g['eps']
    self.weight_decay = g.get('weight_decay', 0.0)
    with torch.no_grad():
      for i, (p, o) in enumerate(zip(self.bucket.params, self.bucket.offsets)):
        st = sd['state'].get(i, sd['state'].get(str(i)))
        if st is None:
          continue
        sl = slice(o, o + p.numel())
        self.exp_avg[sl].as_strided(p.shape, p.stride()).copy_(st['exp_avg'])
        self.exp_avg_sq[sl].as_strided(p.shape, p.stride()).copy_(st['exp_avg_sq'])
        self.step_count.fill_(float(st['step']))
    self.refresh_shadow()

  def step(self):
    from . import ops
    self._check_attached()
    ops.adam_flat(self.flat_params, self.bucket.flat, self.exp_avg, self.exp_avg_sq,
                  self.step_count, self.lr, self.betas[0], self.betas[1], self.eps,
                  self.weight_decay, self.found_inf, self.shadow, self.grad_scale)


def _all_finite(value, group=None):
  """Collective version of train.py:552: every rank takes the same branch."""
  ok = math.isfinite(value)
  import torch.distributed as dist
  if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32,
                        device='cuda' if dist.get_backend(group) == 'nccl' else 'cpu')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    ok = bool(flag.item())
  return ok


class TrainStep(object):
  """step(batch) runs one iteration.  cuda_graph=True: after `graph_warmup`
  eager iterations the whole iteration (three forward/backward passes, gradient
  all-reduces, three Adam steps — ~1500 kernel launches) is captured once per
  batch-shape signature into a CUDA graph and replayed; inputs are copied into
  static buffers, losses are read back after the replay.  Inside a graph there
  can be no host decision, so the reference's "skip the iteration on a
  non-finite loss" (train.py:552-555) becomes an on-device skip: the finite flag
  (all-reduced across ranks) is handed to the fused Adam kernels as `found_inf`,
  which leaves parameters, moments and step counts untouched."""

  def __init__(self, model, obj_discriminator, img_discriminator, args=None,
               fused_adam=None, group=None, cuda_graph=False, graph_warmup=3, weights='oihw',
               max_graphs=4, graph_min_seen=2, capture_cooldown=64):
    a = dict(DEFAULT_ARGS)
    if args is not None:
      a.update(args if isinstance(args, dict) else
               {k: getattr(args, k) for k in DEFAULT_ARGS if hasattr(args, k)})
    self.args = a
    self.model, self.d_obj, self.d_img = model, obj_discriminator, img_discriminator
    self.group = group
    self.gan_g_loss, self.gan_d_loss = get_gan_losses(a['gan_loss_type'])
    dev = next(model.parameters()).device
    if fused_adam is None:
      fused_adam = dev.type == 'cuda'
    kw = dict(lr=a['learning_rate'])
    if fused_adam and fused_adam != 'flat':
      kw['fused'] = True
    self.cuda_graph = bool(cuda_graph) and dev.type == 'cuda'
    if self.cuda_graph:
      kw['fused'] = True
      kw['capturable'] = True
    self.graph_warmup = graph_warmup
    # Captured graphs are keyed by the exact batch-shape signature.  Real VG / COCO batches change
    # their object / triple counts almost every iteration, so the cache is BOUNDED (least recently
    # used graph — and its private memory pool and static input copies — dropped beyond
    # `max_graphs`) and a signature is captured only once it has been seen `graph_min_seen` times;
    # everything else runs eagerly.  Fixed-shape loaders (the benchmark's synthetic batches, a
    # bucketing collate) get the replay path, variable-shape loaders degrade to the eager path
    # instead of recapturing ~1500 launches per step and growing memory without bound.
    # Once the cache is full, a capture that evicts a graph is allowed at most every
    # `capture_cooldown` steps: a loader cycling through more signatures than the cache holds would
    # otherwise re-capture (~130 ms) on every step (measured: 247 img/s instead of the eager 1 000+).
    self.max_graphs = max(1, int(max_graphs))
    self.graph_min_seen = max(1, int(graph_min_seen))
    self.capture_cooldown = max(0, int(capture_cooldown))
    self._steps = 0
    self._last_capture = -(1 << 30)
    self._graphs = collections.OrderedDict()   # shape signature -> captured state (LRU order)
    self._seen = collections.Counter()
    self.graph_evictions = 0
    self._eager_calls = 0
    self.launches_per_replay = 0
    self.replays = 0
    self.nets = {'g': model, 'd_obj': obj_discriminator, 'd_img': img_discriminator}
    if weights == 'kcc':
      # conv / linear weights re-stored in the weight-gradient layout: no pack / unpack passes
      from . import ops
      from .layers import to_kcc_
      for net in self.nets.values():
        if net is not None:
          to_kcc_(net)
    elif weights != 'oihw':
      raise ValueError("weights must be 'oihw' or 'kcc'")
    self.weights = weights
    self.buckets, self.opts = {}, {}
    for name, net in self.nets.items():
      if net is None:
        continue
      if fused_adam == 'flat':
        # one streaming kernel per optimiser over flat parameter / moment buckets
        self.buckets[name] = FlatGrads(net.parameters(), align=4)
        # 'tf32' reads an RN-TF32 shadow of the weights; the bf16 modes split the masters in-kernel
        from . import ops as _ops
        self.opts[name] = FlatAdam(self.buckets[name], lr=a['learning_rate'],
                                   shadow=weights == 'kcc' and _ops.CONV_MATH == 'tf32')
      else:
        # kcc: the weight-gradient kernels write float4 atomics straight into the slots
        self.buckets[name] = FlatGrads(net.parameters(), align=4 if weights == 'kcc' else 1)
        self.opts[name] = torch.optim.Adam(self.buckets[name].params, **kw)
    # bf16 arithmetic with in-place weights: one launch per network and step splits every weight
    # into the bf16 hi / mid operand copies the tensor-core kernels read (ops.SplitShadows)
    from . import ops as _ops
    self.split_shadows = []
    if weights == 'kcc' and _ops.CONV_MATH in ('bf16x3', 'bf16'):
      for net in self.nets.values():
        if net is not None:
          sh = _ops.SplitShadows(list(net.parameters()))
          if sh.table is not None:
            self.split_shadows.append(sh)
    # small zero-initialised scratch (BatchNorm statistics, column sums, tiny gradient accumulators):
    # one arena cleared by one memset per step instead of ~85 fill kernels
    self.zero_arena = _ops.ZeroArena(8 << 20, dev) if weights == 'kcc' else None
    # offset of the cascaded-refinement network's parameters in the generator's flat bucket (they
    # are its tail: `refinement_net` is Sg2ImModel's last child) for the overlapped all-reduce
    self._crn_offset = None
    crn = getattr(model, 'refinement_net', None)
    if crn is not None:
      ids = {id(p) for p in crn.parameters()}
      b = self.buckets['g']
      inside = [id(p) in ids for p in b.params]
      if any(inside):
        first = inside.index(True)
        if all(inside[first:]) and b.offsets[first] % 4 == 0:
          self._crn_offset = b.offsets[first]
    self.skipped = 0
    self._side_stream = None
    self._side_stream2 = None
    self.sync_replicas()

  def _g_backward_and_reduce(self, total):
    """Generator backward + gradient mean.  With more than one rank the all-reduce of the cascaded-
    refinement network's slice of the flat bucket (the last ~95 % of it: `refinement_net` is the
    generator's last child) is issued asynchronously as soon as its backward has finished
    (ops.GRAD_READY_HOOK, fired by the layout's backward) and overlaps the layout / mask-head /
    graph-convolution backward; the remaining head of the bucket is reduced afterwards."""
    import torch.distributed as dist
    from . import ops
    bucket, opt = self.buckets['g'], self.opts['g']
    multi = (dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1)
    # SG2IM_OVERLAP_ALLREDUCE=0: one all-reduce of the whole bucket after the backward
    split = self._crn_offset if (multi and os.environ.get('SG2IM_OVERLAP_ALLREDUCE', '1') != '0') else None
    if not split:
      total.backward()
      bucket.all_reduce_mean(self.group, opt)
      return
    work = []

    def crn_done():
      if not work:
        work.append(dist.all_reduce(bucket.flat[split:], op=dist.ReduceOp.SUM, group=self.group,
                                    async_op=True))
    prev, ops.GRAD_READY_HOOK = ops.GRAD_READY_HOOK, crn_done
    try:
      total.backward()
    finally:
      ops.GRAD_READY_HOOK = prev
    if work:
      dist.all_reduce(bucket.flat[:split], op=dist.ReduceOp.SUM, group=self.group)
      work[0].wait()
    else:                                            # no layout in the graph (never for Sg2ImModel)
      dist.all_reduce(bucket.flat, op=dist.ReduceOp.SUM, group=self.group)
    world = dist.get_world_size(self.group)
    if isinstance(opt, FlatAdam):
      opt.grad_scale = 1.0 / world
    else:
      bucket.flat.div_(world)

  def sync_replicas(self):
    """Data parallel: make every rank start from rank 0's parameters AND buffers (what
    torch DDP does at construction and, for buffers, every forward with broadcast_buffers=True).
    Gradients are averaged each step, so identical replicas stay identical; BatchNorm running
    statistics are rank-LOCAL during training like the reference run at the per-GPU batch (no
    SyncBN) — call this again before saving a checkpoint or evaluating to publish rank 0's.
    No-op without an initialised process group or with one rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) < 2:
      return
    src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
    with torch.no_grad():
      for name, net in self.nets.items():
        if net is None:
          continue
        opt = self.opts.get(name)
        if isinstance(opt, FlatAdam):
          dist.broadcast(opt.flat_params, src, group=self.group)     # every parameter lives in it
          opt.refresh_shadow()
        else:
          for p in net.parameters():
            dist.broadcast(p.data, src, group=self.group)
        for b in net.buffers():
          dist.broadcast(b, src, group=self.group)

  # -- loss assembly, scripts/train.py:387-412 + :539-550
  def generator_losses(self, imgs, imgs_pred, boxes, boxes_pred, masks, masks_pred,
                       predicates, predicate_scores):
    a = self.args
    losses = {}
    total = torch.zeros(1, dtype=imgs.dtype, device=imgs.device)

    def add(name, val, weight):
      nonlocal total
      val = val * weight
      losses[name] = val
      total = total + val

    add('L1_pixel_loss', F.l1_loss(imgs_pred, imgs), a['l1_pixel_loss_weight'])
    add('bbox_pred', F.mse_loss(boxes_pred, boxes), a['bbox_pred_loss_weight'])
    if a['predicate_pred_loss_weight'] > 0:
      add('predicate_pred', F.cross_entropy(predicate_scores, predicates),
          a['predicate_pred_loss_weight'])
    if a['mask_loss_weight'] > 0 and masks is not None and masks_pred is not None:
      add('mask_loss', F.binary_cross_entropy(masks_pred, masks.float()),
          a['mask_loss_weight'])
    return total, losses

  @staticmethod
  def _freeze(net, frozen):
    """The reference lets the generator's backward also produce (and then
    discard) weight gradients of both discriminators; turning requires_grad off
    for the G step skips those wgrads without changing any result."""
    for p in net.parameters():
      p.requires_grad_(not frozen)

  def step(self, batch, noise=None):
    """batch: the collate tuple (6 entries for VG, 7 with masks for COCO), on
    the device (graph mode also accepts pinned host tensors: they are copied
    into the static input buffers).  Returns (losses dict of python floats,
    imgs_pred detached)."""
    from . import ops
    prev = ops.DIRECT_WGRAD, ops.USE_SPLIT_SHADOWS, ops.ZERO_ARENA
    # kcc: the weight-gradient kernels accumulate in place in the flat gradient buckets
    ops.DIRECT_WGRAD = self.weights == 'kcc'
    ops.USE_SPLIT_SHADOWS = bool(self.split_shadows)
    ops.ZERO_ARENA = self.zero_arena
    try:
      if self.cuda_graph:
        return self._step_graphed(batch, noise)
      return self._step_eager(batch, noise)
    finally:
      ops.DIRECT_WGRAD, ops.USE_SPLIT_SHADOWS, ops.ZERO_ARENA = prev

  # ------------------------------------------------------------------ graph mode
  def _step_graphed(self, batch, noise):
    sig = tuple((tuple(t.shape), t.dtype) for t in batch) + (None if noise is None else tuple(noise.shape),)
    st = self._graphs.get(sig)
    dev = next(self.model.parameters()).device
    self._steps += 1
    if st is None:
      self._seen[sig] += 1
      thrash = (len(self._graphs) >= self.max_graphs
                and self._steps - self._last_capture < self.capture_cooldown)
      if self._eager_calls < self.graph_warmup or self._seen[sig] < self.graph_min_seen or thrash:
        self._eager_calls += 1
        return self._step_eager([t.to(dev, non_blocking=True) for t in batch],
                                None if noise is None else noise.to(dev, non_blocking=True))
      while len(self._graphs) >= self.max_graphs:      # drop the least recently used graph first
        _, old = self._graphs.popitem(last=False)
        old.clear()
        self.graph_evictions += 1
      st = self._capture([t.to(dev) for t in batch], noise)
      self._graphs[sig] = st
      self._last_capture = self._steps
      if len(self._seen) > 4096:
        self._seen.clear()
    else:
      self._graphs.move_to_end(sig)
    for dst, src in zip(st['batch'], batch):
      dst.copy_(src, non_blocking=True)
    if noise is not None:
      st['noise'].copy_(noise, non_blocking=True)
    st['graph'].replay()
    self.replays += 1
    vals = st['loss_vec'].tolist()                  # the one D2H sync per iteration
    out = dict(zip(st['keys'], vals))
    if not math.isfinite(out['total_loss']):
      print('WARNING: Got loss = NaN, not backpropping')
      self.skipped += 1
    # NOTE: the returned images are the graph's static output buffer — valid until the next step()
    # with the same batch signature overwrites it (clone it to keep it; the training loop of
    # scripts/train.py only logs / detaches it within the iteration)
    return out, st['imgs_fake']

  def _capture(self, batch, noise):
    dev = batch[0].device
    static_batch = [t.clone() for t in batch]
    static_noise = None if noise is None else noise.to(dev).clone()
    found_inf = torch.zeros((), dtype=torch.float32, device=dev)
    for opt in self.opts.values():
      opt.found_inf = found_inf
    from . import _lib
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    l0 = _lib.launches
    with torch.cuda.graph(graph):
      losses, imgs_fake = self._body(static_batch, static_noise, found_inf)
      keys = list(losses.keys())
      loss_vec = torch.stack([losses[k].reshape(()).detach() for k in keys])
    torch.cuda.synchronize()
    self.launches_per_replay = _lib.launches - l0      # library launches inside one replay
    return dict(graph=graph, batch=static_batch, noise=static_noise, keys=keys,
                loss_vec=loss_vec, imgs_fake=imgs_fake)

  def _body(self, batch, noise, found_inf):
    """The whole iteration with no host synchronisation (graph-capturable)."""
    a = self.args
    masks = None
    if len(batch) == 6:
      imgs, objs, boxes, triples, obj_to_img, _ = batch
    else:
      imgs, objs, boxes, masks, triples, obj_to_img, _ = batch
    N = imgs.size(0)
    for sh in self.split_shadows:          # this step's weights -> bf16 hi / mid operand copies
      sh.refresh()
    if self.zero_arena is not None:
      self.zero_arena.reset()
    imgs_pred, boxes_pred, masks_pred, predicate_scores = self.model(
        objs, triples, obj_to_img, boxes_gt=boxes, masks_gt=masks, num_imgs=N, noise=noise)
    total, losses = self.generator_losses(imgs, imgs_pred, boxes, boxes_pred, masks,
                                          masks_pred, triples[:, 1], predicate_scores)
    # the two discriminators' scores of the generated images (train.py:539-548): the image
    # discriminator runs on a second stream beside the object discriminator, forward and — autograd
    # replays a node on its forward stream — backward (2499 vs 2472 img/s, profiles/r02_call22*;
    # SG2IM_GLOSS_STREAMS=1 serialises them)
    fork = (imgs_pred.is_cuda and self.d_img is not None and self.d_obj is not None
            and os.environ.get('SG2IM_GLOSS_STREAMS', '2') != '1')
    g_img = None
    if fork:
      main = torch.cuda.current_stream()
      if self._side_stream2 is None:
        self._side_stream2 = torch.cuda.Stream()     # not the discriminator iteration's stream: this
      self._side_stream2.wait_stream(main)           # branch's backward is on the generator's chain
      self._freeze(self.d_img, True)
      with torch.cuda.stream(self._side_stream2):
        g_img = self.gan_g_loss(self.d_img(imgs_pred)) * (a['discriminator_loss_weight'] * a['d_img_weight'])
    if self.d_obj is not None:
      self._freeze(self.d_obj, True)
      scores_fake, ac_loss = self.d_obj(imgs_pred, objs, boxes, obj_to_img)
      losses['ac_loss'] = ac_loss * a['ac_loss_weight']
      total = total + losses['ac_loss']
      losses['g_gan_obj_loss'] = self.gan_g_loss(scores_fake) * (
          a['discriminator_loss_weight'] * a['d_obj_weight'])
      total = total + losses['g_gan_obj_loss']
    if self.d_img is not None:
      if fork:
        main.wait_stream(self._side_stream2)
      else:
        self._freeze(self.d_img, True)
        g_img = self.gan_g_loss(self.d_img(imgs_pred)) * (a['discriminator_loss_weight'] * a['d_img_weight'])
      losses['g_gan_img_loss'] = g_img
      total = total + losses['g_gan_img_loss']
    losses['total_loss'] = total
    # on-device, collective finite flag -> fused Adam's found_inf
    bad = (~torch.isfinite(total.detach().reshape(()))).to(torch.float32)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
      dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.group)
    found_inf.copy_(bad)
    imgs_fake = imgs_pred.detach()

    # The discriminator iteration (train.py:566-592) needs only the detached images and the
    # discriminators' pre-update weights, the generator backward only reads those weights: the two
    # run on two streams (one fork / join inside the captured graph) so that the discriminators'
    # medium-sized launches fill the SMs the generator backward leaves idle (tails of the persistent
    # convolution kernels, the 32-CTA launches of the graph-convolution backward).  The
    # discriminators' all-reduce and Adam steps come after the join: their weights must not move
    # under the generator backward, and one communicator's collectives stay on one stream.
    # Measured (profiles/r02_call20*, r02_call21*): 2487 vs 2247 img/s; one stream per discriminator
    # is slower (2418: more contention on the generator's chain), stream priorities change nothing.
    overlap = (imgs_fake.is_cuda and (self.d_obj is not None or self.d_img is not None)
               and os.environ.get('SG2IM_OVERLAP_DSTEP', '1') != '0')
    if overlap:
      main = torch.cuda.current_stream()
      if self._side_stream is None:
        self._side_stream = torch.cuda.Stream()
      self._side_stream.wait_stream(main)
      with torch.cuda.stream(self._side_stream):
        d_losses = self._d_forward_backward(imgs, imgs_fake, objs, boxes, obj_to_img)

    self.buckets['g'].zero()
    self._g_backward_and_reduce(total)
    self.opts['g'].step()

    if overlap:
      main.wait_stream(self._side_stream)
    else:
      d_losses = self._d_forward_backward(imgs, imgs_fake, objs, boxes, obj_to_img)
    losses.update(d_losses)
    for name, net in (('d_obj', self.d_obj), ('d_img', self.d_img)):
      if net is not None:
        self.buckets[name].all_reduce_mean(self.group, self.opts[name])
        self.opts[name].step()
    return losses, imgs_fake

  def _d_forward_backward(self, imgs, imgs_fake, objs, boxes, obj_to_img):
    """Forward + backward of both discriminators on real and generated images (train.py:566-575,
    581-588); gradients land in their buckets, nothing is reduced or stepped here."""
    out = {}
    if self.d_obj is not None:
      self._freeze(self.d_obj, False)
      s_fake, ac_fake = self.d_obj(imgs_fake, objs, boxes, obj_to_img)
      s_real, ac_real = self.d_obj(imgs, objs, boxes, obj_to_img)
      d_obj_gan = self.gan_d_loss(s_real, s_fake)
      out.update(d_obj_gan_loss=d_obj_gan, d_ac_loss_real=ac_real, d_ac_loss_fake=ac_fake)
      self.buckets['d_obj'].zero()
      (d_obj_gan + ac_real + ac_fake).backward()
    if self.d_img is not None:
      self._freeze(self.d_img, False)
      s_fake = self.d_img(imgs_fake)
      s_real = self.d_img(imgs)
      d_img_gan = self.gan_d_loss(s_real, s_fake)
      out['d_img_gan_loss'] = d_img_gan
      self.buckets['d_img'].zero()
      d_img_gan.backward()
    return out

  # ------------------------------------------------------------------ eager mode
  def _step_eager(self, batch, noise=None):
    a = self.args
    masks = None
    if len(batch) == 6:
      imgs, objs, boxes, triples, obj_to_img, _ = batch
    elif len(batch) == 7:
      imgs, objs, boxes, masks, triples, obj_to_img, _ = batch
    else:
      raise ValueError('batch must have 6 or 7 entries')
    N = imgs.size(0)
    predicates = triples[:, 1]
    for sh in self.split_shadows:          # this step's weights -> bf16 hi / mid operand copies
      sh.refresh()
    if self.zero_arena is not None:
      self.zero_arena.reset()

    # ---------------- generator: train.py:524-560
    imgs_pred, boxes_pred, masks_pred, predicate_scores = self.model(
        objs, triples, obj_to_img, boxes_gt=boxes, masks_gt=masks, num_imgs=N, noise=noise)
    total, losses = self.generator_losses(imgs, imgs_pred, boxes, boxes_pred, masks,
                                          masks_pred, predicates, predicate_scores)
    if self.d_obj is not None:
      self._freeze(self.d_obj, True)
      scores_fake, ac_loss = self.d_obj(imgs_pred, objs, boxes, obj_to_img)
      losses['ac_loss'] = ac_loss * a['ac_loss_weight']
      total = total + losses['ac_loss']
      w = a['discriminator_loss_weight'] * a['d_obj_weight']
      losses['g_gan_obj_loss'] = self.gan_g_loss(scores_fake) * w
      total = total + losses['g_gan_obj_loss']
    if self.d_img is not None:
      self._freeze(self.d_img, True)
      scores_fake = self.d_img(imgs_pred)
      w = a['discriminator_loss_weight'] * a['d_img_weight']
      losses['g_gan_img_loss'] = self.gan_g_loss(scores_fake) * w
      total = total + losses['g_gan_img_loss']
    losses['total_loss'] = total

    keys = list(losses.keys())
    vals = torch.stack([losses[k].reshape(()) for k in keys]).tolist()     # one D2H sync
    out = dict(zip(keys, vals))
    imgs_fake = imgs_pred.detach()
    if not _all_finite(out['total_loss'], self.group):
      print('WARNING: Got loss = NaN, not backpropping')
      self.skipped += 1
      for d in (self.d_obj, self.d_img):
        if d is not None:
          self._freeze(d, False)
      return out, imgs_fake

    self.buckets['g'].zero()
    self._g_backward_and_reduce(total)
    self.opts['g'].step()

    # ---------------- object discriminator: train.py:566-579
    d_vals = {}
    if self.d_obj is not None:
      self._freeze(self.d_obj, False)
      s_fake, ac_fake = self.d_obj(imgs_fake, objs, boxes, obj_to_img)
      s_real, ac_real = self.d_obj(imgs, objs, boxes, obj_to_img)
      d_obj_gan = self.gan_d_loss(s_real, s_fake)
      d_total = d_obj_gan + ac_real + ac_fake
      d_vals.update(d_obj_gan_loss=d_obj_gan, d_ac_loss_real=ac_real, d_ac_loss_fake=ac_fake)
      self.buckets['d_obj'].zero()
      d_total.backward()
      self.buckets['d_obj'].all_reduce_mean(self.group, self.opts['d_obj'])
      self.opts['d_obj'].step()

    # ---------------- image discriminator: train.py:581-592
    if self.d_img is not None:
      self._freeze(self.d_img, False)
      s_fake = self.d_img(imgs_fake)
      s_real = self.d_img(imgs)
      d_img_gan = self.gan_d_loss(s_real, s_fake)
      d_vals['d_img_gan_loss'] = d_img_gan
      self.buckets['d_img'].zero()
      d_img_gan.backward()
      self.buckets['d_img'].all_reduce_mean(self.group, self.opts['d_img'])
      self.opts['d_img'].step()

    if d_vals:
      dk = list(d_vals.keys())
      out.update(zip(dk, torch.stack([d_vals[k].reshape(()).detach() for k in dk]).tolist()))
    return out, imgs_fake
