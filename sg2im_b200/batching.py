"""Batch assembly in the reference's collate format, staged for upload in two copies.

The reference's DataLoader hands ``scripts/train.py:515-523`` a tuple built by
``vg_collate_fn`` (sg2im/data/vg.py:144-186: imgs, objs, boxes, triples,
obj_to_img, triple_to_img) or ``coco_collate_fn`` (sg2im/data/coco.py:376-419:
the same with ``masks`` after ``boxes``), i.e. a dozen ``torch.cat`` results that
``train.py:516`` then moves to the GPU one ``.cuda()`` at a time.

Here the same tuple is laid out inside TWO staging buffers -- one float32 (imgs,
boxes), one int64 (objs, [masks,] triples, obj_to_img, triple_to_img) -- that
can live in pinned memory, so a batch reaches the device in two asynchronous
copies and the tuple members are views into them.  The values, dtypes, shapes
and order are those of the reference's collate functions (tests compare
bit-for-bit against both of them).
"""
import torch


def _numel(shape):
  n = 1
  for s in shape:
    n *= int(s)
  return n


class StagedBatch(object):
  """The collate tuple plus the two flat buffers it is carved from."""

  def __init__(self, floats, ints, layout):
    self.floats, self.ints, self.layout = floats, ints, layout

  def _views(self, floats, ints):
    out = []
    for kind, off, shape in self.layout:
      src = floats if kind == 'f' else ints
      out.append(src[off:off + _numel(shape)].view(shape))
    return tuple(out)

  def tensors(self):
    """(imgs, objs, boxes, [masks,] triples, obj_to_img, triple_to_img)."""
    return self._views(self.floats, self.ints)

  def to(self, device, non_blocking=True):
    """Two host->device copies (asynchronous when the staging buffers are
    pinned); returns the tuple as views of the device copies."""
    f = self.floats.to(device, non_blocking=non_blocking)
    i = self.ints.to(device, non_blocking=non_blocking)
    return self._views(f, i)

  def h2d_bytes(self):
    return self.floats.numel() * 4 + self.ints.numel() * 8


def collate(samples, pin=False):
  """samples: list of ``(image, objs, boxes, triples)`` (Visual Genome,
  vg.py:141) or ``(image, objs, boxes, masks, triples)`` (COCO, coco.py:358)
  with per-image object indices in ``triples``.  Returns a StagedBatch.

  Object indices are offset by the number of objects that precede the image in
  the batch; a COCO sample whose ``objs`` or ``triples`` is 0-dimensional
  contributes its image only (coco.py:394-395).  A sample without triples
  (shape (0,) as ``torch.LongTensor([])`` gives) is accepted as zero triples --
  the reference's ``triples[:, 0]`` raises on it."""
  if len(samples) == 0:
    raise ValueError('collate: empty batch')
  with_masks = len(samples[0]) == 5
  rows = []                       # (image index, objs, boxes, masks, triples)
  for i, s in enumerate(samples):
    if len(s) != (5 if with_masks else 4):
      raise ValueError('collate: mixed sample formats in one batch')
    objs, boxes = s[1], s[2]
    masks = s[3] if with_masks else None
    triples = s[-1]
    if with_masks and (objs.dim() == 0 or triples.dim() == 0):
      continue
    if triples.dim() == 1 and triples.numel() == 0:
      triples = triples.view(0, 3)
    rows.append((i, objs, boxes, masks, triples))

  N = len(samples)
  img_shape = tuple(samples[0][0].shape)
  O = sum(r[1].size(0) for r in rows)
  T = sum(r[4].size(0) for r in rows)
  if with_masks and not rows:
    raise RuntimeError('collate: no sample in the batch has objects')   # torch.cat([]) in the reference
  mshape = tuple(rows[0][3].shape[1:]) if with_masks else None

  layout, nf, ni = [], 0, 0

  def add(kind, shape):
    nonlocal nf, ni
    n = _numel(shape)
    if kind == 'f':
      layout.append(('f', nf, shape)); nf += n
    else:
      layout.append(('i', ni, shape)); ni += n

  add('f', (N,) + img_shape)                  # imgs
  add('i', (O,))                              # objs
  add('f', (O, 4))                            # boxes
  if with_masks:
    add('i', (O,) + mshape)                   # masks (int64 in the reference's COCO pipeline)
  add('i', (T, 3))                            # triples
  add('i', (O,))                              # obj_to_img
  add('i', (T,))                              # triple_to_img

  floats = torch.empty(nf, dtype=torch.float32, pin_memory=pin)
  ints = torch.empty(ni, dtype=torch.int64, pin_memory=pin)
  batch = StagedBatch(floats, ints, layout)
  views = batch.tensors()
  if with_masks:
    imgs, objs_o, boxes_o, masks_o, triples_o, o2i, t2i = views
  else:
    imgs, objs_o, boxes_o, triples_o, o2i, t2i = views
    masks_o = None

  for i, s in enumerate(samples):
    imgs[i].copy_(s[0])
  o0 = t0 = 0
  for i, objs, boxes, masks, triples in rows:
    o1, t1 = o0 + objs.size(0), t0 + triples.size(0)
    objs_o[o0:o1].copy_(objs)
    boxes_o[o0:o1].copy_(boxes)
    if masks_o is not None:
      masks_o[o0:o1].copy_(masks)
    if t1 > t0:
      dst = triples_o[t0:t1]
      dst.copy_(triples)
      dst[:, 0] += o0
      dst[:, 2] += o0
    o2i[o0:o1] = i
    t2i[t0:t1] = i
    o0, t0 = o1, t1
  return batch


def vg_collate_fn(batch):
  """Drop-in for sg2im.data.vg.vg_collate_fn (tuple of CPU tensors)."""
  return collate(batch).tensors()


def coco_collate_fn(batch):
  """Drop-in for sg2im.data.coco.coco_collate_fn (tuple of CPU tensors)."""
  return collate(batch).tensors()


def uncollate(batch):
  """Inverse of ``collate`` for the 6-tuple (what vg.py:189-215 ``vg_uncollate_fn``
  returns): per-image ``(img, objs, boxes, triples)`` with image-local object
  indices.  Relies on the collate invariant that objects and triples are
  grouped by image in ascending order."""
  imgs, objs, boxes, triples, obj_to_img, triple_to_img = batch
  N = imgs.size(0)
  n_obj = torch.bincount(obj_to_img, minlength=N).tolist()
  n_tri = torch.bincount(triple_to_img, minlength=N).tolist()
  first_obj = [0]
  for n in n_obj[:-1]:
    first_obj.append(first_obj[-1] + n)
  per_img = zip(imgs, objs.split(n_obj), boxes.split(n_obj), triples.split(n_tri), first_obj)
  out = []
  for img, o, b, t, base in per_img:
    t = t.clone()
    t[:, 0::2] -= base                       # columns 0 (subject) and 2 (object)
    out.append((img, o, b, t))
  return out


# --------------------------------------------------------------------------
# COCO scene-graph synthesis for a whole batch (sg2im/data/coco.py:294-356)
# --------------------------------------------------------------------------
COCO_PREDICATES = ('left of', 'right of', 'above', 'below', 'inside', 'surrounding', '__in_image__')


def coco_relation_draws(obj_counts, include_relationships=True, rng=None):
  """The random part of the reference's graph synthesis, on the host.  obj_counts: objects
  per image INCLUDING the trailing __image__ object.  For every image with at least two real
  objects, per real object in order: ``other = random.choice(the other real objects)`` then
  ``random.random() > 0.5`` keeps (cur, other) as (subject, object) — the same two calls in
  the same order as coco.py:319-327, so with the same seed and sample order the draws are the
  reference's.  Returns (partner int64 [O] of GLOBAL indices, -1 where unused; swap uint8 [O],
  1 = the partner is the subject; obj_off int64 [N+1]; trip_off int64 [N+1])."""
  import random as _random
  rng = _random if rng is None else rng
  O = int(sum(obj_counts))
  partner = [-1] * O
  swap = [0] * O
  obj_off, trip_off = [0], [0]
  for c in obj_counts:
    c = int(c)
    if c < 1:
      raise ValueError('every image needs its __image__ object')
    base, n_real = obj_off[-1], c - 1
    n_rel = n_real if (include_relationships and n_real > 1) else 0
    for pos in range(n_rel):
      k = rng.randrange(n_real - 1)
      partner[base + pos] = base + (k if k < pos else k + 1)
      swap[base + pos] = 0 if rng.random() > 0.5 else 1
    obj_off.append(base + c)
    trip_off.append(trip_off[-1] + n_rel + n_real)
  return (torch.tensor(partner, dtype=torch.int64), torch.tensor(swap, dtype=torch.uint8),
          torch.tensor(obj_off, dtype=torch.int64), torch.tensor(trip_off, dtype=torch.int64))


def coco_relations(boxes, masks, obj_counts, vocab, include_relationships=True, rng=None,
                   device=None):
  """Scene graphs of a collated COCO batch in two kernel launches (sg2im_coco_relations).

  boxes (O,4) float32 and masks (O,M,M) int64 of all objects of the batch, grouped by image
  with the __image__ object last in each (what coco.py:286-292 builds per sample and
  coco_collate_fn concatenates); host tensors are uploaded (two copies), device tensors are
  used in place.  Returns (triples (T,3) int64 with batch-global object indices, triple_to_img
  (T,), obj_to_img (O,)) on the device — the three graph members of the collate tuple
  (coco.py:411-418).  Per image the triple order is the reference's: geometric triples in
  object order, then the __in_image__ triples."""
  from . import _lib, ops
  if device is None:
    device = boxes.device if boxes.is_cuda else torch.device('cuda')
  partner, swap, obj_off, trip_off = coco_relation_draws(obj_counts, include_relationships, rng)
  O, T = int(obj_off[-1]), int(trip_off[-1])
  if boxes.shape != (O, 4) or masks.dim() != 3 or masks.size(0) != O:
    raise ValueError('coco_relations: boxes %s / masks %s do not match %d objects'
                     % (tuple(boxes.shape), tuple(masks.shape), O))
  if masks.dtype != torch.int64 or boxes.dtype != torch.float32:
    raise ValueError('coco_relations: boxes float32 and masks int64 expected (coco.py:291-292)')
  names = vocab['pred_name_to_idx']
  pred_ids = torch.tensor([names[p] for p in COCO_PREDICATES], dtype=torch.int64)   # host array
  counts = torch.as_tensor([int(c) for c in obj_counts], dtype=torch.int64)
  # one staging buffer for the five small index arrays
  obj_to_img_h = torch.repeat_interleave(torch.arange(counts.numel(), dtype=torch.int64), counts)
  ints = torch.cat([obj_off, trip_off, obj_to_img_h, partner]).to(device, non_blocking=True)
  n1 = obj_off.numel()
  d_obj_off, d_trip_off = ints[:n1], ints[n1:2 * n1]
  d_obj_to_img, d_partner = ints[2 * n1:2 * n1 + O], ints[2 * n1 + O:]
  d_swap = swap.to(device, non_blocking=True)
  boxes = ops._chk(boxes.to(device, non_blocking=True).contiguous(), name='boxes')
  masks = ops._chk(masks.to(device, non_blocking=True).contiguous(), dtype=torch.int64, name='masks')
  centers = torch.empty(max(O, 1), 2, dtype=torch.float32, device=device)
  triples = torch.empty(T, 3, dtype=torch.int64, device=device)
  triple_to_img = torch.empty(T, dtype=torch.int64, device=device)
  if T == 0:                                           # only __image__ objects: nothing to write
    return triples, triple_to_img, d_obj_to_img
  _lib.call('sg2im_coco_relations', boxes.data_ptr(), masks.data_ptr(), masks.size(1), masks.size(2),
            d_obj_off.data_ptr(), d_trip_off.data_ptr(), d_obj_to_img.data_ptr(), d_partner.data_ptr(),
            d_swap.data_ptr(), O, pred_ids.data_ptr(), centers.data_ptr(), triples.data_ptr(),
            triple_to_img.data_ptr(), ops._stream())
  return triples, triple_to_img, d_obj_to_img
