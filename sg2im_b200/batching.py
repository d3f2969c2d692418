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
