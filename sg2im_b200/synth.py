"""Synthetic scene-graph batches with exactly the reference's collate format.

Tuple layout follows sg2im/data/vg.py:144-186 (6-tuple, no masks) and
sg2im/data/coco.py:376-419 (7-tuple with (O,M,M) int64 masks): objects grouped
by image in ascending order, the last object of every image is ``__image__``
(category 0, box [0,0,1,1]) and every real object has an
``(i, __in_image__=0, image_idx)`` triple (coco.py:340-358, vg.py:122-139).
Distributions are the ones SURVEY.md §8(d) fixes for the benchmark.
"""
import torch


def make_vocab(num_objs, num_preds):
  """Minimal vocab dict with the keys Sg2ImModel / AcDiscriminator read
  (sg2im/model.py:46-47,196-205; sg2im/discriminators.py:63)."""
  obj_names = ['__image__'] + ['obj%d' % i for i in range(1, num_objs)]
  pred_names = ['__in_image__'] + ['pred%d' % i for i in range(1, num_preds)]
  return {
    'object_idx_to_name': obj_names,
    'object_name_to_idx': {n: i for i, n in enumerate(obj_names)},
    'pred_idx_to_name': pred_names,
    'pred_name_to_idx': {n: i for i, n in enumerate(pred_names)},
  }


CONFIGS = {
  # name: (N per GPU, real objs/img, rels/img, H, W, with gt masks, num_objs, num_preds)
  'coco64': dict(N=32, objs_per_img=6, rels_per_img=6, image_size=(64, 64),
                 masks=True, num_objs=184, num_preds=7),
  'vg128': dict(N=32, objs_per_img=9, rels_per_img=5, image_size=(128, 128),
                masks=False, num_objs=179, num_preds=46),
  # VG-256: the reference ships no 256x256 recipe; SURVEY.md §8(d) assumes one extra refinement
  # stage (dims 1024,512,256,128,64,32), batch 16 per GPU
  'vg256': dict(N=16, objs_per_img=9, rels_per_img=5, image_size=(256, 256),
                masks=False, num_objs=179, num_preds=46,
                refinement_dims=(1024, 512, 256, 128, 64, 32)),
  'dense128': dict(N=64, objs_per_img=32, rels_per_img=32, image_size=(128, 128),
                   masks=False, num_objs=179, num_preds=46),
  'tiny32': dict(N=4, objs_per_img=3, rels_per_img=2, image_size=(32, 32),
                 masks=False, num_objs=9, num_preds=5),
}


def synth_batch(N, objs_per_img, rels_per_img, image_size, num_objs, num_preds,
                masks=False, mask_size=16, seed=0, **_):
  """Returns the collate tuple (CPU tensors); extra config keys are ignored."""
  g = torch.Generator().manual_seed(seed)
  H, W = image_size
  R = objs_per_img
  imgs = torch.randn(N, 3, H, W, generator=g)
  objs, boxes, triples, obj_to_img, triple_to_img = [], [], [], [], []
  for n in range(N):
    base = n * (R + 1)
    cats = torch.randint(1, num_objs, (R,), generator=g)
    xy = torch.rand(R, 2, generator=g) * 0.6
    wh = torch.rand(R, 2, generator=g) * 0.25 + 0.15
    bx = torch.cat([xy, xy + wh], dim=1)
    objs.append(torch.cat([cats, torch.zeros(1, dtype=torch.int64)]))
    boxes.append(torch.cat([bx, torch.tensor([[0., 0., 1., 1.]])]))
    obj_to_img.append(torch.full((R + 1,), n, dtype=torch.int64))
    if rels_per_img > 0 and R > 1:
      s = torch.randint(0, R, (rels_per_img,), generator=g)
      d = torch.randint(1, R, (rels_per_img,), generator=g)
      o = (s + d) % R                                    # s != o
      p = torch.randint(1, max(num_preds, 2), (rels_per_img,), generator=g)
      triples.append(torch.stack([s + base, p, o + base], dim=1))
    in_img = torch.stack([torch.arange(R) + base,
                          torch.zeros(R, dtype=torch.int64),
                          torch.full((R,), base + R, dtype=torch.int64)], dim=1)
    triples.append(in_img)
    triple_to_img.append(torch.full(((rels_per_img if R > 1 else 0) + R,), n,
                                    dtype=torch.int64))
  objs = torch.cat(objs)
  boxes = torch.cat(boxes)
  triples = torch.cat(triples)
  obj_to_img = torch.cat(obj_to_img)
  triple_to_img = torch.cat(triple_to_img)
  if masks:
    m = torch.randint(0, 2, (objs.size(0), mask_size, mask_size), generator=g)
    return (imgs, objs, boxes, m, triples, obj_to_img, triple_to_img)
  return (imgs, objs, boxes, triples, obj_to_img, triple_to_img)


def synth_config(name, seed=0, **overrides):
  cfg = dict(CONFIGS[name])
  cfg.update(overrides)
  return synth_batch(seed=seed, **cfg), cfg
