"""Golden fixtures for the COCO relation synthesis: samples produced by the
UNMODIFIED reference ``CocoSceneGraphDataset.__getitem__`` (sg2im/data/coco.py:225-358) imported
from /root/reference (build container only).

The dataset object is assembled by hand (no COCO annotation files exist here): the attributes
``__getitem__`` reads are set directly, the image is a small PNG written to a temporary
directory, and the two helpers that need absent third-party packages (pycocotools'
``seg_to_mask``, skimage's ``imresize``) are replaced by deterministic stand-ins that paint an
ellipse / half-plane / empty mask per object and resize by nearest neighbour.  Everything after
them — boxes, the masked centroids, Python's ``random`` draws, the geometric predicates, the
``__in_image__`` triples — is the reference's own code.

  python tests/golden/make_golden_coco.py
"""
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from refimport import import_reference  # noqa: E402

PRED_NAMES = ['__in_image__', 'left of', 'right of', 'above', 'below', 'inside', 'surrounding']


def fake_seg_to_mask(seg, width, height):
  """seg = dict(kind, cx, cy, rx, ry): a (height, width) 0/1 array."""
  yy, xx = np.mgrid[0:height, 0:width]
  if seg['kind'] == 'ellipse':
    m = ((xx - seg['cx']) / seg['rx']) ** 2 + ((yy - seg['cy']) / seg['ry']) ** 2 <= 1.0
  elif seg['kind'] == 'half':
    m = (xx - seg['cx']) * seg['rx'] + (yy - seg['cy']) * seg['ry'] >= 0
  elif seg['kind'] == 'empty':
    m = np.zeros((height, width), dtype=bool)
  else:
    m = np.ones((height, width), dtype=bool)
  return m.astype(np.uint8)


def fake_imresize(a, size, mode='constant'):
  h, w = a.shape
  ys = np.minimum((np.arange(size[0]) + 0.5) * h / size[0], h - 1).astype(int)
  xs = np.minimum((np.arange(size[1]) + 0.5) * w / size[1], w - 1).astype(int)
  return a[ys][:, xs]


def make_dataset(coco, tmp, rnd, n_images, mask_size):
  import PIL.Image
  ds = object.__new__(coco.CocoSceneGraphDataset)
  ds.image_dir = tmp
  ds.mask_size = mask_size
  ds.max_samples = None
  ds.include_relationships = True
  ds.image_size = (16, 16)
  ds.transform = lambda im: torch.zeros(3, 16, 16)
  ds.vocab = {'object_name_to_idx': {'__image__': 0},
              'pred_name_to_idx': {n: i for i, n in enumerate(PRED_NAMES)}}
  ds.image_ids, ds.image_id_to_filename, ds.image_id_to_objects = [], {}, {}
  WW, HH = 96, 64
  PIL.Image.new('RGB', (WW, HH)).save(os.path.join(tmp, 'im.png'))
  for i in range(n_images):
    n_obj = [0, 1, 2, 3, 5, 8, 8, 6][i % 8]
    objects = []
    for k in range(n_obj):
      w, h = rnd.uniform(6, 60), rnd.uniform(6, 40)
      x, y = rnd.uniform(0, WW - w), rnd.uniform(0, HH - h)
      if k and rnd.random() < 0.25:                     # nested inside the previous box
        px, py, pw, ph = objects[-1]['bbox']
        x, y, w, h = px + 0.2 * pw, py + 0.2 * ph, 0.5 * pw, 0.5 * ph
      kind = rnd.choice(['ellipse', 'ellipse', 'half', 'full', 'empty'])
      seg = dict(kind=kind, cx=x + rnd.uniform(0.2, 0.8) * w, cy=y + rnd.uniform(0.2, 0.8) * h,
                 rx=max(1.0, rnd.uniform(0.1, 0.5) * w) if kind == 'ellipse' else rnd.uniform(-1, 1),
                 ry=max(1.0, rnd.uniform(0.1, 0.5) * h) if kind == 'ellipse' else rnd.uniform(-1, 1))
      objects.append({'category_id': rnd.randint(1, 20), 'bbox': [x, y, w, h], 'segmentation': seg})
    ds.image_ids.append(i)
    ds.image_id_to_filename[i] = 'im.png'
    ds.image_id_to_objects[i] = objects
  return ds


def main():
  assert import_reference() is not None, 'reference tree not found'
  import sg2im.data.coco as coco
  coco.seg_to_mask = fake_seg_to_mask
  coco.imresize = fake_imresize
  rnd = random.Random(5)
  samples = []
  with tempfile.TemporaryDirectory() as tmp:
    for mask_size in (16, 5):
      ds = make_dataset(coco, tmp, rnd, 16, mask_size)
      for idx in range(len(ds.image_ids)):
        seed = 1000 * mask_size + idx
        random.seed(seed)                               # the module-level generator coco.py draws from
        _, objs, boxes, masks, triples = ds[idx]
        samples.append(dict(seed=seed, objs=objs, boxes=boxes, masks=masks.to(torch.uint8),   # 0/1: stored as bytes
                            triples=triples))
  out = dict(pred_names=PRED_NAMES, samples=samples)
  path = os.path.join(HERE, 'coco_rel.pt')
  torch.save(out, path)
  n_rel = sum(int((s['triples'][:, 1] != 0).sum()) if s['triples'].dim() == 2 else 0 for s in samples)
  print('wrote %s: %d samples, %d geometric triples, %d bytes' % (path, len(samples), n_rel,
                                                                 os.path.getsize(path)))


if __name__ == '__main__':
  main()
