"""batching.collate against the reference's own collate functions
(sg2im/data/vg.py:144-186, sg2im/data/coco.py:376-419), bit for bit."""
import pytest
import torch

from refimport import have_reference, import_reference
from sg2im_b200 import batching

HAVE_REF = have_reference()


def _vg_samples(seed, n_imgs, with_masks=False, M=8):
  g = torch.Generator().manual_seed(seed)
  out = []
  for i in range(n_imgs):
    O = int(torch.randint(2, 7, (1,), generator=g))
    R = int(torch.randint(0, 5, (1,), generator=g))
    img = torch.randn(3, 16, 16, generator=g)
    objs = torch.cat([torch.randint(1, 20, (O - 1,), generator=g), torch.zeros(1, dtype=torch.int64)])
    boxes = torch.rand(O, 4, generator=g)
    rel = torch.stack([torch.randint(0, O - 1, (R,), generator=g),
                       torch.randint(1, 6, (R,), generator=g),
                       torch.randint(0, O - 1, (R,), generator=g)], dim=1)
    in_img = torch.stack([torch.arange(O - 1), torch.zeros(O - 1, dtype=torch.int64),
                          torch.full((O - 1,), O - 1, dtype=torch.int64)], dim=1)
    triples = torch.cat([rel, in_img])
    if with_masks:
      masks = torch.randint(0, 2, (O, M, M), generator=g)
      out.append((img, objs, boxes, masks, triples))
    else:
      out.append((img, objs, boxes, triples))
  return out


def _same(a, b):
  assert len(a) == len(b)
  for x, y in zip(a, b):
    assert x.dtype == y.dtype and x.shape == y.shape
    assert torch.equal(x, y)


@pytest.mark.skipif(not HAVE_REF, reason='reference tree not present on this box')
@pytest.mark.parametrize('seed', [0, 1, 2])
def test_vg_collate_matches_reference(seed):
  import_reference()                                    # stubs h5py etc.
  from sg2im.data.vg import vg_collate_fn, vg_uncollate_fn
  samples = _vg_samples(seed, 5)
  want = vg_collate_fn(samples)
  got = batching.vg_collate_fn(samples)
  _same(got, want)
  # and back
  for a, b in zip(batching.uncollate(got), vg_uncollate_fn(want)):
    _same(a, b)


@pytest.mark.skipif(not HAVE_REF, reason='reference tree not present on this box')
@pytest.mark.parametrize('seed', [0, 3])
def test_coco_collate_matches_reference(seed):
  import_reference()
  from sg2im.data.coco import coco_collate_fn
  samples = _vg_samples(seed, 4, with_masks=True)
  # a sample the reference skips (0-dim objs): image kept, no objects
  img = torch.randn(3, 16, 16)
  samples.insert(2, (img, torch.tensor(3), torch.zeros(4), torch.zeros(8, 8, dtype=torch.int64),
                     torch.tensor(0)))
  want = coco_collate_fn(samples)
  got = batching.coco_collate_fn(samples)
  _same(got, want)
  assert got[5].max().item() == len(samples) - 1 and 2 not in got[5].tolist()


def test_collate_layout_and_staging():
  samples = _vg_samples(7, 3)
  b = batching.collate(samples)
  imgs, objs, boxes, triples, o2i, t2i = b.tensors()
  assert imgs.shape == (3, 3, 16, 16) and boxes.shape == (objs.numel(), 4)
  # tuple members are views of the two staging buffers: one copy each moves the batch
  assert imgs.untyped_storage().data_ptr() == boxes.untyped_storage().data_ptr()
  assert objs.untyped_storage().data_ptr() == triples.untyped_storage().data_ptr() \
      == o2i.untyped_storage().data_ptr() == t2i.untyped_storage().data_ptr()
  assert b.h2d_bytes() == imgs.numel() * 4 + boxes.numel() * 4 + 8 * (
      2 * objs.numel() + 4 * t2i.numel())
  # .to() rebuilds the same tuple from the copies
  for x, y in zip(b.to('cpu'), b.tensors()):
    assert torch.equal(x, y)
  # triples index into the batch-global object list, grouped by image
  assert torch.equal(o2i[triples[:, 0]], t2i) and torch.equal(o2i[triples[:, 2]], t2i)
  assert bool((o2i[1:] >= o2i[:-1]).all())


def test_collate_accepts_imageless_triples_and_rejects_mixed():
  img = torch.zeros(3, 4, 4)
  lone = (img, torch.zeros(1, dtype=torch.int64), torch.tensor([[0., 0., 1., 1.]]), torch.LongTensor([]))
  other = _vg_samples(1, 1)[0]
  other = (other[0][:, :4, :4],) + other[1:]
  out = batching.collate([lone, other]).tensors()
  assert out[3].shape[1] == 3 and out[4][0].item() == 0 and out[5].min().item() == 1
  assert out[3][:, 0].min().item() >= 1                 # offset by the lone image's one object
  with pytest.raises(ValueError):
    batching.collate([lone, _vg_samples(1, 1, with_masks=True)[0]])
  with pytest.raises(ValueError):
    batching.collate([])
