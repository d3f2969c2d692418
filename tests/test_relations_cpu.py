"""COCO scene-graph synthesis without a GPU.

1. the oracle restatement (oracle/relations_oracle.py) against samples produced by the UNMODIFIED
   reference ``CocoSceneGraphDataset.__getitem__`` (tests/golden/coco_rel.pt, seeded ``random``);
2. the product's host half (batching.coco_relation_draws: Python ``random`` in the reference's
   call order) against the oracle's draws;
3. the product's device half — csrc/relations.cu, the file nvcc compiles, built for the host with
   -DSG2IM_EMUL and driven through the shipped Python wrapper (batching.coco_relations) on the
   emulated device — against the oracle, on the golden samples collated into one batch and on
   random batches; plus the exact equivalence of the kernel's comparison-based sector test with
   the reference's atan2 test on diagonals, axes and signed zeros.
"""
import math
import os
import random
import shutil

import pytest
import torch

from conftest import ROOT
from oracle import relations_oracle as RO
from sg2im_b200 import batching

GOLD = os.path.join(ROOT, 'tests', 'golden', 'coco_rel.pt')


def gold():
  d = torch.load(GOLD)
  for s in d['samples']:
    s['masks'] = s['masks'].long()
  return d


def vocab_of(names):
  return {'object_name_to_idx': {'__image__': 0}, 'pred_name_to_idx': {n: i for i, n in enumerate(names)}}


def test_oracle_reproduces_the_reference_samples():
  d = gold()
  idx = {n: i for i, n in enumerate(d['pred_names'])}
  kinds = set()
  for s in d['samples']:
    rng = random.Random(s['seed'])
    triples, margins = RO.sample_triples(s['objs'], s['boxes'], s['masks'], idx, rng=rng)
    ref = s['triples'].tolist() if s['triples'].dim() == 2 else []
    assert len(triples) == len(ref)
    for t, r, m in zip(triples, ref, margins):
      if m is not None and m < 1e-5:
        continue                                         # see the oracle's floating-point note
      assert t == r
    kinds.update(d['pred_names'][t[1]] for t in ref)
  assert kinds == set(d['pred_names'])                   # every predicate occurs in the fixture


def expected_batch(samples, idx, seed, include=True):
  """The oracle per sample with ONE generator across the batch + coco_collate_fn's offsets."""
  rng = random.Random(seed)
  triples, t2i, margins, base = [], [], [], 0
  for n, s in enumerate(samples):
    t, m = RO.sample_triples(s['objs'], s['boxes'], s['masks'], idx, include_relationships=include, rng=rng)
    triples += [[a + base, p, b + base] for a, p, b in t]
    t2i += [n] * len(t)
    margins += m
    base += s['objs'].numel()
  return torch.tensor(triples, dtype=torch.int64).view(-1, 3), torch.tensor(t2i, dtype=torch.int64), margins


def test_host_draws_follow_the_reference_call_order():
  d = gold()
  samples = [s for s in d['samples'] if s['masks'].size(1) == 16]
  idx = {n: i for i, n in enumerate(d['pred_names'])}
  want, _, _ = expected_batch(samples, idx, 77)
  counts = [s['objs'].numel() for s in samples]
  partner, swap, obj_off, trip_off = batching.coco_relation_draws(counts, rng=random.Random(77))
  assert obj_off.tolist() == [0] + torch.tensor(counts).cumsum(0).tolist()
  assert int(trip_off[-1]) == want.size(0)
  for n, c in enumerate(counts):
    base, t0 = int(obj_off[n]), int(trip_off[n])
    n_real = c - 1
    if n_real < 2:
      assert (partner[base:base + c] == -1).all()
      continue
    for pos in range(n_real):
      s, _, o = want[t0 + pos].tolist()
      i, other = base + pos, int(partner[base + pos])
      assert (s, o) == ((other, i) if swap[i] else (i, other))
  # switched off (coco.py:321): no draws at all, only __in_image__ triples
  rng = random.Random(3)
  state = rng.getstate()
  partner, swap, _, trip_off = batching.coco_relation_draws(counts, include_relationships=False, rng=rng)
  assert rng.getstate() == state and (partner == -1).all()
  assert int(trip_off[-1]) == sum(c - 1 for c in counts)
  with pytest.raises(ValueError):
    batching.coco_relation_draws([3, 0])


def test_reference_sector_thresholds_on_the_diagonals():
  """The kernel replaces atan2 by comparisons; that is exact provided atan2 returns exactly
  +-pi/4, +-3pi/4 on the diagonals (it does in IEEE libm; asserted here for this host)."""
  for a in (1.0, 0.3, 1e-7, 0.123456789, float(torch.tensor(0.1, dtype=torch.float32))):
    assert math.atan2(a, a) == math.pi / 4 and math.atan2(-a, a) == -math.pi / 4
    assert math.atan2(a, -a) == 3 * math.pi / 4 and math.atan2(-a, -a) == -3 * math.pi / 4


needs_gxx = pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++ (C++20)')


@pytest.fixture(scope='module')
def emul():
  from emul_device import emulated_device
  with emulated_device() as lib:
    yield lib


def run_product(samples, names, seed, include=True):
  boxes = torch.cat([s['boxes'] for s in samples])
  masks = torch.cat([s['masks'] for s in samples])
  counts = [s['objs'].numel() for s in samples]
  return batching.coco_relations(boxes, masks, counts, vocab_of(names), include_relationships=include,
                                 rng=random.Random(seed), device=torch.device('cpu'))


def assert_same(got, want, margins):
  triples, t2i = got
  assert triples.shape == want[0].shape and torch.equal(t2i, want[1])
  close = torch.tensor([m is not None and m < 1e-5 for m in margins])
  assert close.float().mean() < 0.01
  assert torch.equal(triples[~close], want[0][~close])
  assert torch.equal(triples[close][:, 0::2], want[0][close][:, 0::2])       # only the predicate may differ


@needs_gxx
@pytest.mark.parametrize('mask_size', [16, 5])
@pytest.mark.parametrize('include', [True, False])
def test_kernel_on_the_reference_samples_as_one_batch(emul, mask_size, include):
  d = gold()
  samples = [s for s in d['samples'] if s['masks'].size(1) == mask_size]
  idx = {n: i for i, n in enumerate(d['pred_names'])}
  want_t, want_i, margins = expected_batch(samples, idx, 11, include)
  triples, t2i, o2i = run_product(samples, d['pred_names'], 11, include)
  assert_same((triples, t2i), (want_t, want_i), margins)
  assert torch.equal(o2i, torch.repeat_interleave(torch.arange(len(samples)),
                                                  torch.tensor([s['objs'].numel() for s in samples])))
  # this is what the reference's collate makes of its own per-sample outputs
  if include:
    ref = [s['triples'] for s in samples if s['triples'].dim() == 2]
    assert want_t.size(0) == sum(t.size(0) for t in ref)


@needs_gxx
@pytest.mark.parametrize('seed,MH,MW', [(0, 16, 16), (1, 7, 3), (2, 1, 1), (3, 32, 32)])
def test_kernel_on_random_batches(emul, seed, MH, MW):
  g = torch.Generator().manual_seed(seed)
  names = ['surrounding', '__in_image__', 'below', 'left of', 'inside', 'above', 'right of']   # shuffled ids
  samples = []
  for n in range(9):
    c = [1, 2, 3, 9, 4, 1, 6, 2, 12][n]
    xy = torch.rand(c, 2, generator=g) * 0.6
    wh = torch.rand(c, 2, generator=g) * 0.4
    boxes = torch.cat([xy, xy + wh], 1)
    for k in range(1, c - 1):
      if k % 3 == 0:                                      # strictly nested pair
        boxes[k] = torch.stack([boxes[k - 1][0] + 0.01, boxes[k - 1][1] + 0.01,
                                boxes[k - 1][2] - 0.01, boxes[k - 1][3] - 0.01])
    boxes[-1] = torch.tensor([0., 0., 1., 1.])
    masks = (torch.rand(c, MH, MW, generator=g) < 0.4).long()
    masks[::4] = 0                                         # empty masks: box centre
    masks[-1] = 1
    if c > 2:
      masks[1] = 2 * masks[1]                              # values other than 1 do not count
    objs = torch.cat([torch.randint(1, 30, (c - 1,), generator=g), torch.zeros(1, dtype=torch.int64)])
    samples.append(dict(objs=objs, boxes=boxes, masks=masks))
  idx = {n: i for i, n in enumerate(names)}
  want_t, want_i, margins = expected_batch(samples, idx, seed)
  triples, t2i, _ = run_product(samples, names, seed)
  assert_same((triples, t2i), (want_t, want_i), margins)


@needs_gxx
def test_kernel_sector_test_equals_atan2_on_exact_cases(emul):
  """Point boxes with empty masks make the centroid difference an exact, chosen fp32 pair."""
  names = list(batching.COCO_PREDICATES)
  idx = {n: i for i, n in enumerate(names)}
  f = lambda v: float(torch.tensor(v, dtype=torch.float32))      # noqa: E731
  cases = []
  vals = [0.0, f(0.1), f(0.25), f(0.3), f(0.1) * (1 + 2.0 ** -23), f(0.1) * (1 - 2.0 ** -24)]
  for dx in vals + [-v for v in vals]:
    for dy in vals + [-v for v in vals]:
      cases.append((dx, dy))
  base = 0.5
  samples = []
  for dx, dy in cases:
    # subject = point at (base + dx, base + dy), object = point at (base, base); force the order
    ps = torch.tensor([base + dx, base + dy], dtype=torch.float32)
    boxes = torch.stack([torch.cat([ps, ps]), torch.tensor([base, base, base, base]),
                         torch.tensor([0., 0., 1., 1.])])
    samples.append(dict(objs=torch.tensor([1, 2, 0]), boxes=boxes, masks=torch.zeros(3, 4, 4, dtype=torch.int64)))
  want_t, want_i, margins = expected_batch(samples, idx, 5)
  triples, t2i, _ = run_product(samples, names, 5)
  assert torch.equal(t2i, want_i)
  assert torch.equal(triples, want_t)                     # no margin exclusion: exact agreement
  seen = {names[p] for p in triples[:, 1].tolist()}
  assert {'left of', 'right of', 'above', 'below'} <= seen


@needs_gxx
def test_wrapper_argument_checks(emul):
  v = vocab_of(list(batching.COCO_PREDICATES))
  b, m = torch.rand(3, 4), torch.ones(3, 4, 4, dtype=torch.int64)
  with pytest.raises(ValueError):
    batching.coco_relations(b, m, [2], v, device=torch.device('cpu'))
  with pytest.raises(ValueError):
    batching.coco_relations(b, m.float(), [3], v, device=torch.device('cpu'))
  t, ti, oi = batching.coco_relations(b[:1], m[:1], [1], v, device=torch.device('cpu'))   # only __image__
  assert t.shape == (0, 3) and ti.numel() == 0 and oi.tolist() == [0]
