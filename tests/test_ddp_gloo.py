"""world_size-2 gloo tests (CPU) of the data-parallel host logic: the flat
gradient bucket all-reduce (sum, / world) and the collective finite-loss flag
(SURVEY.md §8e).  No GPU, no kernels: plain tensors stand in for gradients."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from sg2im_b200.train_step import FlatGrads, _all_finite
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    bucket = FlatGrads(lin.parameters())
    # .grad tensors are views into one flat buffer
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in lin.parameters())
    assert bucket.flat.numel() == sum(p.numel() for p in lin.parameters())
    x = torch.full((5, 4), float(rank + 1))
    lin(x).sum().backward()                       # autograd accumulates into the views
    local = bucket.flat.clone()
    bucket.all_reduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    want = sum(gathered) / world
    ok_avg = torch.allclose(bucket.flat, want, rtol=0, atol=1e-6)
    same = [torch.zeros_like(bucket.flat) for _ in range(world)]
    dist.all_gather(same, bucket.flat)
    ok_same = all(torch.equal(same[0], s) for s in same)
    bucket.zero()
    ok_zero = float(bucket.flat.abs().sum()) == 0.0 and all(
        float(p.grad.abs().sum()) == 0.0 for p in lin.parameters())
    # finite flag: one rank sees NaN -> every rank must skip
    f_all = _all_finite(1.0)
    f_one = _all_finite(float('nan') if rank == 1 else 1.0)
    q.put((rank, ok_avg, ok_same, ok_zero, f_all, f_one))
  finally:
    dist.destroy_process_group()


def test_flat_bucket_allreduce_and_collective_finite_flag():
  world, port = 2, 29000 + (os.getpid() % 1000)
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in range(world)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for rank, ok_avg, ok_same, ok_zero, f_all, f_one in res:
    assert ok_avg and ok_same and ok_zero, (rank, ok_avg, ok_same, ok_zero)
    assert f_all is True
    assert f_one is False                         # NaN on rank 1 => both ranks skip


def _train_worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  torch.set_num_threads(2)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    import contextlib
    import io
    from conftest import load_golden
    from cpu_shim import cpu_ops
    from sg2im_b200.model import Sg2ImModel
    from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator
    from sg2im_b200.synth import synth_batch
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
    if rank == 1:
      # a replica that starts from DIFFERENT weights and running statistics (a resumed / differently
      # seeded rank): TrainStep must bring it to rank 0's state before the first step
      with torch.no_grad():
        for net in (m, d_img, d_obj):
          for p in net.parameters():
            p.add_(0.25)
          for b in net.buffers():
            if b.dtype.is_floating_point:
              b.add_(1.0)
    H, W = kw['image_size']
    N = g['batch'][0].size(0)
    # every rank its own shard (same shapes, different content), as bench.py does under torchrun
    shard = synth_batch(N=N, objs_per_img=3, rels_per_img=2, image_size=(H, W), num_objs=9,
                        num_preds=5, seed=1000 * rank)
    with cpu_ops():
      step = TrainStep(m, d_obj, d_img, fused_adam='flat' if os.environ.get('SG2IM_TEST_FLAT') else None)
      synced = all(torch.equal(v, g['sd_g'][k]) for k, v in m.state_dict().items())
      assert synced, 'rank %d did not start from rank 0 weights / buffers' % rank
      out = []
      for it in range(2):
        torch.manual_seed(50 + 10 * it + rank)
        noise = torch.randn(N, kw['layout_noise_dim'], H, W)
        losses, _ = step.step(shard, noise=noise)
        out.append(losses['total_loss'])
    flat = torch.cat([p.detach().reshape(-1) for net in (m, d_obj, d_img) for p in net.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    k0 = next(k for k, v in g['sd_g'].items() if v.dtype.is_floating_point and 'running' not in k)
    moved = float((m.state_dict()[k0] - g['sd_g'][k0]).abs().max())
    q.put((rank, all(torch.equal(gathered[0], x) for x in gathered), out, moved))
  finally:
    dist.destroy_process_group()


def test_two_rank_training_iterations_keep_replicas_identical():
  """Two gloo ranks run two full G + D iterations (kernels swapped for their CPU
  definition) on DIFFERENT shards: the flat-bucket all-reduce must leave every
  parameter bit-identical across ranks while the rank-local losses differ
  (rank-local BatchNorm statistics)."""
  world, port = 2, 30000 + (os.getpid() % 1000)
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  procs = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = sorted(q.get(timeout=300) for _ in range(world))
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  assert all(r[1] for r in res)
  assert res[0][2] != res[1][2]                     # different shards, different losses
  assert all(r[3] > 0 for r in res)                 # and the optimisers did step
