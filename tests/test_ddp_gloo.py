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
