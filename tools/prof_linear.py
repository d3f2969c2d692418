"""Where do the ~16 us of the small per-tap launches (graph-convolution Linears, mask head, discriminator
tails: 62 launches of conv_tc_kernel<64,...> per training step) go?  Replays CUDA graphs of 40
back-to-back launches of one shape and prints the time per launch for a sweep over the reduction
length (fixed cost vs cost per 32-channel k-block), the row count (one tile vs several per CTA) and
the arithmetic (tf32 = no converter warps in the chain).  Not part of the library.
Usage (GPU): python tools/prof_linear.py > gpurun_out/prof_linear.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sg2im_b200 import ops

REPS = 40


def per_launch_us(fn):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  g = torch.cuda.CUDAGraph()
  s = torch.cuda.Stream()
  s.wait_stream(torch.cuda.current_stream())
  with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
      for _ in range(REPS):
        fn()
  torch.cuda.synchronize()
  best = 1e9
  for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e3 / REPS)
  return best


def linear_case(M, Ci, Co, math, H=1, W=1, K=1, P=0):
  dev = torch.device('cuda:0')
  ops.set_conv_math(math)
  x = torch.randn(M, H, W, Ci, device=dev)
  w = torch.randn(Co, Ci, K, K, device=dev) * 0.05
  b = torch.randn(Co, device=dev)
  if math == 'tf32':
    wp = ops.pack_tc_fwd(w)
    return per_launch_us(lambda: ops.conv_tc(x, wp, b, K, K, P, Co, act=1, slope=0.0))
  wk = w.permute(2, 3, 1, 0).contiguous().permute(3, 2, 0, 1)
  sh = ops.SplitShadows([wk])
  sh.refresh()
  sf = wk._split_fwd
  return per_launch_us(lambda: ops.conv_tc_presplit(x, sf, Co, b, K, K, P, Co, 1, 0.0))


def main():
  dev = torch.device('cuda:0')
  z = torch.zeros(1 << 20, device=dev)
  print('floor: %.2f us per launch (40 x fill of 4 MB in a graph)' % per_launch_us(lambda: z.zero_()))
  print('%-34s %9s %9s %9s' % ('shape (rows, Cin -> Cout)', 'tf32', 'bf16x3', 'bf16'))
  cases = [('K sweep', [(448, ci, 512) for ci in (32, 128, 384, 1024)]),
           ('row sweep', [(m, 384, 512) for m in (128, 448, 2048, 8192)]),
           ('Cout sweep', [(448, 512, co) for co in (64, 128, 512, 1152)]),
           ('step shapes', [(448, 384, 512), (448, 512, 1152), (320, 512, 512), (320, 512, 128), (320, 128, 512),
                            (320, 1024, 1024)])]
  for title, shapes in cases:
    print('-- ' + title)
    for (M, Ci, Co) in shapes:
      t = [linear_case(M, Ci, Co, m) for m in ('tf32', 'bf16x3', 'bf16')]
      print('%-34s %9.2f %9.2f %9.2f' % ('%d, %d -> %d' % (M, Ci, Co), t[0], t[1], t[2]))
  print('-- small maps (N images, HxW, 3x3)')
  for (N, HW, Ci, Co) in [(320, 2, 128, 128), (320, 4, 128, 128), (320, 8, 128, 128), (320, 16, 128, 128),
                          (32, 8, 1024, 1024)]:
    t = [linear_case(N, Ci, Co, m, HW, HW, 3, 1) for m in ('tf32', 'bf16x3', 'bf16')]
    print('%-34s %9.2f %9.2f %9.2f' % ('%d x %dx%d, %d -> %d' % (N, HW, HW, Ci, Co), t[0], t[1], t[2]))


if __name__ == '__main__':
  main()
