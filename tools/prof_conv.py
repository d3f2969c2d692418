"""Stand-alone launcher of the dominant convolution shapes for ncu captures
(profiles/).  Not part of the library.

  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 2 \
      -o gpurun_out/prof_conv_tc python tools/prof_conv.py fwd [big|mid|small|n64] [bf16x3|tf32|bf16]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sg2im_b200 import ops

SHAPES = {
    # name: (N, H, W, Cin, Cout, K, P)   CRN stage-4 conv1 at VG-128, batch 32
    'big': (32, 128, 128, 288, 64, 3, 1),
    'mid': (32, 32, 32, 672, 256, 3, 1),
    'small': (32, 8, 8, 1024, 1024, 3, 1),      # CRN stage-0 conv2: 8x8 maps, widest channels
    'n64': (32, 128, 128, 64, 64, 3, 1),        # the two other 128x128 convolutions
}


def main():
  what = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
  shape = SHAPES[sys.argv[2] if len(sys.argv) > 2 else 'big']
  math = sys.argv[3] if len(sys.argv) > 3 else 'bf16x3'
  N, H, W, Ci, Co, K, P = shape
  dev = torch.device('cuda:0')
  ops.set_conv_math(math)
  torch.manual_seed(0)
  x = torch.randn(N, H, W, Ci, device=dev)
  w = torch.randn(Co, Ci, K, K, device=dev) * 0.05
  b = torch.randn(Co, device=dev)
  dy = torch.randn(N, H, W, Co, device=dev)
  flops = 2.0 * N * H * W * Ci * Co * K * K
  reps = 3
  for _ in range(reps):
    if what == 'fwd':
      ops.conv_tc(x, ops.pack_tc_fwd(w), b, K, K, P, Co)
    elif what == 'dgrad':
      ops.conv_tc(dy, ops.pack_tc_dgrad(w), None, K, K, K - 1 - P, Ci)
    elif what == 'wgrad':
      ops.conv_wgrad(x, dy, K, K, 1, P)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  wf, wd = ops.pack_tc_fwd(w), ops.pack_tc_dgrad(w)
  presplit = math in ('bf16x3', 'bf16') and os.environ.get('SG2IM_PROF_PRESPLIT', '1') != '0'
  if presplit:
    # the training step's configuration: weights in the weight-gradient layout, split once per step
    wk = w.permute(2, 3, 1, 0).contiguous().permute(3, 2, 0, 1)
    sh = ops.SplitShadows([wk])
    sh.refresh()
    sf, sd = wk._split_fwd, wk._split_dgrad
    for _ in range(3):
      if what == 'fwd':
        ops.conv_tc_presplit(x, sf, Co, b, K, K, P, Co)
      elif what == 'dgrad':
        ops.conv_tc_presplit(dy, sd, Ci, None, K, K, K - 1 - P, Ci)
    torch.cuda.synchronize()
  e0.record()
  for _ in range(10):
    if what == 'fwd' and presplit:
      ops.conv_tc_presplit(x, sf, Co, b, K, K, P, Co)
    elif what == 'dgrad' and presplit:
      ops.conv_tc_presplit(dy, sd, Ci, None, K, K, K - 1 - P, Ci)
    elif what == 'fwd':
      ops.conv_tc(x, wf, b, K, K, P, Co)
    elif what == 'dgrad':
      ops.conv_tc(dy, wd, None, K, K, K - 1 - P, Ci)
    else:
      ops.conv_wgrad(x, dy, K, K, 1, P)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 10
  print('%s %s %s: %.1f us, %.1f TFLOP/s (algorithmic)' % (what, math, shape, ms * 1e3, flops / ms / 1e9))


if __name__ == '__main__':
  main()
