"""Tile-shape sweep of the tensor-core convolution over the shapes of one
training iteration (round-2 tuning aid; not part of the library).

For every distinct forward / data-gradient shape recorded in
profiles/r01_tf32_conv_shapes_v2.json, time sg2im_conv_tc with the N tile pinned
to 64 / 128 / 256 (SG2IM_TC_BN), with the halo kernel on / off (SG2IM_NO_HALO) and with the
cluster-multicast per-tap kernel (SG2IM_CONV_MC), next to the built-in heuristic.  CUDA events, 3 warm-up + 10 timed launches,
input larger than L2 or L2 flushed between launches is NOT attempted here: this
is a relative comparison of variants on identical inputs.

  python tools/sweep_conv.py [--out gpurun_out/sweep_conv.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from sg2im_b200 import ops  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def shapes():
  rows = json.load(open(os.path.join(HERE, '..', 'profiles', 'r01_tf32_conv_shapes_v2.json')))
  seen, out = set(), []
  for r in rows:
    if r['kernel'] not in ('conv_fwd_tc', 'conv_dgrad_tc'):
      continue
    key = tuple(r['shape(N,H,W,Cin,Cout,K,S)'])
    if key in seen or key[5] not in (1, 2, 3):
      continue
    seen.add(key)
    out.append((key, r['ms_per_step'] / r['launches_per_step']))
  out.sort(key=lambda kv: -kv[1])
  return out


def time_one(x, w, K, P, Co, out_hw, reps=10):
  for _ in range(3):
    ops.conv_tc(x, w, None, K, K, P, Co, out_hw=out_hw)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    ops.conv_tc(x, w, None, K, K, P, Co, out_hw=out_hw)
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps * 1e3            # us


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--out', default='gpurun_out/sweep_conv.json')
  ap.add_argument('--top', type=int, default=40)
  args = ap.parse_args()
  dev = torch.device('cuda:0')
  ops.set_conv_math('tf32')
  # (name, SG2IM_TC_BN, SG2IM_NO_HALO, SG2IM_CONV_MC); 'small' = SG2IM_HALO_SMALL (8-row maps only), 'pair' = SG2IM_HALO_PAIR (halo shapes only)
  variants = [('auto', None, False, False), ('bn64', '64', False, False), ('bn128', '128', False, False),
              ('bn256', '256', False, False), ('nohalo', None, True, False), ('nohalo128', '128', True, False),
              ('mc', None, True, True), ('mc64', '64', True, True), ('mc128', '128', True, True),
              ('mc256', '256', True, True), ('small', 'small', False, False), ('pair', 'pair', False, False)]
  results = []
  for (N, H, W, Ci, Co, K, S), ref_ms in shapes()[:args.top]:
    torch.manual_seed(0)
    x = torch.randn(N, H, W, Ci, device=dev)
    w = ops.pack_tc_fwd(torch.randn(Co, Ci, K, K, device=dev) * 0.05)
    P = (K - 1) // 2 if K == 3 else 0
    out_hw = (H + 2 * P - K + 1, W + 2 * P - K + 1)
    flops = 2.0 * N * out_hw[0] * out_hw[1] * Ci * Co * K * K
    row = {'shape': [N, H, W, Ci, Co, K], 'r01_us': ref_ms * 1e3}
    for name, bn, nohalo, mc in variants:
      for k in ('SG2IM_TC_BN', 'SG2IM_NO_HALO', 'SG2IM_CONV_MC'):
        os.environ.pop(k, None)
      os.environ.pop('SG2IM_HALO_SMALL', None)
      os.environ.pop('SG2IM_HALO_PAIR', None)
      if bn == 'small':
        os.environ['SG2IM_HALO_SMALL'] = '1'
      elif bn == 'pair':
        os.environ['SG2IM_HALO_PAIR'] = '1'
      elif bn:
        os.environ['SG2IM_TC_BN'] = bn
      if nohalo:
        os.environ['SG2IM_NO_HALO'] = '1'
      if mc:
        os.environ['SG2IM_CONV_MC'] = '1'
      us = time_one(x, w, K, P, Co, out_hw)
      row[name] = {'us': us, 'tflops': flops / us / 1e6}
    for k in ('SG2IM_TC_BN', 'SG2IM_NO_HALO', 'SG2IM_CONV_MC', 'SG2IM_HALO_SMALL', 'SG2IM_HALO_PAIR'):
      os.environ.pop(k, None)
    best = min(variants, key=lambda v: row[v[0]]['us'])[0]
    row['best'] = best
    results.append(row)
    print('%-34s auto %7.1f us %6.1f TF/s | %s | best %s' % (
        row['shape'], row['auto']['us'], row['auto']['tflops'],
        ' '.join('%s %.0f' % (v[0], row[v[0]]['us']) for v in variants[1:]), best), flush=True)
  os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
  json.dump(results, open(args.out, 'w'), indent=1)
  gain = sum(r['auto']['us'] - r[r['best']]['us'] for r in results)
  print('sum over shapes of (auto - best): %.1f us' % gain)


if __name__ == '__main__':
  main()
