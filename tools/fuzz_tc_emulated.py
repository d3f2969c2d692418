"""Random-shape fuzzing of the tensor-core kernels under the functional model (no GPU):
forward / data gradient (packed weights and the in-place weight-gradient layout, all tile
heuristics) and the weight gradient, in a random arithmetic per case (tf32 / bf16x3 / bf16), under
random adversarial schedules of the asynchronous model (late loads, lagging tensor pipe, lagging
epilogue, few SMs so that pipeline slots and barrier phases wrap), each against torch in fp64 on
operands that are exact in that arithmetic (bf16x3: arbitrary fp32, tolerance 3e-5).
Usage: python tools/fuzz_tc_emulated.py [cases] [seed]
"""
import ctypes
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from emul_device import build_lib  # noqa: E402
from sg2im_b200._lib import SIGNATURES  # noqa: E402

KEYS = ('SG2IM_NO_HALO', 'SG2IM_TC_BN', 'SG2IM_EMUL_SMS', 'SG2IM_EMUL_ASYNC_SLOW3D', 'SG2IM_EMUL_SLOW_EPILOGUE',
        'SG2IM_EMUL_SLOW_PIPE')
MATH = 0


def tf32(t):
  """Operands exact in the case's arithmetic (MATH: 0 tf32, 1 bf16x3 = any fp32, 2 bf16)."""
  if MATH == 1:
    return t
  return (t.view(torch.int32) & (~0x1fff if MATH == 0 else ~0xffff)).view(torch.float32)


def p(t):
  return None if t is None else t.data_ptr()


def rel(a, b):
  return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))


def main():
  cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
  rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
  L = ctypes.CDLL(build_lib())
  for name, sig in SIGNATURES.items():
    if hasattr(L, name):
      getattr(L, name).argtypes = sig
  L.emul_last_error.restype = ctypes.c_char_p
  worst = 0.0
  for case in range(cases):
    K = rng.choice([1, 2, 3, 3, 3])
    P = rng.choice([0, (K - 1) // 2]) if K == 3 else 0
    N = rng.randint(1, 5)
    H, W = rng.randint(max(K, 1), 20), rng.randint(max(K, 1), 20)
    Ci = 4 * rng.randint(1, 40)
    Co = 4 * rng.randint(1, 70)
    Cf = Ci + 4 * rng.randint(0, 3)                       # weights may be wider than the channels used
    env = {}
    if rng.random() < 0.3: env['SG2IM_NO_HALO'] = '1'
    if rng.random() < 0.4: env['SG2IM_TC_BN'] = rng.choice(['64', '128', '256'])
    if rng.random() < 0.5 and K > 1:                       # halo shapes: >= 16 rows, >= 8 columns
      H, W = rng.randint(16, 40) + K - 1 - 2 * P, rng.randint(8, 20) + K - 1 - 2 * P
    global MATH
    MATH = rng.choice([0, 1, 1, 1, 2])
    if rng.random() < 0.5:                                 # few SMs: persistent loops iterate; adversarial schedules
      env['SG2IM_EMUL_SMS'] = rng.choice(['2', '4', '8'])
      r = rng.random()
      if r < 0.25: env['SG2IM_EMUL_ASYNC_SLOW3D'] = '60'
      elif r < 0.5: env['SG2IM_EMUL_SLOW_EPILOGUE'] = '40'
      elif r < 0.75: env['SG2IM_EMUL_SLOW_PIPE'] = '25'
    for k in KEYS:
      os.environ.pop(k, None)
    os.environ.update(env)
    g = torch.Generator().manual_seed(case)
    T = K * K
    Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
    if Ho < 1 or Wo < 1:
      continue
    x = tf32(torch.randn(N, H, W, Ci, generator=g))
    wf = tf32(torch.randn(Co, Cf, K, K, generator=g) * 0.1)
    w = wf[:, :Ci]
    b = torch.randn(Co, generator=g)
    xr = x.permute(0, 3, 1, 2).double().clone().requires_grad_(True)
    wr = w.double().clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, b.double(), padding=P)
    gy = tf32(torch.randn(N, Ho, Wo, Co, generator=g))
    ref.backward(gy.permute(0, 3, 1, 2).double())
    errs = {}
    # forward, packed weights
    y = torch.empty(N, Ho, Wo, Co)
    wt = w.permute(2, 3, 0, 1).reshape(T, Co, Ci).contiguous()
    assert L.sg2im_conv_tc(p(x), Ci, N, H, W, Ci, p(wt), p(b), K, K, P, Ho, Wo, Co, 0, 0.0, p(y), Co, 0,
                           None, 0, MATH, None) == 0, L.emul_last_error()
    errs['fwd'] = rel(y, ref.permute(0, 2, 3, 1))
    # forward + dgrad from the weight-gradient layout
    kcc = wf.permute(2, 3, 1, 0).reshape(T, Cf, Co).contiguous()
    y2 = torch.empty(N, Ho, Wo, Co)
    assert L.sg2im_conv_tc_kcc(p(x), Ci, N, H, W, Ci, p(kcc), Cf, 0, p(b), K, K, P, Ho, Wo, Co, 0, 0.0,
                               p(y2), Co, 0, None, 0, MATH, None) == 0, L.emul_last_error()
    errs['fwd_kcc'] = rel(y2, ref.permute(0, 2, 3, 1))
    if K - 1 - P >= 0:
      dx = torch.empty(N, H, W, Ci)
      assert L.sg2im_conv_tc_kcc(p(gy), Co, N, Ho, Wo, Co, p(kcc), Cf, 1, None, K, K, K - 1 - P, H, W, Ci,
                                 0, 0.0, p(dx), Ci, 0, None, 0, MATH, None) == 0, L.emul_last_error()
      errs['dgrad_kcc'] = rel(dx, xr.grad.permute(0, 2, 3, 1))
    # weight gradient
    if L.sg2im_conv_wgrad_tc_supported(N, H, W, Ci, Ci, K, K, 1, P, Ho, Wo, Co):
      dw = torch.zeros(T * Ci, Co)
      assert L.sg2im_conv_wgrad_tc(p(x), Ci, N, H, W, Ci, p(gy), K, K, P, Ho, Wo, Co, p(dw), MATH, 0, None) == 0, \
          L.emul_last_error()
      errs['wgrad'] = rel(dw, wr.grad.permute(2, 3, 1, 0).reshape(T * Ci, Co))
    bad = {k: v for k, v in errs.items() if v > (3e-5 if MATH == 1 else 5e-6)}
    worst = max([worst] + list(errs.values()))
    print('%3d %s math=%d env=%s  %s%s' % (case, (N, H, W, Ci, Co, K, P, Cf), MATH, env,
                                   ' '.join('%s=%.1e' % kv for kv in errs.items()),
                                   '   <<<<< MISMATCH' if bad else ''), flush=True)
    assert not bad, bad
  for k in KEYS:
    os.environ.pop(k, None)
  print('all cases agree; worst relative error %.2e' % worst)


if __name__ == '__main__':
  main()
