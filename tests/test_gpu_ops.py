"""GPU parity tests, op by op: the CUDA path (through the ctypes C-ABI) against
the CPU oracle / torch-CPU primitives on the same seeded inputs.

Tolerances (relative = max|a-b| / max|b|):
  * index / scatter work: bit-exact (torch.equal)
  * exact-fp32 kernels (FFMA convs, BN, layout, crop): 1e-4 (accumulation-order
    differences only; north_star allows 1e-3)
"""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err, load_golden

pytestmark = pytest.mark.gpu

TOL = 1e-4


def dev():
  return torch.device('cuda:0')


def test_library_loads_and_device_ok():
  from sg2im_b200 import _lib
  lib = _lib.load()
  assert lib.sg2im_abi_version() >= 1
  assert lib.sg2im_device_ok() == 1


def test_cpu_tensor_is_refused():
  from sg2im_b200 import ops
  with pytest.raises(RuntimeError):
    ops.conv2d(torch.zeros(1, 4, 4, 4), torch.zeros(4, 4, 3, 3), None, 1, 1)


@pytest.mark.parametrize('T,O,H,D', [(10, 7, 16, 8), (448, 320, 512, 128), (4096, 2112, 512, 128),
                                     (33, 5, 6, 3)])
@pytest.mark.parametrize('avg', [True, False])
def test_graph_pool_bit_exact(T, O, H, D, avg):
  from sg2im_b200 import ops
  from oracle import sg2im_oracle as orc
  g = torch.Generator().manual_seed(T + O)
  new_t = torch.randn(T, 2 * H + D, generator=g)
  edges = torch.randint(0, max(O - 2, 1), (T, 2), generator=g)      # last rows unused
  want = orc.graph_pool(new_t, edges, O, H, D, 'avg' if avg else 'sum')
  e = edges.to(dev())
  row_ptr, entries = ops.csr_build(e, 2, O)
  got = ops.segment_sum(new_t.to(dev()), 0, H + D, H, row_ptr, entries, O, avg)
  assert torch.equal(got.cpu(), want)
  # CSR invariants
  rp = row_ptr.cpu()
  assert rp[0] == 0 and rp[-1] == 2 * T
  ent = entries.cpu()[:2 * T]
  for r in (0, O // 2):
    seg = ent[rp[r]:rp[r + 1]]
    keys = (seg & 1) * (T + 1) + (seg >> 1)
    assert torch.equal(keys, keys.sort().values)


def test_graph_pool_empty_and_single():
  from sg2im_b200 import ops
  # no triples at all: every pooled row is zero (graph.py:92, clamp(min=1))
  e = torch.zeros(0, 2, dtype=torch.int64, device=dev())
  row_ptr, entries = ops.csr_build(e, 2, 3)
  assert row_ptr.cpu().tolist() == [0, 0, 0, 0]


def test_gconv_layer_forward_backward():
  from sg2im_b200.graph import GraphTripleConv
  from oracle import sg2im_oracle as orc
  g = load_golden('gconv.pt')
  layer = GraphTripleConv(input_dim=8, output_dim=8, hidden_dim=16, pooling='avg')
  layer.load_state_dict(g['sd'])
  layer = layer.to(dev())
  ov = g['obj_vecs'].to(dev()).requires_grad_(True)
  pv = g['pred_vecs'].to(dev()).requires_grad_(True)
  new_obj, new_p = layer(ov, pv, g['edges'].to(dev()))
  assert rel_err(new_obj, g['new_obj']) < TOL
  assert rel_err(new_p, g['new_p']) < TOL
  # gradients vs oracle autograd on CPU
  sd = {'L.' + k: v.clone().requires_grad_(True) for k, v in g['sd'].items()}
  ov_c = g['obj_vecs'].detach().clone().requires_grad_(True)
  pv_c = g['pred_vecs'].detach().clone().requires_grad_(True)
  ro, rp = orc.graph_triple_conv(sd, 'L', ov_c, pv_c, g['edges'])
  wo, wp = torch.randn_like(ro), torch.randn_like(rp)
  ((ro * wo).sum() + (rp * wp).sum()).backward()
  ((new_obj * wo.to(dev())).sum() + (new_p * wp.to(dev())).sum()).backward()
  assert rel_err(ov.grad, ov_c.grad) < TOL
  assert rel_err(pv.grad, pv_c.grad) < TOL
  for k, p in layer.named_parameters():
    assert rel_err(p.grad, sd['L.' + k].grad) < TOL, k


CONV_CASES = [
    # N, H, W, Cin, Cout, K, S, P
    (2, 8, 8, 16, 32, 3, 1, 1),
    (2, 9, 7, 5, 3, 3, 1, 1),          # odd everything, skinny output
    (3, 16, 16, 3, 64, 4, 2, 0),       # discriminator first layer
    (2, 15, 15, 64, 128, 4, 2, 0),     # discriminator second layer (odd input)
    (2, 6, 6, 128, 256, 4, 2, 0),
    (4, 16, 16, 64, 3, 1, 1, 0),       # RGB head
    (2, 16, 16, 128, 1, 1, 1, 0),      # mask head
    (2, 8, 8, 161, 64, 3, 1, 1),       # CRN stage-0 width (not a multiple of 4)
    (1, 32, 32, 288, 64, 3, 1, 1),
    (1, 4, 4, 200, 260, 3, 1, 1),      # > one N tile, ragged
    (320, 1, 1, 1024, 1, 1, 1, 0),     # discriminator real / fake head: warp-per-row skinny kernel
    (37, 1, 1, 260, 3, 1, 1, 0),
]


@pytest.mark.parametrize('N,H,W,Ci,Co,K,S,P', CONV_CASES)
def test_conv_forward_dgrad_wgrad(N, H, W, Ci, Co, K, S, P):
  from sg2im_b200 import ops
  g = torch.Generator().manual_seed(Ci * 131 + Co)
  x = torch.randn(N, Ci, H, W, generator=g)
  w = torch.randn(Co, Ci, K, K, generator=g) * 0.1
  b = torch.randn(Co, generator=g)
  xr = x.clone().requires_grad_(True)
  wr = w.clone().requires_grad_(True)
  br = b.clone().requires_grad_(True)
  yr = F.leaky_relu(F.conv2d(xr, wr, br, stride=S, padding=P), 0.2)
  gy = torch.randn(yr.shape, generator=g)
  yr.backward(gy)

  xd = x.to(dev()).permute(0, 2, 3, 1).requires_grad_(True)         # NCHW memory, NHWC view
  wd = w.to(dev()).requires_grad_(True)
  bd = b.to(dev()).requires_grad_(True)
  y = ops.conv2d(xd, wd, bd, S, P, 1, 0.2)
  assert rel_err(y.permute(0, 3, 1, 2), yr) < TOL
  y.backward(gy.to(dev()).permute(0, 2, 3, 1))
  assert rel_err(xd.grad.permute(0, 3, 1, 2), xr.grad) < TOL
  assert rel_err(wd.grad, wr.grad) < TOL
  assert rel_err(bd.grad, br.grad) < TOL


@pytest.mark.parametrize('M,K,Nn', [(448, 384, 512), (320, 512, 4), (7, 24, 46), (1, 256, 1024),
                                    (320, 1024, 1), (37, 260, 3)])    # heads: warp-per-row skinny kernel
def test_linear_relu(M, K, Nn):
  from sg2im_b200 import ops
  g = torch.Generator().manual_seed(M + K)
  x, w, b = torch.randn(M, K, generator=g), torch.randn(Nn, K, generator=g) * 0.1, torch.randn(Nn, generator=g)
  xr, wr, br = [t.clone().requires_grad_(True) for t in (x, w, b)]
  yr = torch.relu(F.linear(xr, wr, br))
  gy = torch.randn(yr.shape, generator=g)
  yr.backward(gy)
  xd, wd, bd = [t.to(dev()).requires_grad_(True) for t in (x, w, b)]
  y = ops.linear(xd, wd, bd, 1, 0.0)
  assert rel_err(y, yr) < TOL
  y.backward(gy.to(dev()))
  assert rel_err(xd.grad, xr.grad) < TOL
  assert rel_err(wd.grad, wr.grad) < TOL
  assert rel_err(bd.grad, br.grad) < TOL


@pytest.mark.parametrize('training', [True, False])
@pytest.mark.parametrize('up,slope,C', [(1, 0.2, 32), (2, 0.2, 64), (2, 1.0, 16), (1, 0.0, 7)])
def test_bn_act_upsample_slice(training, up, slope, C):
  """BatchNorm2d -> LeakyReLU -> nearest upsample, written into a channel
  slice, forward + backward, vs the torch CPU composition (crn.py:43-47,107)."""
  from sg2im_b200 import ops
  from sg2im_b200.layers import BatchNorm2d
  g = torch.Generator().manual_seed(C + up)
  N, H, W, off, extra = 3, 6, 5, 8, 4
  x = torch.randn(N, C, H, W, generator=g) * 2 + 0.5
  bn_ref = torch.nn.BatchNorm2d(C)
  with torch.no_grad():
    bn_ref.weight.copy_(torch.rand(C, generator=g) + 0.5)
    bn_ref.bias.copy_(torch.randn(C, generator=g))
    bn_ref.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
    bn_ref.running_var.copy_(torch.rand(C, generator=g) + 0.5)
  bn = BatchNorm2d(C)
  bn.load_state_dict(bn_ref.state_dict())
  bn = bn.to(dev())
  bn_ref.train(training)
  bn.train(training)
  xr = x.clone().requires_grad_(True)
  yr = F.leaky_relu(bn_ref(xr), slope)
  if up > 1:
    yr = F.interpolate(yr, scale_factor=up, mode='nearest')
  gy = torch.randn(yr.shape, generator=g)
  yr.backward(gy)

  xd = x.to(dev()).permute(0, 2, 3, 1).contiguous().requires_grad_(True)
  buf = torch.zeros(N, H * up, W * up, off + C + extra, device=dev())
  base = buf.clone().requires_grad_(True)
  work = base * 1.0
  out = ops.bn_act(xd, bn, slope, up=up, unbias_mult=1, out=work, out_coff=off)
  got = out[..., off:off + C].permute(0, 3, 1, 2)
  assert rel_err(got, yr) < TOL
  assert float(out[..., :off].abs().max()) == 0 and float(out[..., off + C:].abs().max()) == 0
  gfull = torch.zeros_like(buf)
  gfull[..., off:off + C] = gy.to(dev()).permute(0, 2, 3, 1)
  out.backward(gfull)
  assert rel_err(xd.grad.permute(0, 3, 1, 2), xr.grad) < 5 * TOL
  assert rel_err(bn.weight.grad, bn_ref.weight.grad) < 5 * TOL
  assert rel_err(bn.bias.grad, bn_ref.bias.grad) < 5 * TOL
  if training:
    assert rel_err(bn.running_mean, bn_ref.running_mean) < TOL
    assert rel_err(bn.running_var, bn_ref.running_var) < TOL
    assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked)


def test_upsample_then_bn_running_var():
  """mask_net order Upsample -> BN (model.py:98-99): statistics of the
  upsampled tensor, unbiased factor from the post-upsample count."""
  from sg2im_b200.layers import FusedSequential, Upsample, BatchNorm2d
  torch.manual_seed(0)
  x = torch.randn(5, 8, 2, 2)
  ref = torch.nn.Sequential(torch.nn.Upsample(scale_factor=2, mode='nearest'), torch.nn.BatchNorm2d(8))
  mine = FusedSequential(Upsample(scale_factor=2, mode='nearest'), BatchNorm2d(8)).to(dev())
  yr = ref(x)
  y = mine(x.to(dev()))
  assert rel_err(y, yr) < TOL
  assert rel_err(mine[1].running_var, ref[1].running_var) < TOL
  assert rel_err(mine[1].running_mean, ref[1].running_mean) < TOL


def test_layout_golden_demo_and_random():
  from sg2im_b200.layout import masks_to_layout, boxes_to_layout
  g = load_golden('layout.pt')
  d = dev()
  out_b = boxes_to_layout(g['vecs'].to(d), g['boxes'].to(d), g['obj_to_img'].to(d), 32)
  out_m = masks_to_layout(g['vecs'].to(d), g['boxes'].to(d), g['masks'].to(d),
                          g['obj_to_img'].to(d), 32)
  out_r = masks_to_layout(g['rvecs'].to(d), g['rboxes'].to(d), g['rmasks'].to(d),
                          g['robj_to_img'].to(d), 24, 40)
  assert rel_err(out_b, g['out_boxes']) < TOL
  assert rel_err(out_m, g['out_masks']) < TOL
  assert rel_err(out_r, g['out_rand']) < TOL


@pytest.mark.parametrize('O,N,D,M,H,W', [(12, 3, 16, 8, 24, 20), (40, 4, 128, 16, 64, 64)])
def test_layout_backward(O, N, D, M, H, W):
  from sg2im_b200.layout import masks_to_layout
  from oracle import sg2im_oracle as orc
  g = torch.Generator().manual_seed(O)
  vecs = torch.randn(O, D, generator=g)
  xy = torch.rand(O, 2, generator=g) * 0.6
  boxes = torch.cat([xy, xy + torch.rand(O, 2, generator=g) * 0.35 + 0.1], 1)
  boxes[-1] = torch.tensor([0., 0., 1., 1.])
  masks = torch.rand(O, M, M, generator=g)
  o2i = torch.sort(torch.randint(0, N, (O,), generator=g)).values
  vr, mr = vecs.clone().requires_grad_(True), masks.clone().requires_grad_(True)
  ref = orc.masks_to_layout(vr, boxes, mr, o2i, H, W, N)
  gy = torch.randn(ref.shape, generator=g)
  ref.backward(gy)
  d = dev()
  vd, md = vecs.to(d).requires_grad_(True), masks.to(d).requires_grad_(True)
  out = masks_to_layout(vd, boxes.to(d), md, o2i.to(d), H, W, num_imgs=N)
  assert rel_err(out, ref) < TOL
  out.backward(gy.to(d))
  assert rel_err(vd.grad, vr.grad) < TOL
  assert rel_err(md.grad, mr.grad) < TOL


def test_crop_golden_and_backward():
  from sg2im_b200.bilinear import crop_bbox_batch
  from oracle import sg2im_oracle as orc
  g = load_golden('crop.pt')
  d = dev()
  fr = g['feats'].clone().requires_grad_(True)
  ref = orc.crop_bbox_batch(fr, g['boxes'], g['bbox_to_feats'], 8)
  assert rel_err(ref, g['crops']) < 1e-6
  gy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
  ref.backward(gy)
  fd = g['feats'].to(d).requires_grad_(True)
  out = crop_bbox_batch(fd, g['boxes'].to(d), g['bbox_to_feats'].to(d), 8)
  assert rel_err(out, g['crops']) < TOL
  out.backward(gy.to(d))
  assert rel_err(fd.grad, fr.grad) < TOL


# ---------------------------------------------------------------------------
# tcgen05 tensor-core convolution (TF32 multiply, fp32 accumulate)
# ---------------------------------------------------------------------------

def _tf32_exact(t):
  """Zero the 13 low mantissa bits: values the tensor core consumes exactly."""
  return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


TC_CASES = [
    # N, H, W, Cin, Cout, K, P
    (2, 16, 16, 32, 64, 3, 1),
    (3, 8, 8, 160, 256, 3, 1),         # 8x8: two images per tile, ragged image count
    (2, 32, 32, 96, 128, 3, 1),
    (1, 64, 64, 64, 64, 3, 1),
    (5, 4, 4, 128, 128, 3, 1),         # mask head sizes
    (7, 2, 2, 128, 128, 3, 1),
    (2, 16, 16, 288, 512, 3, 1),       # two N tiles
    (448, 1, 1, 384, 512, 1, 0),       # Linear as a 1x1 convolution over rows
    (70, 1, 1, 512, 1152, 1, 0),
    (2, 16, 16, 64, 64, 1, 0),
    (2, 32, 16, 36, 64, 3, 1),         # Cin not a multiple of 32 (TMA zero fill), H != W
    (2, 16, 16, 288, 64, 3, 1),        # dgrad output width 288 = 256 + 32: partial last N tile
    (1, 32, 32, 160, 96, 3, 1),        # Cout 96: partial N tile in the forward
    (4, 64, 64, 64, 64, 3, 1),         # wgrad: several pixel splits, 5+4 tap passes
    (1, 128, 128, 288, 64, 3, 1),      # CRN stage-4 conv1 shape (one image)
    (2, 16, 16, 512, 256, 3, 1),       # wgrad: 4 ci tiles, 2 taps per pass
]


@pytest.mark.parametrize('N,H,W,Ci,Co,K,P', TC_CASES)
def test_conv_tc_forward_dgrad(N, H, W, Ci, Co, K, P):
  """With TF32-exact operands every product is exact in fp32, so the tensor
  core result must agree with the fp32 reference to accumulation-order level
  (2e-5); with arbitrary fp32 operands the TF32 operand truncation (2^-10
  relative per operand, biased toward zero) bounds the error: 3e-3 stated.
  The backward check uses a linear conv (no activation): with an activation in
  between, y ~ 0 elements may legitimately take the other LeakyReLU branch."""
  from sg2im_b200 import ops
  g = torch.Generator().manual_seed(Ci + 7 * Co + H)
  x = torch.randn(N, Ci, H, W, generator=g)
  w = torch.randn(Co, Ci, K, K, generator=g) * 0.1
  b = torch.randn(Co, generator=g)
  gy = torch.randn(N, Co, H + 2 * P - K + 1, W + 2 * P - K + 1, generator=g)
  ops.set_conv_math('tf32')
  try:
    for exact in (True, False):
      xx, ww, gg = (_tf32_exact(x), _tf32_exact(w), _tf32_exact(gy)) if exact else (x, w, gy)
      tol = 2e-5 if exact else 3e-3
      xd = xx.to(dev()).permute(0, 2, 3, 1).contiguous().requires_grad_(True)
      wd = ww.to(dev()).requires_grad_(True)
      bd = b.to(dev()).requires_grad_(True)
      assert ops.conv_tc_ok(xd, K, K, 1, P, Co), 'shape should take the tensor-core path'
      # forward with fused bias + LeakyReLU epilogue
      y_act = ops.conv2d(xd, wd, bd, 1, P, 1, 0.2)
      yr_act = F.leaky_relu(F.conv2d(xx, ww, b, padding=P), 0.2)
      assert rel_err(y_act.permute(0, 3, 1, 2), yr_act) < tol, ('fwd', exact)
      # linear conv: forward + dgrad (tensor core) + wgrad + bias grad
      xr, wr, br = xx.clone().requires_grad_(True), ww.clone().requires_grad_(True), b.clone().requires_grad_(True)
      yr = F.conv2d(xr, wr, br, padding=P)
      yr.backward(gg)
      y = ops.conv2d(xd, wd, bd, 1, P)
      assert rel_err(y.permute(0, 3, 1, 2), yr) < tol, ('fwd-linear', exact)
      y.backward(gg.to(dev()).permute(0, 2, 3, 1))
      assert rel_err(xd.grad.permute(0, 3, 1, 2), xr.grad) < tol, ('dgrad', exact)
      assert rel_err(wd.grad, wr.grad) < tol, ('wgrad', exact)
      assert rel_err(bd.grad, br.grad) < 1e-4
  finally:
    ops.set_conv_math('fp32')


def test_conv_tc_dgrad_exact_and_slice_output():
  """dgrad with TF32-exact dY (no activation in between) is exact to 1e-5; and
  the kernel writes into a channel slice of a wider buffer."""
  from sg2im_b200 import ops
  g = torch.Generator().manual_seed(3)
  N, H, W, Ci, Co, K, P = 2, 16, 16, 64, 128, 3, 1
  x = _tf32_exact(torch.randn(N, Ci, H, W, generator=g))
  w = _tf32_exact(torch.randn(Co, Ci, K, K, generator=g) * 0.1)
  gy = _tf32_exact(torch.randn(N, Co, H, W, generator=g))
  xr = x.clone().requires_grad_(True)
  yr = F.conv2d(xr, w, None, padding=P)
  yr.backward(gy)
  ops.set_conv_math('tf32')
  try:
    xd = x.to(dev()).permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    y = ops.conv2d(xd, w.to(dev()), None, 1, P)
    assert rel_err(y.permute(0, 3, 1, 2), yr) < 2e-5
    y.backward(gy.to(dev()).permute(0, 2, 3, 1))
    assert rel_err(xd.grad.permute(0, 3, 1, 2), xr.grad) < 2e-5
    buf = torch.zeros(N, H, W, 32 + Co + 8, device=dev())
    ops.conv_tc(xd.detach(), ops.pack_tc_fwd(w.to(dev())), None, K, K, P, Co, out=buf, out_coff=32)
    assert rel_err(buf[..., 32:32 + Co].permute(0, 3, 1, 2), yr) < 2e-5
    assert float(buf[..., :32].abs().max()) == 0 and float(buf[..., 32 + Co:].abs().max()) == 0
  finally:
    ops.set_conv_math('fp32')


S2_CASES = [
    # N, H, W, Cin, Cout  (4x4 stride-2 'valid' convs of the discriminators)
    (2, 32, 32, 3, 64),        # first layer on crops: 12 s2d channels
    (2, 15, 15, 64, 128),      # odd input size: zero-padded s2d, cropped output
    (3, 6, 6, 128, 256),
    (2, 63, 63, 64, 128),      # D_img layer 2 geometry
    (2, 30, 30, 128, 256),
    (5, 16, 20, 8, 32),        # H != W
]


@pytest.mark.parametrize('N,H,W,Ci,Co', S2_CASES)
def test_conv_stride2_space_to_depth_route(N, H, W, Ci, Co):
  """4x4/s2 conv == 2x2/s1 conv on the space-to-depth input, on the tensor-core
  kernels (fwd, dgrad incl. the 12-channel image gradient, wgrad).  TF32-exact
  operands -> 2e-5; also checks the NCHW-strided image input."""
  from sg2im_b200 import ops
  g = torch.Generator().manual_seed(H * 7 + Ci)
  x = _tf32_exact(torch.randn(N, Ci, H, W, generator=g))
  w = _tf32_exact(torch.randn(Co, Ci, 4, 4, generator=g) * 0.1)
  b = torch.randn(Co, generator=g)
  xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
  yr = F.conv2d(xr, wr, br, stride=2)
  gy = _tf32_exact(torch.randn(yr.shape, generator=g))
  yr.backward(gy)
  ops.set_conv_math('tf32')
  try:
    xd = x.to(dev()).requires_grad_(True)                 # NCHW memory, viewed NHWC
    wd, bd = w.to(dev()).requires_grad_(True), b.to(dev()).requires_grad_(True)
    y = ops.conv2d(xd.permute(0, 2, 3, 1), wd, bd, 2, 0)
    assert y.shape == (N, yr.size(2), yr.size(3), Co)
    assert rel_err(y.permute(0, 3, 1, 2), yr) < 2e-5
    y.backward(gy.to(dev()).permute(0, 2, 3, 1))
    assert rel_err(xd.grad, xr.grad) < 2e-5
    assert rel_err(wd.grad, wr.grad) < 2e-5
    assert rel_err(bd.grad, br.grad) < 1e-4
  finally:
    ops.set_conv_math('fp32')


# ---------------------------------------------------------------------------
# full benchmark sizes (VG-128, batch 32): size-independent properties
# ---------------------------------------------------------------------------

def test_full_size_conv_tensor_core_vs_exact_fp32_kernel():
  """CRN stage-4 conv1 at the benchmark size (32 x 128 x 128, 288 -> 64): with
  TF32-exact operands the tcgen05 kernels (halo forward, dgrad, wgrad) must
  reproduce the exact-fp32 FFMA kernels, which the small-size tests pin to the
  CPU oracle.  Also linearity of the forward at full size."""
  from sg2im_b200 import ops
  torch.manual_seed(0)
  N, H, W, Ci, Co = 32, 128, 128, 288, 64
  d = dev()
  x = _tf32_exact(torch.randn(N, H, W, Ci, device=d))
  w = _tf32_exact(torch.randn(Co, Ci, 3, 3, device=d) * 0.05)
  gy = _tf32_exact(torch.randn(N, H, W, Co, device=d))
  res = {}
  for mode in ('fp32', 'tf32'):
    ops.set_conv_math(mode)
    xx = x.clone().requires_grad_(True)
    ww = w.clone().requires_grad_(True)
    y = ops.conv2d(xx, ww, None, 1, 1)
    y.backward(gy)
    res[mode] = (y.detach(), xx.grad, ww.grad)
  ops.set_conv_math('fp32')
  for a, b, name in zip(res['tf32'], res['fp32'], ('fwd', 'dgrad', 'wgrad')):
    assert rel_err(a, b) < 5e-5, name
  ops.set_conv_math('tf32')
  try:
    x2 = _tf32_exact(torch.randn(N, H, W, Ci, device=d))
    y1 = ops.conv2d(x, w, None, 1, 1)
    y2 = ops.conv2d(x2, w, None, 1, 1)
    y12 = ops.conv2d(_tf32_exact(0.5 * x + 0.25 * x2), w, None, 1, 1)
    # 0.5*x + 0.25*x2 is TF32-exact only up to one rounding: tolerance of one TF32 ulp
    assert rel_err(y12, 0.5 * y1 + 0.25 * y2) < 2e-3
  finally:
    ops.set_conv_math('fp32')


def test_full_size_layout_linearity_and_mass():
  """masks_to_layout at the benchmark size (O=320, N=32, D=128, 128x128): linear
  in the object vectors; with all-ones masks and the full-image box the layout
  equals the vector at every pixel (the bilinear weights of a constant mask sum
  to 1 inside the box)."""
  from sg2im_b200.layout import masks_to_layout
  from sg2im_b200.synth import synth_config
  (imgs, objs, boxes, triples, o2i, _), cfg = synth_config('vg128', seed=5)
  d = dev()
  O = objs.numel()
  g = torch.Generator().manual_seed(1)
  v1, v2 = torch.randn(O, 128, generator=g).to(d), torch.randn(O, 128, generator=g).to(d)
  masks = torch.rand(O, 16, 16, generator=g).to(d)
  b, o = boxes.to(d), o2i.to(d)
  l1 = masks_to_layout(v1, b, masks, o, 128, 128, num_imgs=32)
  l2 = masks_to_layout(v2, b, masks, o, 128, 128, num_imgs=32)
  l12 = masks_to_layout(2.0 * v1 - 0.5 * v2, b, masks, o, 128, 128, num_imgs=32)
  assert rel_err(l12, 2.0 * l1 - 0.5 * l2) < 1e-5
  # only the __image__ objects (box [0,0,1,1], last of every image), all-ones masks
  keep = torch.zeros(O, 1, device=d)
  keep[9::10] = 1.0
  lay = masks_to_layout(v1 * keep, b, torch.ones(O, 16, 16, device=d), o, 128, 128, num_imgs=32)
  want = v1[9::10].view(32, 128, 1, 1).expand(32, 128, 128, 128)
  inner = (slice(None), slice(None), slice(4, 124), slice(4, 124))
  assert rel_err(lay[inner], want[inner]) < 1e-5
