// Differentiable box crops for the object discriminator: one gather kernel
// reading feats[idx[b]] directly — replaces the per-image Python loop with
// nonzero() host syncs, expand().contiguous() copies and F.grid_sample of
// sg2im/bilinear.py:28-132 (+ tensor_linspace :249-278).  HBM/latency-bound.
#include "common.cuh"

namespace {

// sampling coordinate j of `steps` between 2*b0-1 and 2*b1-1 (tensor_linspace:
// start_w*start + end_w*end with start_w = linspace(1,0), end_w = linspace(0,1))
__device__ __forceinline__ float crop_coord(int j, int steps, float b0, float b1) {
  float a = steps > 1 ? (float)j / (float)(steps - 1) : 0.f;
  float s = 2.f * b0 - 1.f, e = 2.f * b1 - 1.f;
  return (1.f - a) * s + a * e;
}

__global__ void crop_fwd_kernel(const float* __restrict__ feats, int64_t sfn, int64_t sfh,
                                int64_t sfw, int64_t sfc, int64_t N, int H, int W, int C,
                                const float* __restrict__ boxes, const int64_t* __restrict__ idx,
                                int64_t B, int HH, int WW, int align, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HH * WW) return;
  int j = (int)(i % WW);
  int64_t t = i / WW;
  int ii = (int)(t % HH);
  int64_t b = t / HH;
  int64_t n = idx[b];
  float* op = out + i * C;
  if (n < 0 || n >= N) { for (int c = 0; c < C; ++c) op[c] = 0.f; return; }
  float4 bx = *reinterpret_cast<const float4*>(boxes + b * 4);
  int xl, yl; float wx, wy;
  bilinear_axis(crop_coord(j, WW, bx.x, bx.z), W, align, xl, wx);
  bilinear_axis(crop_coord(ii, HH, bx.y, bx.w), H, align, yl, wy);
  const float* base = feats + n * sfn;
  bool x0 = xl >= 0 && xl < W, x1 = xl + 1 >= 0 && xl + 1 < W;
  bool y0 = yl >= 0 && yl < H, y1 = yl + 1 >= 0 && yl + 1 < H;
  float w00 = (1.f - wx) * (1.f - wy), w01 = wx * (1.f - wy), w10 = (1.f - wx) * wy, w11 = wx * wy;
  for (int c = 0; c < C; ++c) {
    const float* pc = base + (int64_t)c * sfc;
    float v = 0.f;
    if (y0 && x0) v += pc[(int64_t)yl * sfh + (int64_t)xl * sfw] * w00;
    if (y0 && x1) v += pc[(int64_t)yl * sfh + (int64_t)(xl + 1) * sfw] * w01;
    if (y1 && x0) v += pc[(int64_t)(yl + 1) * sfh + (int64_t)xl * sfw] * w10;
    if (y1 && x1) v += pc[(int64_t)(yl + 1) * sfh + (int64_t)(xl + 1) * sfw] * w11;
    op[c] = v;
  }
}

__global__ void crop_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ boxes,
                                const int64_t* __restrict__ idx, int64_t N, int H, int W, int C,
                                int64_t B, int HH, int WW, int align, float* __restrict__ dfeats) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HH * WW) return;
  int j = (int)(i % WW);
  int64_t t = i / WW;
  int ii = (int)(t % HH);
  int64_t b = t / HH;
  int64_t n = idx[b];
  if (n < 0 || n >= N) return;
  float4 bx = *reinterpret_cast<const float4*>(boxes + b * 4);
  int xl, yl; float wx, wy;
  bilinear_axis(crop_coord(j, WW, bx.x, bx.z), W, align, xl, wx);
  bilinear_axis(crop_coord(ii, HH, bx.y, bx.w), H, align, yl, wy);
  bool x0 = xl >= 0 && xl < W, x1 = xl + 1 >= 0 && xl + 1 < W;
  bool y0 = yl >= 0 && yl < H, y1 = yl + 1 >= 0 && yl + 1 < H;
  float w00 = (1.f - wx) * (1.f - wy), w01 = wx * (1.f - wy), w10 = (1.f - wx) * wy, w11 = wx * wy;
  const float* dp = dout + i * C;
  float* base = dfeats + n * (int64_t)H * W * C;
  for (int c = 0; c < C; ++c) {
    float g = dp[c];
    if (y0 && x0) atomicAdd(base + ((int64_t)yl * W + xl) * C + c, g * w00);
    if (y0 && x1) atomicAdd(base + ((int64_t)yl * W + xl + 1) * C + c, g * w01);
    if (y1 && x0) atomicAdd(base + ((int64_t)(yl + 1) * W + xl) * C + c, g * w10);
    if (y1 && x1) atomicAdd(base + ((int64_t)(yl + 1) * W + xl + 1) * C + c, g * w11);
  }
}

}  // namespace

extern "C" int sg2im_crop_fwd(const float* feats, int64_t sfn, int64_t sfh, int64_t sfw,
                              int64_t sfc, int64_t N, int64_t H, int64_t W, int64_t C,
                              const float* boxes, const int64_t* idx, int64_t B, int64_t HH,
                              int64_t WW, int align_corners, float* out, sg2im_stream_t stream) {
  SG_ARG(feats && boxes && idx && out);
  SG_ARG(N >= 1 && H >= 1 && W >= 1 && C >= 1 && B >= 0 && HH >= 1 && WW >= 1);
  SG_ARG(aligned16(boxes));
  if (B == 0) return 0;
  int64_t total = B * HH * WW;
  SG_LAUNCH(crop_fwd_kernel, (unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream), 
      feats, sfn, sfh, sfw, sfc, N, (int)H, (int)W, (int)C, boxes, idx, B, (int)HH, (int)WW,
      align_corners, out);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_crop_bwd(const float* dout, const float* boxes, const int64_t* idx,
                              int64_t N, int64_t H, int64_t W, int64_t C, int64_t B, int64_t HH,
                              int64_t WW, int align_corners, float* dfeats,
                              sg2im_stream_t stream) {
  SG_ARG(dout && boxes && idx && dfeats);
  SG_ARG(N >= 1 && H >= 1 && W >= 1 && C >= 1 && B >= 0 && HH >= 1 && WW >= 1);
  SG_ARG(aligned16(boxes));
  if (B == 0) return 0;
  int64_t total = B * HH * WW;
  SG_LAUNCH(crop_bwd_kernel, (unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream), 
      dout, boxes, idx, N, (int)H, (int)W, (int)C, B, (int)HH, (int)WW, align_corners, dfeats);
  SG_LAUNCH_OK();
  return 0;
}
