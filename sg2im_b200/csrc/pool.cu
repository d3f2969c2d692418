// Non-overlapping 2-D pooling on NHWC activations: nn.MaxPool2d / nn.AvgPool2d with
// kernel_size = stride = f, what build_cnn's 'PX' token builds (sg2im/layers.py:195-201).
// Floor mode like torch: trailing rows / columns that do not fill a window are dropped
// (their gradient is zero).  HBM-bound: one thread per (window, channel vector); the channel
// index is fastest so a warp reads consecutive addresses in every tap of the window.
//
//   max: ties go to the first element in row-major window order (strict '>' scan, like ATen);
//        the backward pass re-derives the arg-max from x — every input element belongs to
//        exactly one window (kernel = stride), so it writes all f*f gradients of its window
//        without atomics or a zero fill.
//   avg: y = (sum of the window in row-major order) / (f*f), dx = dy / (f*f).
#include "common.cuh"

namespace {

template <int VEC> struct Vec;
template <> struct Vec<1> { typedef float T; };
template <> struct Vec<4> { typedef float4 T; };

__device__ __forceinline__ float lane(float v, int) { return v; }
__device__ __forceinline__ float lane(const float4& v, int i) {
  return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
}

__device__ __forceinline__ void store(float* p, const float (&a)[1]) { *p = a[0]; }
__device__ __forceinline__ void store(float* p, const float (&a)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
}

// decompose a flat (window, channel-group) index
struct Win {
  uint32_t n, yo, xo, c;
};
__device__ __forceinline__ Win decompose(uint32_t ii, uint32_t Ho, uint32_t Wo, uint32_t cg, int V) {
  Win w;
  w.c = (ii % cg) * (uint32_t)V;
  uint32_t t = ii / cg;
  w.xo = t % Wo; t /= Wo;
  w.yo = t % Ho;
  w.n = t / Ho;
  return w;
}

template <int MODE, int VEC>   // MODE 0 = average, 1 = max
__global__ void pool_fwd_kernel(const float* __restrict__ x, uint32_t N, uint32_t H, uint32_t W,
                                uint32_t C, int f, float* __restrict__ y) {
  uint32_t Ho = H / (uint32_t)f, Wo = W / (uint32_t)f, cg = C / VEC;
  uint32_t ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= N * Ho * Wo * cg) return;
  Win w = decompose(ii, Ho, Wo, cg, VEC);
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
  for (int dy = 0; dy < f; ++dy) {
    for (int dx = 0; dx < f; ++dx) {
      int64_t off = (((int64_t)w.n * H + (w.yo * f + dy)) * W + (w.xo * f + dx)) * C + w.c;
      typename Vec<VEC>::T v = *reinterpret_cast<const typename Vec<VEC>::T*>(x + off);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float e = lane(v, k);
        if (MODE == 0) acc[k] += e;
        else if ((dy | dx) == 0 || e > acc[k] || e != e) acc[k] = e;     // NaN propagates like ATen
      }
    }
  }
  float inv = 1.f / (float)(f * f);
  int64_t o = (((int64_t)w.n * Ho + w.yo) * Wo + w.xo) * C + w.c;
  if (MODE == 0) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] *= inv;
  }
  store(y + o, acc);
}

template <int MODE, int VEC>
__global__ void pool_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                uint32_t N, uint32_t H, uint32_t W, uint32_t C, int f,
                                float* __restrict__ dx) {
  uint32_t Ho = H / (uint32_t)f, Wo = W / (uint32_t)f, cg = C / VEC;
  uint32_t ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= N * Ho * Wo * cg) return;
  Win w = decompose(ii, Ho, Wo, cg, VEC);
  int64_t o = (((int64_t)w.n * Ho + w.yo) * Wo + w.xo) * C + w.c;
  float g[VEC];
  int best[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { g[k] = dy[o + k]; best[k] = 0; }
  if (MODE == 1) {
    float m[VEC];
    for (int t = 0; t < f * f; ++t) {
      int64_t off = (((int64_t)w.n * H + (w.yo * f + t / f)) * W + (w.xo * f + t % f)) * C + w.c;
      typename Vec<VEC>::T v = *reinterpret_cast<const typename Vec<VEC>::T*>(x + off);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float e = lane(v, k);
        if (t == 0 || e > m[k] || e != e) { m[k] = e; best[k] = t; }
      }
    }
  }
  float inv = 1.f / (float)(f * f);
  for (int t = 0; t < f * f; ++t) {
    int64_t off = (((int64_t)w.n * H + (w.yo * f + t / f)) * W + (w.xo * f + t % f)) * C + w.c;
    float r[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) r[k] = MODE == 0 ? g[k] * inv : (best[k] == t ? g[k] : 0.f);
    store(dx + off, r);
  }
}

bool pool_dims_ok(int64_t N, int64_t H, int64_t W, int64_t C, int f) {
  return N >= 1 && H >= 1 && W >= 1 && C >= 1 && f >= 1 && f <= 64 && H / f >= 1 && W / f >= 1 &&
         N * H * W * C < (1ll << 31);
}

}  // namespace

extern "C" int sg2im_pool2d_fwd(const float* x, int64_t N, int64_t H, int64_t W, int64_t C,
                                int factor, int mode, float* y, sg2im_stream_t stream) {
  SG_ARG(x && y && (mode == 0 || mode == 1));
  SG_ARG(pool_dims_ok(N, H, W, C, factor));
  bool vec = C % 4 == 0 && aligned16(x) && aligned16(y);
  int64_t total = N * (H / factor) * (W / factor) * (C / (vec ? 4 : 1));
  unsigned grid = (unsigned)ceil_div64(total, 256);
  cudaStream_t st = as_stream(stream);
  uint32_t n = (uint32_t)N, h = (uint32_t)H, w = (uint32_t)W, c = (uint32_t)C;
  if (mode == 0 && vec)       SG_LAUNCH((pool_fwd_kernel<0, 4>), grid, 256, 0, st, x, n, h, w, c, factor, y);
  else if (mode == 0)         SG_LAUNCH((pool_fwd_kernel<0, 1>), grid, 256, 0, st, x, n, h, w, c, factor, y);
  else if (vec)               SG_LAUNCH((pool_fwd_kernel<1, 4>), grid, 256, 0, st, x, n, h, w, c, factor, y);
  else                        SG_LAUNCH((pool_fwd_kernel<1, 1>), grid, 256, 0, st, x, n, h, w, c, factor, y);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_pool2d_bwd(const float* dy, const float* x, int64_t N, int64_t H, int64_t W,
                                int64_t C, int factor, int mode, float* dx,
                                sg2im_stream_t stream) {
  SG_ARG(dy && dx && (mode == 0 || (mode == 1 && x)));
  SG_ARG(pool_dims_ok(N, H, W, C, factor));
  bool vec = C % 4 == 0 && aligned16(dy) && aligned16(dx) && (mode == 0 || aligned16(x));
  int64_t total = N * (H / factor) * (W / factor) * (C / (vec ? 4 : 1));
  unsigned grid = (unsigned)ceil_div64(total, 256);
  cudaStream_t st = as_stream(stream);
  uint32_t n = (uint32_t)N, h = (uint32_t)H, w = (uint32_t)W, c = (uint32_t)C;
  if (mode == 0 && vec)       SG_LAUNCH((pool_bwd_kernel<0, 4>), grid, 256, 0, st, dy, x, n, h, w, c, factor, dx);
  else if (mode == 0)         SG_LAUNCH((pool_bwd_kernel<0, 1>), grid, 256, 0, st, dy, x, n, h, w, c, factor, dx);
  else if (vec)               SG_LAUNCH((pool_bwd_kernel<1, 4>), grid, 256, 0, st, dy, x, n, h, w, c, factor, dx);
  else                        SG_LAUNCH((pool_bwd_kernel<1, 1>), grid, 256, 0, st, dy, x, n, h, w, c, factor, dx);
  SG_LAUNCH_OK();
  return 0;
}
