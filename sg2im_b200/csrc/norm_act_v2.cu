// Second-generation BatchNorm + LeakyReLU (+ nearest-x2 upsample) BACKWARD
// kernels — same mathematics as scale_act_bwd_reduce / scale_act_bwd_apply in
// norm_act.cu, restructured for memory-level parallelism.
//
// Why: the first-generation reduce kernel reaches ~20 % of HBM bandwidth and the
// apply kernel ~37 % (profiles/r01_kernel_table_tf32.txt: 1.04 + 0.80 ms per
// step for ~3.3 GB of algorithmic traffic).  Its SASS shows why: the generic
// ActGrad functor carries run-time `up`, `scale != NULL`, `save != NULL`
// branches, so the four rows a thread has "in flight" are in fact loaded one
// after the other (each behind its own branches), and the per-channel constants
// (scale, shift, mean, invstd) are re-read from L1 for every row; the apply
// kernel additionally performs 8 fp64 divisions per float4.  Here
//   * UP is a template parameter and BN is mandatory (the no-BN / eval cases stay
//     on the generic kernels), so the row loop has no branches;
//   * every thread owns ONE group of 4 channels: its constants live in registers,
//     loaded once; per-channel means of the reduced sums are divided once per
//     thread, not per element;
//   * R rows are processed per iteration (R = 4 for up = 1: 8 independent 16-byte
//     loads; R = 2 for up = 2: 10) with no control flow between them; ptxas
//     schedules ~5 of them back to back (cuobjdump -sass), and 3-4 resident CTAs
//     per SM supply the rest of the latency-bandwidth product.
// Default since round 2 (validated and timed on the B200: apply 6.0 TB/s, reduce 3.0-3.7 TB/s,
// profiles/r02_hbm_kernels_ncu.txt); SG2IM_BNBWD_V2=0 selects the first-generation kernels (norm_act.cu).
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// per-thread channel-group constants
struct Chan4 {
  float4 scale, shift, mean, invstd;
};

__device__ __forceinline__ Chan4 load_chan(const float* scale, const float* shift,
                                           const float* save, uint32_t C, uint32_t c) {
  Chan4 k;
  k.scale = ld4(scale + c); k.shift = ld4(shift + c);
  k.mean = ld4(save + c);   k.invstd = ld4(save + C + c);
  return k;
}

// leaky'(x*scale+shift) applied to g, component-wise
__device__ __forceinline__ float4 act_grad(float4 g, float4 x, const Chan4& k, float slope) {
  g.x *= fmaf(x.x, k.scale.x, k.shift.x) > 0.f ? 1.f : slope;
  g.y *= fmaf(x.y, k.scale.y, k.shift.y) > 0.f ? 1.f : slope;
  g.z *= fmaf(x.z, k.scale.z, k.shift.z) > 0.f ? 1.f : slope;
  g.w *= fmaf(x.w, k.scale.w, k.shift.w) > 0.f ? 1.f : slope;
  return g;
}

__device__ __forceinline__ float4 xhat4(float4 x, const Chan4& k) {
  return make_float4((x.x - k.mean.x) * k.invstd.x, (x.y - k.mean.y) * k.invstd.y,
                     (x.z - k.mean.z) * k.invstd.z, (x.w - k.mean.w) * k.invstd.w);
}

// dy address of input row m (= (n, yy, xx) flattened) at sub-position (0,0)
template <int UP>
__device__ __forceinline__ const float* dy_row(const float* dy, uint32_t m, uint32_t H, uint32_t W,
                                               uint32_t dcs) {
  if (UP == 1) return dy + (size_t)m * dcs;
  uint32_t xx = m % W, t = m / W;                  // t = n*H + yy
  // output pixel ((n*H + yy)*2) * (2W) + 2*xx = t*4W + 2*xx
  return dy + ((size_t)t * 4u * W + 2u * xx) * dcs;
}

// sum of the UP x UP output gradients feeding input row m; loads only
template <int UP>
struct DyLoads {
  float4 v[UP * UP];
  __device__ __forceinline__ void load(const float* base, uint32_t W, uint32_t dcs) {
    if (UP == 1) {
      v[0] = ld4(base);
    } else {
      const size_t row = (size_t)2u * W * dcs;     // one output row down
      v[0] = ld4(base); v[1] = ld4(base + dcs);
      v[2] = ld4(base + row); v[3] = ld4(base + row + dcs);
    }
  }
  __device__ __forceinline__ float4 sum() const {
    if (UP == 1) return v[0];
    // same association as the generic kernel: ((a + b) + c) + d from zero
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < UP * UP; ++i) { g.x += v[i].x; g.y += v[i].y; g.z += v[i].z; g.w += v[i].w; }
    return g;
  }
};

template <int UP> struct RowsPerIter { static constexpr int value = UP == 1 ? 4 : 2; };

// ---------------------------------------------------------------- reduce ---
// sums[c] += sum_m g, sums[C+c] += sum_m g * xhat.  Block = TX channel groups x
// TY = 256/TX row lanes; grid (cblocks, rblocks).
template <int UP>
__global__ void __launch_bounds__(256, UP == 1 ? 4 : 3)
bn_bwd_reduce_v2_kernel(const float* __restrict__ dy, uint32_t dcs, uint32_t dco,
                        const float* __restrict__ x, uint32_t H, uint32_t W, uint32_t C,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ save, float slope, uint32_t M,
                        uint32_t rows_per_block, double* __restrict__ sums, int TX) {
  constexpr int R = RowsPerIter<UP>::value;
  __shared__ double sh[2][4][256];
  const uint32_t tx = threadIdx.x % TX, ty = threadIdx.x / TX, TY = 256 / TX;
  const uint32_t c = (blockIdx.x * TX + tx) * 4;
  const uint32_t mb = blockIdx.y * rows_per_block;
  const uint32_t me = min(mb + rows_per_block, M);
  double d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0};
  if (c < C) {
    const Chan4 k = load_chan(scale, shift, save, C, c);
    const float* xc = x + c;
    const float* dyc = dy + dco + c;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    int cnt = 0;
    uint32_t m = mb + ty;
    for (; m + (R - 1) * TY < me; m += R * TY) {
      float4 xv[R];
      DyLoads<UP> g[R];
#pragma unroll
      for (int r = 0; r < R; ++r) xv[r] = ld4(xc + (size_t)(m + r * TY) * C);
#pragma unroll
      for (int r = 0; r < R; ++r) g[r].load(dy_row<UP>(dyc, m + r * TY, H, W, dcs), W, dcs);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float4 gg = act_grad(g[r].sum(), xv[r], k, slope);
        float4 xh = xhat4(xv[r], k);
        s0.x += gg.x; s0.y += gg.y; s0.z += gg.z; s0.w += gg.w;
        s1.x += gg.x * xh.x; s1.y += gg.y * xh.y; s1.z += gg.z * xh.z; s1.w += gg.w * xh.w;
      }
      if (++cnt == 8) {                               // bound the fp32 partial sums
        d0[0] += s0.x; d0[1] += s0.y; d0[2] += s0.z; d0[3] += s0.w;
        d1[0] += s1.x; d1[1] += s1.y; d1[2] += s1.z; d1[3] += s1.w;
        s0 = make_float4(0.f, 0.f, 0.f, 0.f); s1 = s0; cnt = 0;
      }
    }
    for (; m < me; m += TY) {
      float4 xv = ld4(xc + (size_t)m * C);
      DyLoads<UP> g;
      g.load(dy_row<UP>(dyc, m, H, W, dcs), W, dcs);
      float4 gg = act_grad(g.sum(), xv, k, slope);
      float4 xh = xhat4(xv, k);
      s0.x += gg.x; s0.y += gg.y; s0.z += gg.z; s0.w += gg.w;
      s1.x += gg.x * xh.x; s1.y += gg.y * xh.y; s1.z += gg.z * xh.z; s1.w += gg.w * xh.w;
    }
    d0[0] += s0.x; d0[1] += s0.y; d0[2] += s0.z; d0[3] += s0.w;
    d1[0] += s1.x; d1[1] += s1.y; d1[2] += s1.z; d1[3] += s1.w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[0][j][threadIdx.x] = d0[j]; sh[1][j][threadIdx.x] = d1[j]; }
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double a = 0, b = 0;
      for (uint32_t y = 0; y < TY; ++y) { a += sh[0][j][y * TX + tx]; b += sh[1][j][y * TX + tx]; }
      atomicAdd(sums + c + j, a);
      atomicAdd(sums + C + c + j, b);
    }
  }
}

// ----------------------------------------------------------------- apply ---
// dx = scale * (g - mean(g) - xhat * mean(g*xhat)), train-mode BatchNorm.
template <int UP>
__global__ void __launch_bounds__(256, 4)
bn_bwd_apply_v2_kernel(const float* __restrict__ dy, uint32_t dcs, uint32_t dco,
                       const float* __restrict__ x, uint32_t H, uint32_t W, uint32_t C,
                       const float* __restrict__ scale, const float* __restrict__ shift,
                       const float* __restrict__ save, float slope, uint32_t M,
                       uint32_t rows_per_block, const double* __restrict__ sums,
                       float* __restrict__ dx, int TX) {
  constexpr int R = RowsPerIter<UP>::value;
  const uint32_t tx = threadIdx.x % TX, ty = threadIdx.x / TX, TY = 256 / TX;
  const uint32_t c = (blockIdx.x * TX + tx) * 4;
  if (c >= C) return;
  const uint32_t mb = blockIdx.y * rows_per_block;
  const uint32_t me = min(mb + rows_per_block, M);
  const Chan4 k = load_chan(scale, shift, save, C, c);
  // the generic kernel's (float)(sums[c] / (double)M), once per thread
  float mg[4], mgx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mg[j] = (float)(sums[c + j] / (double)M);
    mgx[j] = (float)(sums[C + c + j] / (double)M);
  }
  const float* xc = x + c;
  const float* dyc = dy + dco + c;
  float* dxc = dx + c;
  auto finish = [&](float4 gsum, float4 xv) {
    float4 g = act_grad(gsum, xv, k, slope);
    float4 xh = xhat4(xv, k);
    return make_float4(k.scale.x * (g.x - mg[0] - xh.x * mgx[0]),
                       k.scale.y * (g.y - mg[1] - xh.y * mgx[1]),
                       k.scale.z * (g.z - mg[2] - xh.z * mgx[2]),
                       k.scale.w * (g.w - mg[3] - xh.w * mgx[3]));
  };
  uint32_t m = mb + ty;
  for (; m + (R - 1) * TY < me; m += R * TY) {
    float4 xv[R];
    DyLoads<UP> g[R];
#pragma unroll
    for (int r = 0; r < R; ++r) xv[r] = ld4(xc + (size_t)(m + r * TY) * C);
#pragma unroll
    for (int r = 0; r < R; ++r) g[r].load(dy_row<UP>(dyc, m + r * TY, H, W, dcs), W, dcs);
#pragma unroll
    for (int r = 0; r < R; ++r)
      *reinterpret_cast<float4*>(dxc + (size_t)(m + r * TY) * C) = finish(g[r].sum(), xv[r]);
  }
  for (; m < me; m += TY) {
    float4 xv = ld4(xc + (size_t)m * C);
    DyLoads<UP> g;
    g.load(dy_row<UP>(dyc, m, H, W, dcs), W, dcs);
    *reinterpret_cast<float4*>(dxc + (size_t)m * C) = finish(g.sum(), xv);
  }
}

// channel-group tiling shared by both kernels (same rule as launch_colreduce4)
void tile_channels(int64_t C, int& TX, int64_t& cblocks) {
  int64_t groups = C / 4;
  TX = 1;
  while (TX < 32 && TX < groups) TX <<= 1;
  cblocks = ceil_div64(groups, TX);
}

}  // namespace

// Preconditions (checked by the callers in norm_act.cu): float4 path eligible
// (C, dcs, dco multiples of 4, 16-byte aligned bases), scale/shift/save non-NULL,
// up in {1, 2}, N*H*W*C < 2^31 and (up*up) * N*H*W * dcs < 2^32 pixels*stride
// handled in size_t.
int sg2im_bn_bwd_reduce_v2(const float* dy, int64_t dcs, int64_t dco, const float* x, int64_t N,
                           int64_t H, int64_t W, int64_t C, const float* scale, const float* shift,
                           const float* save, float slope, int up, double* sums, cudaStream_t st) {
  const int64_t M = N * H * W;
  int TX; int64_t cblocks;
  tile_channels(C, TX, cblocks);
  const int TY = 256 / TX;
  const int R = up == 1 ? 4 : 2;
  // 3-4 CTAs per SM (register budget of the __launch_bounds__): ptxas keeps ~5 16-byte loads
  // in flight per thread, so 768 threads/SM cover the HBM latency-bandwidth product; more CTAs
  // would only lengthen the tail of same-address fp64 atomics (2*4*TX per CTA and channel group)
  int64_t want = ceil_div64(148 * (up == 1 ? 4 : 3), cblocks);   // up = 1 fits 64 registers: 4 CTAs per SM
  int64_t rpb = ceil_div64(M, want);
  if (rpb < (int64_t)R * TY) rpb = (int64_t)R * TY;
  int64_t rblocks = ceil_div64(M, rpb);
  if (rblocks > 65535) { rblocks = 65535; rpb = ceil_div64(M, rblocks); rblocks = ceil_div64(M, rpb); }
  dim3 grid((unsigned)cblocks, (unsigned)rblocks);
  if (up == 1)
    SG_LAUNCH(bn_bwd_reduce_v2_kernel<1>, grid, 256, 0, st, dy, (uint32_t)dcs, (uint32_t)dco, x, (uint32_t)H,
                                                     (uint32_t)W, (uint32_t)C, scale, shift, save,
                                                     slope, (uint32_t)M, (uint32_t)rpb, sums, TX);
  else
    SG_LAUNCH(bn_bwd_reduce_v2_kernel<2>, grid, 256, 0, st, dy, (uint32_t)dcs, (uint32_t)dco, x, (uint32_t)H,
                                                     (uint32_t)W, (uint32_t)C, scale, shift, save,
                                                     slope, (uint32_t)M, (uint32_t)rpb, sums, TX);
  return 0;
}

int sg2im_bn_bwd_apply_v2(const float* dy, int64_t dcs, int64_t dco, const float* x, int64_t N,
                          int64_t H, int64_t W, int64_t C, const float* scale, const float* shift,
                          const float* save, float slope, int up, const double* sums, float* dx,
                          cudaStream_t st) {
  const int64_t M = N * H * W;
  int TX; int64_t cblocks;
  tile_channels(C, TX, cblocks);
  const int TY = 256 / TX;
  const int R = up == 1 ? 4 : 2;
  // no cross-CTA traffic here: two waves of the 4 resident CTAs per SM, each CTA at least 4
  // iterations deep
  int64_t want = ceil_div64(148 * 8, cblocks);
  int64_t rpb = ceil_div64(M, want);
  if (rpb < (int64_t)4 * R * TY) rpb = (int64_t)4 * R * TY;
  int64_t rblocks = ceil_div64(M, rpb);
  if (rblocks > 65535) { rblocks = 65535; rpb = ceil_div64(M, rblocks); rblocks = ceil_div64(M, rpb); }
  dim3 grid((unsigned)cblocks, (unsigned)rblocks);
  if (up == 1)
    SG_LAUNCH(bn_bwd_apply_v2_kernel<1>, grid, 256, 0, st, dy, (uint32_t)dcs, (uint32_t)dco, x, (uint32_t)H,
                                                    (uint32_t)W, (uint32_t)C, scale, shift, save,
                                                    slope, (uint32_t)M, (uint32_t)rpb, sums, dx, TX);
  else
    SG_LAUNCH(bn_bwd_apply_v2_kernel<2>, grid, 256, 0, st, dy, (uint32_t)dcs, (uint32_t)dco, x, (uint32_t)H,
                                                    (uint32_t)W, (uint32_t)C, scale, shift, save,
                                                    slope, (uint32_t)M, (uint32_t)rpb, sums, dx, TX);
  return 0;
}

// ---------------------------------------------------------------- forward ---
// y[n, Y, X, yco + c] = leaky(x[n, Y/UP, X/UP, c] * scale[c] + shift[c])   (UP in {1, 2})
// First generation: one OUTPUT float4 per thread, i.e. every input element is
// fetched UP*UP times and each output pays eight 32-bit divisions.  Here a thread
// owns an INPUT float4 (x is contiguous: element i of the float4 stream), computes
// it once and writes the UP*UP replicas; four elements per thread are in flight.
// Default since round 2 (5.8-6.9 TB/s on the B200); SG2IM_BNFWD_V2=0 selects the generic kernel (norm_act.cu).
namespace {

template <int UP>
__global__ void __launch_bounds__(256)
scale_act_fwd_v2_kernel(const float* __restrict__ x, uint32_t total /* float4s */, uint32_t W,
                        uint32_t C, const float* __restrict__ scale,
                        const float* __restrict__ shift, float slope, float* __restrict__ y,
                        uint32_t ycs, uint32_t yco, int rnd) {
  constexpr int ILP = 4;
  const uint32_t cg = C >> 2;
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i0 < total; i0 += ILP * stride) {
    float4 v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      uint32_t i = i0 + k * stride;
      v[k] = i < total ? ld4(x + (size_t)i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      uint32_t i = i0 + k * stride;
      if (i >= total) continue;
      const uint32_t m = i / cg, c = (i - m * cg) * 4;
      float4 r = v[k];
      if (scale) {
        float4 s = ld4(scale + c), b = ld4(shift + c);
        r.x = fmaf(r.x, s.x, b.x); r.y = fmaf(r.y, s.y, b.y);
        r.z = fmaf(r.z, s.z, b.z); r.w = fmaf(r.w, s.w, b.w);
      }
      r.x = leaky(r.x, slope); r.y = leaky(r.y, slope); r.z = leaky(r.z, slope); r.w = leaky(r.w, slope);
      if (rnd) { r.x = tf32_rn(r.x); r.y = tf32_rn(r.y); r.z = tf32_rn(r.z); r.w = tf32_rn(r.w); }
      if (UP == 1) {
        *reinterpret_cast<float4*>(y + (size_t)m * ycs + yco + c) = r;
      } else {
        const uint32_t xx = m % W, t = m / W;            // t = n*H + yy
        float* p = y + ((size_t)t * 4u * W + 2u * xx) * ycs + yco + c;
        const size_t row = (size_t)2u * W * ycs;
        *reinterpret_cast<float4*>(p) = r;
        *reinterpret_cast<float4*>(p + ycs) = r;
        *reinterpret_cast<float4*>(p + row) = r;
        *reinterpret_cast<float4*>(p + row + ycs) = r;
      }
    }
  }
}

}  // namespace

// Preconditions (checked by sg2im_scale_act_fwd): float4 path eligible, up in {1, 2},
// N*H*W*C < 2^31, y_cstride < 2^31.
int sg2im_scale_act_fwd_v2(const float* x, int64_t N, int64_t H, int64_t W, int64_t C,
                           const float* scale, const float* shift, float slope, int up, float* y,
                           int64_t ycs, int64_t yco, int rnd, cudaStream_t st) {
  const int64_t total = N * H * W * (C / 4);
  int64_t blocks = ceil_div64(total, 256 * 4);
  if (blocks < 1) blocks = 1;
  if (up == 1)
    SG_LAUNCH(scale_act_fwd_v2_kernel<1>, (unsigned)blocks, 256, 0, st, x, (uint32_t)total, (uint32_t)W,
                                                                 (uint32_t)C, scale, shift, slope, y,
                                                                 (uint32_t)ycs, (uint32_t)yco, rnd);
  else
    SG_LAUNCH(scale_act_fwd_v2_kernel<2>, (unsigned)blocks, 256, 0, st, x, (uint32_t)total, (uint32_t)W,
                                                                 (uint32_t)C, scale, shift, slope, y,
                                                                 (uint32_t)ycs, (uint32_t)yco, rnd);
  return 0;
}

// ----------------------------------------------------------- small colsum ---
// Bias gradients of the small GEMMs (scene-graph MLPs: 320/448 rows; discriminator
// heads): out[c] = sum_m x[m, c].  The generic path is three launches (zero fp64
// scratch, reduce with atomics, convert); for M <= 1024 one CTA per 32 columns
// finishes the job in a single launch with no atomics and no scratch
// (SG2IM_COLSUM_V2=0 disables it).
namespace {

__global__ void __launch_bounds__(256)
colsum_small_kernel(const float* __restrict__ x, uint32_t M, uint32_t C, float* __restrict__ out) {
  __shared__ double sh[8][33];
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const uint32_t c = blockIdx.x * 32 + tx;
  double acc = 0.0;
  if (c < C) {
    uint32_t m = ty;
    for (; m + 24 < M; m += 32) {                       // 4 independent loads per iteration
      float a = x[(size_t)m * C + c], b = x[(size_t)(m + 8) * C + c];
      float d = x[(size_t)(m + 16) * C + c], e = x[(size_t)(m + 24) * C + c];
      acc += (double)a + (double)b + (double)d + (double)e;
    }
    for (; m < M; m += 8) acc += (double)x[(size_t)m * C + c];
  }
  sh[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
    double s = 0.0;
#pragma unroll
    for (int y = 0; y < 8; ++y) s += sh[y][tx];
    out[c] = (float)s;
  }
}

}  // namespace

int sg2im_colsum_small(const float* x, int64_t M, int64_t C, float* out, cudaStream_t st) {
  SG_LAUNCH(colsum_small_kernel, (unsigned)ceil_div64(C, 32), 256, 0, st, x, (uint32_t)M, (uint32_t)C, out);
  return 0;
}

// -------------------------------------------- activation backward + bias gradient ---
// dx = dy * leaky'(y)  AND  db[c] += sum_m dx[m, c]  in ONE pass over dy (the generic path is
// act_bwd followed by the three-launch colsum, i.e. two passes over the gradient and four
// launches).  db is ACCUMULATED into (float atomics of per-CTA fp64 partials): pass a zeroed
// buffer, or the bias' slot of the flat gradient bucket.  Opt-in (ops.FUSE_ACT_BWD).
namespace {

__global__ void __launch_bounds__(256)
act_bwd_colsum_kernel(const float* __restrict__ dy, const float* __restrict__ y, float slope,
                      uint32_t M, uint32_t C, uint32_t rows_per_block, float* __restrict__ dx,
                      float* __restrict__ db, int TX) {
  __shared__ double sh[4][256];
  const uint32_t tx = threadIdx.x % TX, ty = threadIdx.x / TX, TY = 256 / TX;
  const uint32_t c = (blockIdx.x * TX + tx) * 4;
  const uint32_t mb = blockIdx.y * rows_per_block;
  const uint32_t me = min(mb + rows_per_block, M);
  double d[4] = {0, 0, 0, 0};
  if (c < C) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
    for (uint32_t m = mb + ty; m < me; m += TY) {
      const size_t o = (size_t)m * C + c;
      float4 g = ld4(dy + o), yy = ld4(y + o);
      g.x *= yy.x > 0.f ? 1.f : slope; g.y *= yy.y > 0.f ? 1.f : slope;
      g.z *= yy.z > 0.f ? 1.f : slope; g.w *= yy.w > 0.f ? 1.f : slope;
      *reinterpret_cast<float4*>(dx + o) = g;
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
      if (++cnt == 32) {
        d[0] += s.x; d[1] += s.y; d[2] += s.z; d[3] += s.w;
        s = make_float4(0.f, 0.f, 0.f, 0.f); cnt = 0;
      }
    }
    d[0] += s.x; d[1] += s.y; d[2] += s.z; d[3] += s.w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sh[j][threadIdx.x] = d[j];
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double a = 0;
      for (uint32_t r = 0; r < TY; ++r) a += sh[j][r * TX + tx];
      atomicAdd(db + c + j, (float)a);
    }
  }
}

}  // namespace

extern "C" int sg2im_act_bwd_colsum(const float* dy, const float* y, float slope, int64_t M, int64_t C,
                                    float* dx, float* db, sg2im_stream_t stream) {
  SG_ARG(dy && y && dx && db && M >= 1 && C >= 4 && C % 4 == 0 && M * C < (1ll << 31));
  SG_ARG(aligned16(dy) && aligned16(y) && aligned16(dx));
  int TX; int64_t cblocks;
  tile_channels(C, TX, cblocks);
  const int TY = 256 / TX;
  int64_t want = ceil_div64(148 * 4, cblocks);
  int64_t rpb = ceil_div64(M, want);
  if (rpb < (int64_t)4 * TY) rpb = (int64_t)4 * TY;
  int64_t rblocks = ceil_div64(M, rpb);
  if (rblocks > 65535) { rblocks = 65535; rpb = ceil_div64(M, rblocks); rblocks = ceil_div64(M, rpb); }
  dim3 grid((unsigned)cblocks, (unsigned)rblocks);
  SG_LAUNCH(act_bwd_colsum_kernel, grid, 256, 0, as_stream(stream), dy, y, slope, (uint32_t)M,
            (uint32_t)C, (uint32_t)rpb, dx, db, TX);
  SG_LAUNCH_OK();
  return 0;
}
