// Train-mode BatchNorm statistics, the fused affine + LeakyReLU + nearest-x2
// upsample + channel-slice write ("virtual concat"), their backward, the 2x2
// average-pool cascade, bias-gradient column sums and the epilogue-activation
// backward.  HBM-bound elementwise / reduction kernels (float4, coalesced on
// the NHWC channel axis, fp64 cross-CTA accumulation).
// Replaces nn.BatchNorm2d + nn.LeakyReLU + F.upsample + torch.cat +
// F.avg_pool2d of sg2im/crn.py:41-47,58-63,107 and model.py:98-99.
#include <cstdlib>
#include "common.cuh"

// second-generation BatchNorm kernels (norm_act_v2.cu): the defaults since round 2
int sg2im_bn_bwd_reduce_v2(const float* dy, int64_t dcs, int64_t dco, const float* x, int64_t N,
                           int64_t H, int64_t W, int64_t C, const float* scale, const float* shift,
                           const float* save, float slope, int up, double* sums, cudaStream_t st);
int sg2im_bn_bwd_apply_v2(const float* dy, int64_t dcs, int64_t dco, const float* x, int64_t N,
                          int64_t H, int64_t W, int64_t C, const float* scale, const float* shift,
                          const float* save, float slope, int up, const double* sums, float* dx,
                          cudaStream_t st);

int sg2im_scale_act_fwd_v2(const float* x, int64_t N, int64_t H, int64_t W, int64_t C,
                           const float* scale, const float* shift, float slope, int up, float* y,
                           int64_t ycs, int64_t yco, int rnd, cudaStream_t st);

int sg2im_colsum_small(const float* x, int64_t M, int64_t C, float* out, cudaStream_t st);

static bool bn_fwd_v2_enabled() {
  const char* e = getenv("SG2IM_BNFWD_V2");          // default since round 2 (validated on the B200); "0" = first generation
  return !(e && e[0] == '0');
}

static bool bn_bwd_v2_enabled() {
  const char* e = getenv("SG2IM_BNBWD_V2");          // read per call: tests toggle it in-process ("0" = first generation)
  return !(e && e[0] == '0');
}

namespace {

constexpr int RED_TX = 32, RED_TY = 8;

// ---- per-channel sums over rows: out[c] += sum f(x), out[C+c] += sum g(x) ----
// Functor F: (row m, channel c) -> (v0, v1).  Block = 32 channels x 8 row lanes.
template <class F>
__global__ void __launch_bounds__(RED_TX * RED_TY)
colreduce_kernel(F f, int64_t M, int64_t C, int64_t rows_per_block, double* __restrict__ sums,
                 int nout) {
  __shared__ double sh[2][RED_TY][RED_TX];
  int64_t c = (int64_t)blockIdx.x * RED_TX + threadIdx.x;
  int64_t mb = (int64_t)blockIdx.y * rows_per_block;
  int64_t me = mb + rows_per_block < M ? mb + rows_per_block : M;
  double d0 = 0.0, d1 = 0.0;
  if (c < C) {
    float s0 = 0.f, s1 = 0.f;
    int cnt = 0;
    for (int64_t m = mb + threadIdx.y; m < me; m += RED_TY) {
      float v0, v1;
      f(m, c, v0, v1);
      s0 += v0; s1 += v1;
      if (++cnt == 32) { d0 += s0; d1 += s1; s0 = s1 = 0.f; cnt = 0; }
    }
    d0 += s0; d1 += s1;
  }
  sh[0][threadIdx.y][threadIdx.x] = d0;
  sh[1][threadIdx.y][threadIdx.x] = d1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int y = 1; y < RED_TY; ++y) { d0 += sh[0][y][threadIdx.x]; d1 += sh[1][y][threadIdx.x]; }
    atomicAdd(sums + c, d0);
    if (nout > 1) atomicAdd(sums + C + c, d1);
  }
}

// float4 variant: thread = 4 consecutive channels, TX channel groups x (256/TX) row
// lanes per block, 2 rows in flight per thread.  F4::at4(m, c, v0, v1).
template <class F4>
__global__ void __launch_bounds__(256)
colreduce4_kernel(F4 f, int64_t M, int64_t C, int64_t rows_per_block, double* __restrict__ sums,
                  int nout, int TX) {
  __shared__ double sh[2][4][256];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX, TY = 256 / TX;
  int64_t c = ((int64_t)blockIdx.x * TX + tx) * 4;
  int64_t mb = (int64_t)blockIdx.y * rows_per_block;
  int64_t me = mb + rows_per_block < M ? mb + rows_per_block : M;
  double d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0};
  if (c < C) {
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    int cnt = 0;
    int64_t m = mb + ty;
    for (; m + 3 * TY < me; m += 4 * TY) {
      float4 a0, a1, b0, b1, c0, c1, e0, e1;
      f.at4(m, c, a0, a1);
      f.at4(m + TY, c, b0, b1);
      f.at4(m + 2 * TY, c, c0, c1);
      f.at4(m + 3 * TY, c, e0, e1);
      s0.x += (a0.x + b0.x) + (c0.x + e0.x); s0.y += (a0.y + b0.y) + (c0.y + e0.y);
      s0.z += (a0.z + b0.z) + (c0.z + e0.z); s0.w += (a0.w + b0.w) + (c0.w + e0.w);
      s1.x += (a1.x + b1.x) + (c1.x + e1.x); s1.y += (a1.y + b1.y) + (c1.y + e1.y);
      s1.z += (a1.z + b1.z) + (c1.z + e1.z); s1.w += (a1.w + b1.w) + (c1.w + e1.w);
      if (++cnt == 8) {
        d0[0] += s0.x; d0[1] += s0.y; d0[2] += s0.z; d0[3] += s0.w;
        d1[0] += s1.x; d1[1] += s1.y; d1[2] += s1.z; d1[3] += s1.w;
        s0 = make_float4(0.f, 0.f, 0.f, 0.f); s1 = s0; cnt = 0;
      }
    }
    for (; m < me; m += TY) {
      float4 a0, a1;
      f.at4(m, c, a0, a1);
      s0.x += a0.x; s0.y += a0.y; s0.z += a0.z; s0.w += a0.w;
      s1.x += a1.x; s1.y += a1.y; s1.z += a1.z; s1.w += a1.w;
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
      for (int y = 0; y < TY; ++y) { a += sh[0][j][y * TX + tx]; b += sh[1][j][y * TX + tx]; }
      atomicAdd(sums + c + j, a);
      if (nout > 1) atomicAdd(sums + C + c + j, b);
    }
  }
}

template <class F4>
int launch_colreduce4(F4 f, int64_t M, int64_t C, double* sums, int nout, cudaStream_t st) {
  int64_t groups = C / 4;
  int TX = 1;
  while (TX < 32 && TX < groups) TX <<= 1;
  int64_t cblocks = ceil_div64(groups, TX);
  int TY = 256 / TX;
  // every block ends with 2*4*TX same-address fp64 atomics per channel group: with
  // ~1200 blocks those serialised at the L2 (25 us per launch); ~2 blocks per SM keep
  // HBM busy (4 independent 16-byte loads in flight per thread) with 4x fewer atomics
  int64_t want = ceil_div64(148 * 2, cblocks);
  int64_t rpb = ceil_div64(M, want);
  if (rpb < 4 * TY) rpb = 4 * TY;
  int64_t rblocks = ceil_div64(M, rpb);
  if (rblocks > 65535) { rblocks = 65535; rpb = ceil_div64(M, rblocks); rblocks = ceil_div64(M, rpb); }
  dim3 grid((unsigned)cblocks, (unsigned)rblocks);
  SG_LAUNCH(colreduce4_kernel<F4>, grid, 256, 0, st, f, M, C, rpb, sums, nout, TX);
  return 0;
}

struct StatsF {
  const float* x; int64_t C;
  __device__ void operator()(int64_t m, int64_t c, float& v0, float& v1) const {
    float v = x[m * C + c];
    v0 = v; v1 = v * v;
  }
  __device__ __forceinline__ void at4(int64_t m, int64_t c, float4& v0, float4& v1) const {
    float4 v = *reinterpret_cast<const float4*>(x + m * C + c);
    v0 = v; v1 = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
  }
};

template <class F>
int launch_colreduce(F f, int64_t M, int64_t C, double* sums, int nout, cudaStream_t st) {
  int64_t cblocks = ceil_div64(C, RED_TX);
  // ~8 waves of CTAs, at least 64 rows each
  int64_t want = ceil_div64(148 * 8, cblocks);
  int64_t rpb = ceil_div64(M, want);
  if (rpb < 64) rpb = 64;
  int64_t rblocks = ceil_div64(M, rpb);
  if (rblocks > 65535) { rblocks = 65535; rpb = ceil_div64(M, rblocks); rblocks = ceil_div64(M, rpb); }
  dim3 grid((unsigned)cblocks, (unsigned)rblocks);
  SG_LAUNCH(colreduce_kernel<F>, grid, dim3(RED_TX, RED_TY), 0, st, f, M, C, rpb, sums, nout);
  return 0;
}

__global__ void zero_doubles(double* p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0;
}
__global__ void doubles_to_float(const double* p, float* o, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = (float)p[i];
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, int64_t count,
                                   int64_t unbias_mult, int64_t C, const float* gamma,
                                   const float* beta, float eps, float momentum, int training,
                                   float* running_mean, float* running_var, float* scale,
                                   float* shift, float* save, long long* num_batches_tracked) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && num_batches_tracked) *num_batches_tracked += 1;   // nn.BatchNorm's counter
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    double m = sums[c] / (double)count;
    double var = sums[C + c] / (double)count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      double n = (double)count * (double)unbias_mult;
      double unb = n > 1.0 ? var * n / (n - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c];
    invstd = 1.f / sqrtf(running_var[c] + eps);
  }
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - mean * g * invstd;
  save[c] = mean;
  save[C + c] = invstd;
}

// ---- y[n,Y,X,coff+c] = leaky(x[n,Y/up,X/up,c]*scale+shift) -------------------
// I = index type: uint32_t when every offset fits (64-bit integer division costs
// >100 instructions per element and made these passes compute-bound)
template <int VEC, typename I>
__global__ void scale_act_fwd_kernel(const float* __restrict__ x, I N, I H, I W,
                                     I C, const float* __restrict__ scale,
                                     const float* __restrict__ shift, float slope, int up,
                                     float* __restrict__ y, I ycs, I yco, int rnd) {
  I cg = C / VEC;
  I Ho = H * up, Wo = W * up;
  I i = (I)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Ho * Wo * cg) return;
  I c = (i % cg) * VEC;
  I pix = i / cg;
  I X = pix % Wo;
  I t = pix / Wo;
  I Y = t % Ho;
  I n = t / Ho;
  const float* xp = x + ((n * H + Y / up) * W + X / up) * C + c;
  float* yp = y + pix * ycs + yco + c;
  if (VEC == 4) {
    float4 v = *reinterpret_cast<const float4*>(xp);
    if (scale) {
      float4 s = *reinterpret_cast<const float4*>(scale + c);
      float4 b = *reinterpret_cast<const float4*>(shift + c);
      v.x = fmaf(v.x, s.x, b.x); v.y = fmaf(v.y, s.y, b.y);
      v.z = fmaf(v.z, s.z, b.z); v.w = fmaf(v.w, s.w, b.w);
    }
    v.x = leaky(v.x, slope); v.y = leaky(v.y, slope);
    v.z = leaky(v.z, slope); v.w = leaky(v.w, slope);
    if (rnd) { v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w); }
    *reinterpret_cast<float4*>(yp) = v;
  } else {
    float v = xp[0];
    if (scale) v = fmaf(v, scale[c], shift[c]);
    v = leaky(v, slope);
    yp[0] = rnd ? tf32_rn(v) : v;
  }
}

// g = leaky'(pre) * sum_{up x up} dy   for input element (n,y,x,c)
struct ActGrad {
  const float* dy; int64_t dcs, dco;
  const float* x; int64_t H, W, C;
  const float* scale; const float* shift;
  float slope; int up;
  __device__ __forceinline__ float at(int64_t m, int64_t c, float& xv) const {
    xv = x[m * C + c];
    float pre = scale ? fmaf(xv, scale[c], shift[c]) : xv;
    float d = pre > 0.f ? 1.f : slope;
    float g;
    if (up == 1) {
      g = dy[m * dcs + dco + c];
    } else {
      int64_t xx = m % W; int64_t t = m / W; int64_t yy = t % H; int64_t n = t / H;
      int64_t Wo = W * up;
      g = 0.f;
      for (int a = 0; a < up; ++a)
        for (int b = 0; b < up; ++b)
          g += dy[((n * H * up + yy * up + a) * Wo + xx * up + b) * dcs + dco + c];
    }
    return g * d;
  }
};

// float4 form of ActGrad::at (C, dcs, dco multiples of 4; 16-byte aligned bases)
__device__ __forceinline__ float4 actgrad4(const ActGrad& a, int64_t m, int64_t c, float4& xv) {
  xv = *reinterpret_cast<const float4*>(a.x + m * a.C + c);
  float4 pre = xv;
  if (a.scale) {
    float4 s = *reinterpret_cast<const float4*>(a.scale + c);
    float4 b = *reinterpret_cast<const float4*>(a.shift + c);
    pre.x = fmaf(xv.x, s.x, b.x); pre.y = fmaf(xv.y, s.y, b.y);
    pre.z = fmaf(xv.z, s.z, b.z); pre.w = fmaf(xv.w, s.w, b.w);
  }
  float4 g;
  if (a.up == 1) {
    g = *reinterpret_cast<const float4*>(a.dy + m * a.dcs + a.dco + c);
  } else {
    // rows < 2^31 is guaranteed by the host for the float4 path: 32-bit division
    uint32_t mm = (uint32_t)m, Wd = (uint32_t)a.W, Hd = (uint32_t)a.H;
    uint32_t xx = mm % Wd; uint32_t t = mm / Wd; uint32_t yy = t % Hd; uint32_t n = t / Hd;
    int64_t Wo = a.W * a.up;
    g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < a.up; ++i)
      for (int j = 0; j < a.up; ++j) {
        float4 d = *reinterpret_cast<const float4*>(
            a.dy + (((int64_t)n * a.H * a.up + yy * a.up + i) * Wo + xx * a.up + j) * a.dcs + a.dco + c);
        g.x += d.x; g.y += d.y; g.z += d.z; g.w += d.w;
      }
  }
  g.x *= pre.x > 0.f ? 1.f : a.slope; g.y *= pre.y > 0.f ? 1.f : a.slope;
  g.z *= pre.z > 0.f ? 1.f : a.slope; g.w *= pre.w > 0.f ? 1.f : a.slope;
  return g;
}

struct BwdReduceF {
  ActGrad ag; const float* save; int64_t C;
  __device__ void operator()(int64_t m, int64_t c, float& v0, float& v1) const {
    float xv;
    float g = ag.at(m, c, xv);
    float xhat = save ? (xv - save[c]) * save[C + c] : xv;
    v0 = g; v1 = g * xhat;
  }
  __device__ __forceinline__ void at4(int64_t m, int64_t c, float4& v0, float4& v1) const {
    float4 xv;
    float4 g = actgrad4(ag, m, c, xv);
    float4 xh = xv;
    if (save) {
      float4 mu = *reinterpret_cast<const float4*>(save + c);
      float4 is = *reinterpret_cast<const float4*>(save + C + c);
      xh.x = (xv.x - mu.x) * is.x; xh.y = (xv.y - mu.y) * is.y;
      xh.z = (xv.z - mu.z) * is.z; xh.w = (xv.w - mu.w) * is.w;
    }
    v0 = g; v1 = make_float4(g.x * xh.x, g.y * xh.y, g.z * xh.z, g.w * xh.w);
  }
};

__global__ void scale_act_bwd_apply4_kernel(ActGrad ag, const float* __restrict__ save,
                                            int64_t M, int64_t C, int training,
                                            const double* __restrict__ sums,
                                            float* __restrict__ dx) {
  // host guarantees M * C < 2^31 on this path
  uint32_t cg = (uint32_t)(C / 4);
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint32_t)M * cg) return;
  uint32_t mrow = i / cg;
  int64_t m = mrow, c = (int64_t)(i - mrow * cg) * 4;
  float4 xv;
  float4 g = actgrad4(ag, m, c, xv);
  float4 sc = ag.scale ? *reinterpret_cast<const float4*>(ag.scale + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  float gv[4] = {g.x, g.y, g.z, g.w}, xs[4] = {xv.x, xv.y, xv.z, xv.w}, ss[4] = {sc.x, sc.y, sc.z, sc.w};
  float r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (training && save) {
      float xhat = (xs[j] - save[c + j]) * save[C + c + j];
      float mg = (float)(sums[c + j] / (double)M);
      float mgx = (float)(sums[C + c + j] / (double)M);
      r[j] = ss[j] * (gv[j] - mg - xhat * mgx);
    } else {
      r[j] = ss[j] * gv[j];
    }
  }
  *reinterpret_cast<float4*>(dx + m * C + c) = make_float4(r[0], r[1], r[2], r[3]);
}

__global__ void scale_act_bwd_apply_kernel(ActGrad ag, const float* __restrict__ save,
                                           int64_t M, int64_t C, int training,
                                           const double* __restrict__ sums,
                                           float* __restrict__ dx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  int64_t m = i / C, c = i - m * C;
  float xv;
  float g = ag.at(m, c, xv);
  float sc = ag.scale ? ag.scale[c] : 1.f;
  float r;
  if (training && save) {
    float xhat = (xv - save[c]) * save[C + c];
    float mg = (float)(sums[c] / (double)M);
    float mgx = (float)(sums[C + c] / (double)M);
    r = sc * (g - mg - xhat * mgx);
  } else {
    r = sc * g;
  }
  dx[i] = r;
}

__global__ void bn_param_grads(const double* __restrict__ sums, int64_t C, float* dgamma,
                               float* dbeta, int accumulate) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // accumulate: dgamma / dbeta are the parameters' slots of a gradient bucket (zeroed at the start
  // of the step; a layer that runs twice per backward, like the discriminators', adds twice)
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)sums[c];
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)sums[C + c];
}

template <int VEC>
__global__ void avgpool2_fwd_kernel(const float* __restrict__ x, int64_t xcs, int64_t xco,
                                    int64_t N, int64_t H, int64_t W, int64_t C,
                                    float* __restrict__ y, int64_t ycs, int64_t yco) {
  // element counts < 2^31 (checked on the host): 32-bit index splitting
  uint32_t cg = (uint32_t)(C / VEC), Ho = (uint32_t)(H / 2), Wo = (uint32_t)(W / 2);
  uint32_t ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= (uint32_t)N * Ho * Wo * cg) return;
  int64_t c = (int64_t)(ii % cg) * VEC;
  uint32_t pixu = ii / cg;
  int64_t pix = pixu;
  int64_t X = pixu % Wo; uint32_t tu = pixu / Wo; int64_t Y = tu % Ho; int64_t n = tu / Ho;
  const float* p00 = x + ((n * H + 2 * Y) * W + 2 * X) * xcs + xco + c;
  const float* p10 = p00 + W * xcs;
  float* yp = y + pix * ycs + yco + c;
  if (VEC == 4) {
    float4 a = *reinterpret_cast<const float4*>(p00);
    float4 b = *reinterpret_cast<const float4*>(p00 + xcs);
    float4 d = *reinterpret_cast<const float4*>(p10);
    float4 e = *reinterpret_cast<const float4*>(p10 + xcs);
    float4 r;
    r.x = (a.x + b.x + d.x + e.x) * 0.25f; r.y = (a.y + b.y + d.y + e.y) * 0.25f;
    r.z = (a.z + b.z + d.z + e.z) * 0.25f; r.w = (a.w + b.w + d.w + e.w) * 0.25f;
    *reinterpret_cast<float4*>(yp) = r;
  } else {
    yp[0] = (p00[0] + p00[xcs] + p10[0] + p10[xcs]) * 0.25f;
  }
}

template <int VEC>
__global__ void avgpool2_bwd_kernel(const float* __restrict__ dc, int64_t dcs, int64_t dco,
                                    int64_t N, int64_t H, int64_t W, int64_t C,
                                    float* __restrict__ df, int64_t dfs, int64_t dfo, int acc) {
  // (H, W) are the FINE dims
  uint32_t cg = (uint32_t)(C / VEC);
  uint32_t ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= (uint32_t)N * (uint32_t)H * (uint32_t)W * cg) return;
  int64_t c = (int64_t)(ii % cg) * VEC;
  uint32_t pixu = ii / cg;
  int64_t pix = pixu;
  int64_t X = pixu % (uint32_t)W; uint32_t tu = pixu / (uint32_t)W; int64_t Y = tu % (uint32_t)H; int64_t n = tu / (uint32_t)H;
  const float* cp = dc + ((n * (H / 2) + Y / 2) * (W / 2) + X / 2) * dcs + dco + c;
  float* fp = df + pix * dfs + dfo + c;
  if (VEC == 4) {
    float4 g = *reinterpret_cast<const float4*>(cp);
    float4 r = make_float4(g.x * 0.25f, g.y * 0.25f, g.z * 0.25f, g.w * 0.25f);
    if (acc) {
      float4 o = *reinterpret_cast<const float4*>(fp);
      r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
    }
    *reinterpret_cast<float4*>(fp) = r;
  } else {
    float r = cp[0] * 0.25f;
    fp[0] = acc ? fp[0] + r : r;
  }
}

__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                               float slope, int64_t n, float* __restrict__ dx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = y[i] > 0.f ? dy[i] : dy[i] * slope;
}

}  // namespace

extern "C" int sg2im_bn_stats(const float* x, int64_t M, int64_t C, double* sums,
                              sg2im_stream_t stream) {
  SG_ARG(x && sums && M >= 1 && C >= 1);
  StatsF f{x, C};
  if (C % 4 == 0 && aligned16(x)) launch_colreduce4(f, M, C, sums, 2, as_stream(stream));
  else launch_colreduce(f, M, C, sums, 2, as_stream(stream));
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_colsum(const float* x, int64_t M, int64_t C, float* out, double* scratch,
                            sg2im_stream_t stream) {
  SG_ARG(x && out && scratch && M >= 1 && C >= 1);
  cudaStream_t st = as_stream(stream);
  {
    const char* e = getenv("SG2IM_COLSUM_V2");            // read per call: tests toggle it in-process
    // one CTA per 32 columns walks all M rows: right for the small GEMMs (a few hundred rows);
    // taller inputs (6272 x 256 of the image discriminator: 61 us on 8 CTAs, measured) go to the
    // split reduction below
    if (!(e && e[0] == '0') && M <= 1024 && M * C < (1ll << 31)) {
      sg2im_colsum_small(x, M, C, out, st);
      SG_LAUNCH_OK();
      return 0;
    }
  }
  SG_LAUNCH(zero_doubles, (unsigned)ceil_div64(C, 256), 256, 0, st, scratch, C);
  StatsF f{x, C};
  if (C % 4 == 0 && aligned16(x)) launch_colreduce4(f, M, C, scratch, 1, st);
  else launch_colreduce(f, M, C, scratch, 1, st);
  SG_LAUNCH(doubles_to_float, (unsigned)ceil_div64(C, 256), 256, 0, st, scratch, out, C);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_bn_finalize(const double* sums, int64_t count, int64_t unbias_mult, int64_t C,
                                 const float* gamma, const float* beta, float eps, float momentum,
                                 int training, float* running_mean, float* running_var,
                                 float* scale, float* shift, float* save, int64_t* num_batches_tracked, sg2im_stream_t stream) {
  SG_ARG(scale && shift && save && C >= 1 && count >= 1 && unbias_mult >= 1);
  SG_ARG(training ? sums != nullptr : (running_mean && running_var));
  SG_LAUNCH(bn_finalize_kernel, (unsigned)ceil_div64(C, 128), 128, 0, as_stream(stream), 
      sums, count, unbias_mult, C, gamma, beta, eps, momentum, training, running_mean,
      running_var, scale, shift, save, reinterpret_cast<long long*>(num_batches_tracked));
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_scale_act_fwd(const float* x, int64_t N, int64_t H, int64_t W, int64_t C,
                                   const float* scale, const float* shift, float slope, int up,
                                   float* y, int64_t y_cstride, int64_t y_coff, int round_tf32,
                                   sg2im_stream_t stream) {
  SG_ARG(x && y && N >= 1 && H >= 1 && W >= 1 && C >= 1 && up >= 1);
  SG_ARG((scale == nullptr) == (shift == nullptr));
  SG_ARG(y_coff >= 0 && y_cstride >= y_coff + C);
  bool vec = (C % 4 == 0) && (y_cstride % 4 == 0) && (y_coff % 4 == 0) && aligned16(x) &&
             aligned16(y) && (!scale || (aligned16(scale) && aligned16(shift)));
  int64_t total = N * H * up * W * up * (C / (vec ? 4 : 1));
  unsigned grid = (unsigned)ceil_div64(total, 256);
  cudaStream_t st = as_stream(stream);
  bool small = N * H * up * W * up * y_cstride < (1ll << 31) && N * H * W * C < (1ll << 31);
  typedef uint32_t U;
  if (vec && small && (up == 1 || up == 2) && bn_fwd_v2_enabled())
    sg2im_scale_act_fwd_v2(x, N, H, W, C, scale, shift, slope, up, y, y_cstride, y_coff, round_tf32,
                           st);
  else if (vec && small)
    SG_LAUNCH((scale_act_fwd_kernel<4, U>), grid, 256, 0, st, x, (U)N, (U)H, (U)W, (U)C, scale, shift, slope, up, y, (U)y_cstride, (U)y_coff, round_tf32);
  else if (vec)
    SG_LAUNCH((scale_act_fwd_kernel<4, int64_t>), grid, 256, 0, st, x, N, H, W, C, scale, shift, slope, up, y, y_cstride, y_coff, round_tf32);
  else
    SG_LAUNCH((scale_act_fwd_kernel<1, int64_t>), grid, 256, 0, st, x, N, H, W, C, scale, shift, slope, up, y, y_cstride, y_coff, round_tf32);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_scale_act_bwd_reduce(const float* dy, int64_t dy_cstride, int64_t dy_coff,
                                          const float* x, int64_t N, int64_t H, int64_t W,
                                          int64_t C, const float* scale, const float* shift,
                                          const float* save, float slope, int up, double* sums,
                                          sg2im_stream_t stream) {
  SG_ARG(dy && x && sums && N >= 1 && H >= 1 && W >= 1 && C >= 1 && up >= 1);
  BwdReduceF f{{dy, dy_cstride, dy_coff, x, H, W, C, scale, shift, slope, up}, save, C};
  bool vec = (C % 4 == 0) && (dy_cstride % 4 == 0) && (dy_coff % 4 == 0) && aligned16(dy) &&
             aligned16(x) && (!scale || (aligned16(scale) && aligned16(shift))) &&
             (!save || aligned16(save)) && N * H * W * C < (1ll << 31);
  if (vec && scale && shift && save && (up == 1 || up == 2) && bn_bwd_v2_enabled())
    sg2im_bn_bwd_reduce_v2(dy, dy_cstride, dy_coff, x, N, H, W, C, scale, shift, save, slope, up,
                           sums, as_stream(stream));
  else if (vec) launch_colreduce4(f, N * H * W, C, sums, 2, as_stream(stream));
  else launch_colreduce(f, N * H * W, C, sums, 2, as_stream(stream));
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_scale_act_bwd_apply(const float* dy, int64_t dy_cstride, int64_t dy_coff,
                                         const float* x, int64_t N, int64_t H, int64_t W,
                                         int64_t C, const float* scale, const float* shift,
                                         const float* save, float slope, int up, int training,
                                         const double* sums, float* dx, float* dgamma,
                                         float* dbeta, sg2im_stream_t stream) {
  SG_ARG(dy && x && dx && N >= 1 && H >= 1 && W >= 1 && C >= 1 && up >= 1);
  const int accumulate = (training >> 1) & 1;       // bit 1: ADD the parameter gradients into dgamma / dbeta
  training &= 1;
  SG_ARG(!training || !save || sums);
  cudaStream_t st = as_stream(stream);
  ActGrad ag{dy, dy_cstride, dy_coff, x, H, W, C, scale, shift, slope, up};
  int64_t M = N * H * W;
  bool vec = (C % 4 == 0) && (dy_cstride % 4 == 0) && (dy_coff % 4 == 0) && aligned16(dy) &&
             aligned16(x) && aligned16(dx) && (!scale || (aligned16(scale) && aligned16(shift))) &&
             M * C < (1ll << 31);
  if (vec && training && scale && shift && save && aligned16(save) && sums &&
      (up == 1 || up == 2) && bn_bwd_v2_enabled())
    sg2im_bn_bwd_apply_v2(dy, dy_cstride, dy_coff, x, N, H, W, C, scale, shift, save, slope, up,
                          sums, dx, st);
  else if (vec)
    SG_LAUNCH(scale_act_bwd_apply4_kernel, (unsigned)ceil_div64(M * (C / 4), 256), 256, 0, st, 
        ag, save, M, C, training, sums, dx);
  else
    SG_LAUNCH(scale_act_bwd_apply_kernel, (unsigned)ceil_div64(M * C, 256), 256, 0, st, 
        ag, save, M, C, training, sums, dx);
  if ((dgamma || dbeta) && sums)
    SG_LAUNCH(bn_param_grads, (unsigned)ceil_div64(C, 128), 128, 0, st, sums, C, dgamma, dbeta, accumulate);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_avgpool2_fwd(const float* x, int64_t x_cstride, int64_t x_coff,
                                  int64_t N, int64_t H, int64_t W, int64_t C,
                                  float* y, int64_t y_cstride, int64_t y_coff,
                                  sg2im_stream_t stream) {
  SG_ARG(x && y && N >= 1 && C >= 1 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0);
  SG_ARG(N * H * W * C < (1ll << 31));
  bool vec = (C % 4 == 0) && (x_cstride % 4 == 0) && (x_coff % 4 == 0) && (y_cstride % 4 == 0) &&
             (y_coff % 4 == 0) && aligned16(x) && aligned16(y);
  int64_t total = N * (H / 2) * (W / 2) * (C / (vec ? 4 : 1));
  unsigned grid = (unsigned)ceil_div64(total, 256);
  cudaStream_t st = as_stream(stream);
  if (vec) SG_LAUNCH(avgpool2_fwd_kernel<4>, grid, 256, 0, st, x, x_cstride, x_coff, N, H, W, C, y, y_cstride, y_coff);
  else     SG_LAUNCH(avgpool2_fwd_kernel<1>, grid, 256, 0, st, x, x_cstride, x_coff, N, H, W, C, y, y_cstride, y_coff);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_avgpool2_bwd(const float* dcoarse, int64_t dc_cstride, int64_t dc_coff,
                                  int64_t N, int64_t H, int64_t W, int64_t C,
                                  float* dfine, int64_t df_cstride, int64_t df_coff,
                                  int accumulate, sg2im_stream_t stream) {
  SG_ARG(dcoarse && dfine && N >= 1 && C >= 1 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0);
  SG_ARG(N * H * W * C < (1ll << 31));
  bool vec = (C % 4 == 0) && (dc_cstride % 4 == 0) && (dc_coff % 4 == 0) &&
             (df_cstride % 4 == 0) && (df_coff % 4 == 0) && aligned16(dcoarse) && aligned16(dfine);
  int64_t total = N * H * W * (C / (vec ? 4 : 1));
  unsigned grid = (unsigned)ceil_div64(total, 256);
  cudaStream_t st = as_stream(stream);
  if (vec) SG_LAUNCH(avgpool2_bwd_kernel<4>, grid, 256, 0, st, dcoarse, dc_cstride, dc_coff, N, H, W, C, dfine, df_cstride, df_coff, accumulate);
  else     SG_LAUNCH(avgpool2_bwd_kernel<1>, grid, 256, 0, st, dcoarse, dc_cstride, dc_coff, N, H, W, C, dfine, df_cstride, df_coff, accumulate);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_act_bwd(const float* dy, const float* y, float slope, int64_t n, float* dx,
                             sg2im_stream_t stream) {
  SG_ARG(dy && y && dx && n >= 0);
  if (n == 0) return 0;
  SG_LAUNCH(act_bwd_kernel, (unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream), dy, y, slope, n, dx);
  SG_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------- space to depth ---
namespace {
__global__ void s2d_fwd_kernel(const float* __restrict__ x, int64_t sxn, int64_t sxh, int64_t sxw,
                               int64_t sxc, int64_t N, int64_t H, int64_t W, int64_t C,
                               float* __restrict__ out) {
  int64_t H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H2 * W2 * 4 * C) return;
  int64_t c = i % C; int64_t t = i / C;
  int ph = (int)(t % 4); t /= 4;
  int64_t x2 = t % W2; t /= W2;
  int64_t y2 = t % H2; int64_t n = t / H2;
  int64_t y = 2 * y2 + (ph >> 1), xx = 2 * x2 + (ph & 1);
  out[i] = (y < H && xx < W) ? x[n * sxn + y * sxh + xx * sxw + c * sxc] : 0.f;
}
// float4 variants (C % 4 == 0, unit channel stride)
__global__ void s2d_fwd4_kernel(const float* __restrict__ x, int64_t sxn, int64_t sxh, int64_t sxw,
                                int64_t N, int64_t H, int64_t W, int64_t C, float* __restrict__ out) {
  int64_t H2 = (H + 1) / 2, W2 = (W + 1) / 2, cg = C / 4;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H2 * W2 * 4 * cg) return;
  int64_t c = (i % cg) * 4; int64_t t = i / cg;
  int ph = (int)(t % 4); t /= 4;
  int64_t x2 = t % W2; t /= W2;
  int64_t y2 = t % H2; int64_t n = t / H2;
  int64_t y = 2 * y2 + (ph >> 1), xx = 2 * x2 + (ph & 1);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (y < H && xx < W) v = *reinterpret_cast<const float4*>(x + n * sxn + y * sxh + xx * sxw + c);
  *reinterpret_cast<float4*>(out + (((n * H2 + y2) * W2 + x2) * 4 + ph) * C + c) = v;
}
__global__ void s2d_bwd4_kernel(const float* __restrict__ dout, int64_t N, int64_t H, int64_t W,
                                int64_t C, float* __restrict__ dx) {
  int64_t H2 = (H + 1) / 2, W2 = (W + 1) / 2, cg = C / 4;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H * W * cg) return;
  int64_t c = (i % cg) * 4; int64_t t = i / cg;
  int64_t xx = t % W; t /= W;
  int64_t y = t % H; int64_t n = t / H;
  int ph = (int)((y & 1) * 2 + (xx & 1));
  *reinterpret_cast<float4*>(dx + ((n * H + y) * W + xx) * C + c) =
      *reinterpret_cast<const float4*>(dout + (((n * H2 + y / 2) * W2 + xx / 2) * 4 + ph) * C + c);
}
__global__ void s2d_bwd_kernel(const float* __restrict__ dout, int64_t N, int64_t H, int64_t W,
                               int64_t C, float* __restrict__ dx) {
  int64_t H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H * W * C) return;
  int64_t c = i % C; int64_t t = i / C;
  int64_t xx = t % W; t /= W;
  int64_t y = t % H; int64_t n = t / H;
  int ph = (int)((y & 1) * 2 + (xx & 1));
  dx[i] = dout[(((n * H2 + y / 2) * W2 + xx / 2) * 4 + ph) * C + c];
}
}  // namespace

extern "C" int sg2im_s2d_fwd(const float* x, int64_t sxn, int64_t sxh, int64_t sxw, int64_t sxc,
                             int64_t N, int64_t H, int64_t W, int64_t C, float* out,
                             sg2im_stream_t stream) {
  SG_ARG(x && out && N >= 1 && H >= 1 && W >= 1 && C >= 1);
  int64_t total = N * ((H + 1) / 2) * ((W + 1) / 2) * 4 * C;
  if (sxc == 1 && C % 4 == 0 && sxn % 4 == 0 && sxh % 4 == 0 && sxw % 4 == 0 && aligned16(x) &&
      aligned16(out))
    SG_LAUNCH(s2d_fwd4_kernel, (unsigned)ceil_div64(total / 4, 256), 256, 0, as_stream(stream), 
        x, sxn, sxh, sxw, N, H, W, C, out);
  else
    SG_LAUNCH(s2d_fwd_kernel, (unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream), 
        x, sxn, sxh, sxw, sxc, N, H, W, C, out);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_s2d_bwd(const float* dout, int64_t N, int64_t H, int64_t W, int64_t C,
                             float* dx, sg2im_stream_t stream) {
  SG_ARG(dout && dx && N >= 1 && H >= 1 && W >= 1 && C >= 1);
  int64_t total = N * H * W * C;
  if (C % 4 == 0 && aligned16(dout) && aligned16(dx))
    SG_LAUNCH(s2d_bwd4_kernel, (unsigned)ceil_div64(total / 4, 256), 256, 0, as_stream(stream), dout, N, H, W, C, dx);
  else
    SG_LAUNCH(s2d_bwd_kernel, (unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream), dout, N, H, W, C, dx);
  SG_LAUNCH_OK();
  return 0;
}
