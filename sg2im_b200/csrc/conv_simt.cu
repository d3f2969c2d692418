// Exact-fp32 implicit-GEMM convolution on the FFMA pipe (NHWC):
//   conv_igemm  — forward and data-gradient (mode 0 / 1)
//   conv_wgrad  — weight gradient, pixel reduction split over CTAs
//   conv_skinny — Cout <= 4 outputs (RGB head, mask head): one thread per pixel
// This is the reference-precision path (bit-for-bit fp32 products, fp32
// accumulate) used for parity and for the shapes the tcgen05 kernel
// (conv_tc.cu) does not take.  Replaces nn.Conv2d / nn.Linear of
// sg2im/crn.py:41-45,80-82, model.py:100,105, layers.py:178,221.
#include "common.cuh"

namespace {

struct ConvP {
  const float* x; int64_t sxn, sxh, sxw, sxc;
  int64_t N, Hin, Win, Cin;
  const float* w; const float* bias;
  int KH, KW, S, P;
  int64_t Hout, Wout, Cout;
  int act; float slope;
  float* y; int64_t y_cstride, y_coff;
  int64_t M;                 // N*Hout*Wout
  int vecA, vecB, vecY;
};

constexpr int BM = 128, BN = 128, BK = 8, AS_LD = BM + 4;

// map an output pixel + tap to the input pixel it reads; returns validity
template <int MODE>
__device__ __forceinline__ bool tap_src(const ConvP& p, int oy, int ox, int ky, int kx,
                                        int& iy, int& ix) {
  if (MODE == 0) {
    iy = oy * p.S - p.P + ky;
    ix = ox * p.S - p.P + kx;
    return iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
  } else {
    int ty = oy + p.P - ky, tx = ox + p.P - kx;
    if (ty < 0 || tx < 0) return false;
    if (p.S > 1) {
      if ((ty % p.S) | (tx % p.S)) return false;
      ty /= p.S; tx /= p.S;
    }
    iy = ty; ix = tx;
    return iy < p.Hin && ix < p.Win;
  }
}

template <int MODE>
__global__ void __launch_bounds__(256, 2) conv_igemm_kernel(ConvP p) {
  __shared__ __align__(16) float As[2][BK][AS_LD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;

  // --- A loader: row a_row, 4 consecutive k starting at a_k
  const int a_row = tid >> 1, a_k = (tid & 1) * 4;
  const int64_t am = m0 + a_row;
  const bool a_ok = am < p.M;
  int a_n = 0, a_oy = 0, a_ox = 0;
  if (a_ok) {
    int64_t hw = p.Hout * p.Wout;
    a_n = (int)(am / hw);
    int64_t r = am - (int64_t)a_n * hw;
    a_oy = (int)(r / p.Wout);
    a_ox = (int)(r - (int64_t)a_oy * p.Wout);
  }
  // --- B loader: k row b_k, 4 consecutive columns at b_col
  const int b_k = tid >> 5, b_col = (tid & 31) * 4;

  const int cpt = (int)((p.Cin + BK - 1) / BK);         // K chunks per tap
  const int taps = p.KH * p.KW;
  const int iters = taps * cpt;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra, rb;
  auto load_tile = [&](int it) {
    int tap = it / cpt;
    int c0 = (it - tap * cpt) * BK;
    int ky = tap / p.KW, kx = tap - ky * p.KW;
    // A
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    int iy, ix;
    if (a_ok && tap_src<MODE>(p, a_oy, a_ox, ky, kx, iy, ix)) {
      int ci = c0 + a_k;
      const float* src = p.x + (int64_t)a_n * p.sxn + (int64_t)iy * p.sxh + (int64_t)ix * p.sxw;
      if (p.vecA && ci + 3 < p.Cin) {
        ra = *reinterpret_cast<const float4*>(src + ci);
      } else {
        if (ci + 0 < p.Cin) ra.x = src[(int64_t)(ci + 0) * p.sxc];
        if (ci + 1 < p.Cin) ra.y = src[(int64_t)(ci + 1) * p.sxc];
        if (ci + 2 < p.Cin) ra.z = src[(int64_t)(ci + 2) * p.sxc];
        if (ci + 3 < p.Cin) ra.w = src[(int64_t)(ci + 3) * p.sxc];
      }
    }
    // B
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    int ck = c0 + b_k;
    if (ck < p.Cin) {
      const float* wr = p.w + ((int64_t)tap * p.Cin + ck) * p.Cout;
      int64_t col = n0 + b_col;
      if (p.vecB && col + 3 < p.Cout) {
        rb = *reinterpret_cast<const float4*>(wr + col);
      } else {
        if (col + 0 < p.Cout) rb.x = wr[col + 0];
        if (col + 1 < p.Cout) rb.y = wr[col + 1];
        if (col + 2 < p.Cout) rb.z = wr[col + 2];
        if (col + 3 < p.Cout) rb.w = wr[col + 3];
      }
    }
  };
  auto store_tile = [&](int buf) {
    As[buf][a_k + 0][a_row] = ra.x;
    As[buf][a_k + 1][a_row] = ra.y;
    As[buf][a_k + 2][a_row] = ra.z;
    As[buf][a_k + 3][a_row] = ra.w;
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_col]) = rb;
  };

  load_tile(0);
  store_tile(0);
  __syncthreads();

  for (int it = 0; it < iters; ++it) {
    int buf = it & 1;
    if (it + 1 < iters) load_tile(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < iters) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // --- epilogue: bias + activation, write the channel slice
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    float* yrow = p.y + m * p.y_cstride + p.y_coff;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int64_t col = n0 + (h ? 64 + tx * 4 : tx * 4);
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = acc[i][h * 4 + j];
        if (p.bias && col + j < p.Cout) t += p.bias[col + j];
        if (p.act) t = leaky(t, p.slope);
        v[j] = t;
      }
      if (p.vecY && col + 3 < p.Cout) {
        *reinterpret_cast<float4*>(yrow + col) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < p.Cout) yrow[col + j] = v[j];
      }
    }
  }
}

// ------------------------------------------------------------ skinny rows ---
// Cout <= 4, 1x1, few rows (the discriminator heads: 320 rows x 1024 features -> 1): a WARP per row,
// lanes stride over the features in float4 steps, shuffle reduction.  The thread-per-row kernel
// below walks 1024 features serially on 320 threads (66 us measured for a 1.3 MB read).
__global__ void __launch_bounds__(256) conv_skinny_rows_kernel(ConvP p) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= p.M) return;                                   // whole warps leave together
  const int64_t hw = p.Hout * p.Wout;
  const int n = (int)(m / hw);
  const int64_t r = m - (int64_t)n * hw;
  const int oy = (int)(r / p.Wout), ox = (int)(r - (int64_t)oy * p.Wout);
  const float* src = p.x + (int64_t)n * p.sxn + (int64_t)oy * p.sxh + (int64_t)ox * p.sxw;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ci = lane * 4; ci + 3 < p.Cin; ci += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(src + ci);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      for (int j = 0; j < p.Cout; ++j) acc[j] = fmaf(xs[q], p.w[(int64_t)(ci + q) * p.Cout + j], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  if (lane == 0) {
    float* yrow = p.y + m * p.y_cstride + p.y_coff;
    for (int j = 0; j < p.Cout; ++j) {
      float t = acc[j] + (p.bias ? p.bias[j] : 0.f);
      if (p.act) t = leaky(t, p.slope);
      yrow[j] = t;
    }
  }
}

// ---------------------------------------------------------------- skinny ---
// Cout <= 4 forward: thread per output pixel, weights (K x 4, zero padded) in
// shared memory.  Memory-bound on x.
__global__ void __launch_bounds__(256) conv_skinny_kernel(ConvP p) {
  SG_DYN_SMEM(float, ws);            // [K][4]
  const int64_t K = (int64_t)p.KH * p.KW * p.Cin;
  for (int64_t i = threadIdx.x; i < K * 4; i += blockDim.x) {
    int64_t k = i >> 2; int j = (int)(i & 3);
    ws[i] = j < p.Cout ? p.w[k * p.Cout + j] : 0.f;
  }
  __syncthreads();
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  int64_t hw = p.Hout * p.Wout;
  int n = (int)(m / hw);
  int64_t r = m - (int64_t)n * hw;
  int oy = (int)(r / p.Wout), ox = (int)(r - (int64_t)oy * p.Wout);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ky = 0; ky < p.KH; ++ky) {
    for (int kx = 0; kx < p.KW; ++kx) {
      int iy, ix;
      if (!tap_src<0>(p, oy, ox, ky, kx, iy, ix)) continue;
      const float* src = p.x + (int64_t)n * p.sxn + (int64_t)iy * p.sxh + (int64_t)ix * p.sxw;
      const float4* wt = reinterpret_cast<const float4*>(ws) + (int64_t)(ky * p.KW + kx) * p.Cin;
      int ci = 0;
      if (p.vecA) {
        for (; ci + 3 < p.Cin; ci += 4) {
          float4 xv = *reinterpret_cast<const float4*>(src + ci);
          float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 wv = wt[ci + q];
            acc[0] = fmaf(xs[q], wv.x, acc[0]); acc[1] = fmaf(xs[q], wv.y, acc[1]);
            acc[2] = fmaf(xs[q], wv.z, acc[2]); acc[3] = fmaf(xs[q], wv.w, acc[3]);
          }
        }
      }
      for (; ci < p.Cin; ++ci) {
        float xv = src[(int64_t)ci * p.sxc];
        float4 wv = wt[ci];
        acc[0] = fmaf(xv, wv.x, acc[0]); acc[1] = fmaf(xv, wv.y, acc[1]);
        acc[2] = fmaf(xv, wv.z, acc[2]); acc[3] = fmaf(xv, wv.w, acc[3]);
      }
    }
  }
  float* yrow = p.y + m * p.y_cstride + p.y_coff;
  for (int j = 0; j < p.Cout; ++j) {
    float t = acc[j] + (p.bias ? p.bias[j] : 0.f);
    if (p.act) t = leaky(t, p.slope);
    yrow[j] = t;
  }
}

// ----------------------------------------------------------------- wgrad ---
struct WgradP {
  const float* x; int64_t sxn, sxh, sxw, sxc;
  int64_t N, Hin, Win, Cin;
  const float* dy;
  int KH, KW, S, P;
  int64_t Hout, Wout, Cout;
  float* dw;
  int64_t M, m_per_split;
  int ci_tiles;
  int vecX, vecY;
};

__global__ void __launch_bounds__(256, 2) conv_wgrad_kernel(WgradP p) {
  __shared__ __align__(16) float Xs[2][BK][BM];          // [pixel][ci]
  __shared__ __align__(16) float Ys[2][BK][BN];          // [pixel][co]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int tap = blockIdx.x / p.ci_tiles;
  const int64_t ci0 = (int64_t)(blockIdx.x - tap * p.ci_tiles) * BM;
  const int64_t co0 = (int64_t)blockIdx.y * BN;
  const int ky = tap / p.KW, kx = tap - ky * p.KW;
  const int64_t m_begin = (int64_t)blockIdx.z * p.m_per_split;
  const int64_t m_end = m_begin + p.m_per_split < p.M ? m_begin + p.m_per_split : p.M;
  if (m_begin >= m_end) return;
  const int iters = (int)((m_end - m_begin + BK - 1) / BK);
  const int lp = tid >> 5, lc = (tid & 31) * 4;          // loader pixel / channel
  const int64_t hw = p.Hout * p.Wout;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 rx, ry;
  auto load_tile = [&](int it) {
    rx = make_float4(0.f, 0.f, 0.f, 0.f);
    ry = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t m = m_begin + (int64_t)it * BK + lp;
    if (m >= m_end) return;
    int n = (int)(m / hw);
    int64_t r = m - (int64_t)n * hw;
    int oy = (int)(r / p.Wout), ox = (int)(r - (int64_t)oy * p.Wout);
    int iy = oy * p.S - p.P + ky, ix = ox * p.S - p.P + kx;
    if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win) {
      const float* src = p.x + (int64_t)n * p.sxn + (int64_t)iy * p.sxh + (int64_t)ix * p.sxw;
      int64_t ci = ci0 + lc;
      if (p.vecX && ci + 3 < p.Cin) {
        rx = *reinterpret_cast<const float4*>(src + ci);
      } else {
        if (ci + 0 < p.Cin) rx.x = src[(ci + 0) * p.sxc];
        if (ci + 1 < p.Cin) rx.y = src[(ci + 1) * p.sxc];
        if (ci + 2 < p.Cin) rx.z = src[(ci + 2) * p.sxc];
        if (ci + 3 < p.Cin) rx.w = src[(ci + 3) * p.sxc];
      }
    }
    const float* dyr = p.dy + m * p.Cout;
    int64_t co = co0 + lc;
    if (p.vecY && co + 3 < p.Cout) {
      ry = *reinterpret_cast<const float4*>(dyr + co);
    } else {
      if (co + 0 < p.Cout) ry.x = dyr[co + 0];
      if (co + 1 < p.Cout) ry.y = dyr[co + 1];
      if (co + 2 < p.Cout) ry.z = dyr[co + 2];
      if (co + 3 < p.Cout) ry.w = dyr[co + 3];
    }
  };
  auto store_tile = [&](int buf) {
    *reinterpret_cast<float4*>(&Xs[buf][lp][lc]) = rx;
    *reinterpret_cast<float4*>(&Ys[buf][lp][lc]) = ry;
  };

  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    int buf = it & 1;
    if (it + 1 < iters) load_tile(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&Xs[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&Xs[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Ys[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Ys[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < iters) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t ci = ci0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (ci >= p.Cin) continue;
    float* row = p.dw + ((int64_t)tap * p.Cin + ci) * p.Cout;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int64_t co = co0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (co < p.Cout) atomicAdd(row + co, acc[i][j]);
    }
  }
}

// 1x1 weight gradient with Cout <= 4 (RGB head 64->3, mask head 128->1, box head):
// dw[ci][co] += sum_m x[m][ci] * dy[m][co].  HBM-bound on x: threads map to
// channels (coalesced rows), 4+ pixel lanes per block, block reduce, one atomic
// per (ci, co) per block.
__global__ void __launch_bounds__(256)
wgrad_skinny_kernel(const float* __restrict__ x, int64_t xs, int64_t M, int Cin,
                    const float* __restrict__ dy, int Cout, int64_t rows_per_block,
                    float* __restrict__ dw) {
  __shared__ float red[4][256];
  const int lanes = 256 / Cin > 0 ? 256 / Cin : 1;       // pixel lanes (Cin <= 256)
  const int ci = threadIdx.x % Cin, pl = threadIdx.x / Cin;
  int64_t mb = (int64_t)blockIdx.x * rows_per_block;
  int64_t me = mb + rows_per_block < M ? mb + rows_per_block : M;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    for (int64_t m = mb + pl; m < me; m += lanes) {
      float xv = x[m * xs + ci];
      const float* d = dy + m * Cout;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < Cout) acc[j] = fmaf(xv, d[j], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[j][threadIdx.x] = (pl < lanes) ? acc[j] : 0.f;
  __syncthreads();
  if (pl == 0) {
    for (int j = 0; j < Cout; ++j) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[j][l * Cin + ci];
      atomicAdd(dw + (int64_t)ci * Cout + j, s);
    }
  }
}

}  // namespace

extern "C" int sg2im_conv_igemm(int mode, const float* x, int64_t sxn, int64_t sxh, int64_t sxw,
                                int64_t sxc, int64_t N, int64_t Hin, int64_t Win, int64_t Cin,
                                const float* w, const float* bias, int KH, int KW, int S, int P,
                                int64_t Hout, int64_t Wout, int64_t Cout, int act, float slope,
                                float* y, int64_t y_cstride, int64_t y_coff,
                                sg2im_stream_t stream) {
  SG_ARG(mode == 0 || mode == 1);
  SG_ARG(x && w && y);
  SG_ARG(N >= 1 && Hin >= 1 && Win >= 1 && Cin >= 1 && Hout >= 1 && Wout >= 1 && Cout >= 1);
  SG_ARG(KH >= 1 && KW >= 1 && S >= 1 && P >= 0);
  SG_ARG(y_cstride >= y_coff + Cout && y_coff >= 0);
  SG_ARG(Hin < (1 << 30) && Win < (1 << 30) && Cin < (1 << 30) && N < (1 << 30));
  if (mode == 0) {
    SG_ARG((Hin + 2 * P - KH) / S + 1 == Hout && (Win + 2 * P - KW) / S + 1 == Wout);
  } else {
    SG_ARG((Hout + 2 * P - KH) / S + 1 == Hin && (Wout + 2 * P - KW) / S + 1 == Win);
  }
  ConvP p;
  p.x = x; p.sxn = sxn; p.sxh = sxh; p.sxw = sxw; p.sxc = sxc;
  p.N = N; p.Hin = Hin; p.Win = Win; p.Cin = Cin;
  p.w = w; p.bias = bias; p.KH = KH; p.KW = KW; p.S = S; p.P = P;
  p.Hout = Hout; p.Wout = Wout; p.Cout = Cout; p.act = act; p.slope = slope;
  p.y = y; p.y_cstride = y_cstride; p.y_coff = y_coff;
  p.M = N * Hout * Wout;
  p.vecA = (sxc == 1) && (Cin % 4 == 0) && (sxn % 4 == 0) && (sxh % 4 == 0) && (sxw % 4 == 0) &&
           aligned16(x);
  p.vecB = (Cout % 4 == 0) && aligned16(w);
  p.vecY = (Cout % 4 == 0) && (y_cstride % 4 == 0) && (y_coff % 4 == 0) && aligned16(y);
  cudaStream_t st = as_stream(stream);
  int64_t K = (int64_t)KH * KW * Cin;
  if (mode == 0 && Cout <= 4 && KH == 1 && KW == 1 && S == 1 && P == 0 && p.vecA && Cin % 4 == 0 &&
      Cin >= 256 && p.M <= 16384) {
    SG_LAUNCH(conv_skinny_rows_kernel, (unsigned)ceil_div64(p.M, 8), 256, 0, st, p);
  } else if (mode == 0 && Cout <= 4 && K * 16 <= 48 * 1024) {
    unsigned grid = (unsigned)ceil_div64(p.M, 256);
    SG_LAUNCH(conv_skinny_kernel, grid, 256, (size_t)(K * 16), st, p);
  } else {
    dim3 grid((unsigned)ceil_div64(p.M, BM), (unsigned)ceil_div64(Cout, BN));
    if (mode == 0) SG_LAUNCH(conv_igemm_kernel<0>, grid, 256, 0, st, p);
    else           SG_LAUNCH(conv_igemm_kernel<1>, grid, 256, 0, st, p);
  }
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_conv_wgrad(const float* x, int64_t sxn, int64_t sxh, int64_t sxw, int64_t sxc,
                                int64_t N, int64_t Hin, int64_t Win, int64_t Cin,
                                const float* dy, int KH, int KW, int S, int P,
                                int64_t Hout, int64_t Wout, int64_t Cout,
                                float* dw, sg2im_stream_t stream) {
  SG_ARG(x && dy && dw);
  SG_ARG(N >= 1 && Hin >= 1 && Win >= 1 && Cin >= 1 && Hout >= 1 && Wout >= 1 && Cout >= 1);
  SG_ARG(KH >= 1 && KW >= 1 && S >= 1 && P >= 0);
  SG_ARG((Hin + 2 * P - KH) / S + 1 == Hout && (Win + 2 * P - KW) / S + 1 == Wout);
  if (KH == 1 && KW == 1 && S == 1 && P == 0 && Cout <= 4 && Cin <= 256 && sxc == 1 &&
      (Hin * Win == 1 || (sxw * Win == sxh && sxh * Hin == sxn))) {
    // pixels form one dense list of rows with stride sxw (or sxn when H = W = 1)
    int64_t M = N * Hin * Win;
    int64_t xs = (Hin * Win == 1) ? sxn : sxw;
    int64_t blocks = 148 * 8;
    int64_t rpb = ceil_div64(M, blocks);
    if (rpb < 64) rpb = 64;
    blocks = ceil_div64(M, rpb);
    SG_LAUNCH(wgrad_skinny_kernel, (unsigned)blocks, 256, 0, as_stream(stream), x, xs, M, (int)Cin, dy,
                                                                           (int)Cout, rpb, dw);
    SG_LAUNCH_OK();
    return 0;
  }
  WgradP p;
  p.x = x; p.sxn = sxn; p.sxh = sxh; p.sxw = sxw; p.sxc = sxc;
  p.N = N; p.Hin = Hin; p.Win = Win; p.Cin = Cin; p.dy = dy;
  p.KH = KH; p.KW = KW; p.S = S; p.P = P; p.Hout = Hout; p.Wout = Wout; p.Cout = Cout;
  p.dw = dw; p.M = N * Hout * Wout;
  p.ci_tiles = (int)ceil_div64(Cin, BM);
  p.vecX = (sxc == 1) && (Cin % 4 == 0) && (sxn % 4 == 0) && (sxh % 4 == 0) && (sxw % 4 == 0) &&
           aligned16(x);
  p.vecY = (Cout % 4 == 0) && aligned16(dy);
  int64_t out_tiles = (int64_t)KH * KW * p.ci_tiles * ceil_div64(Cout, BN);
  // enough CTAs for ~4 waves of 148 SMs x 2 resident CTAs, >= 64 pixels each
  int64_t split = ceil_div64(148 * 2 * 4, out_tiles);
  int64_t max_split = ceil_div64(p.M, 64);
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  if (split > 65535) split = 65535;
  p.m_per_split = ceil_div64(ceil_div64(p.M, split), BK) * BK;
  split = ceil_div64(p.M, p.m_per_split);
  dim3 grid((unsigned)(KH * KW * p.ci_tiles), (unsigned)ceil_div64(Cout, BN), (unsigned)split);
  SG_LAUNCH(conv_wgrad_kernel, grid, 256, 0, as_stream(stream), p);
  SG_LAUNCH_OK();
  return 0;
}
