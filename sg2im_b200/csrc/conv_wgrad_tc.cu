// tcgen05 weight gradient of a stride-1 convolution (NHWC fp32 in HBM, TF32
// multiply, fp32 accumulate in TMEM):
//
//   dW[tap][ci][co] += sum_pix X[pix + tap - P, ci] * dY[pix, co]
//
// GEMM view per tap: D[ci (M=128), co (N=BN)] over K = pixels.  Both operands
// are "MN-major" exactly as NHWC stores them (a pixel row holds 32 consecutive
// channels = one 128-byte swizzle row), so TMA boxes land in the layout
// tcgen05 consumes (SWIZZLE_128B_ATOM_32B <-> UMMA layout SWIZZLE_128B_BASE32B,
// the only MN-major layout for 32-bit operands) and no transpose ever runs.
//
// * Per pipeline stage one 8 x RH pixel tile: dY tile (32 px rows per 32-channel
//   atom) and ONE X halo tile ((RH+KH-1) x (8+KW-1) pixels).  Every tap's A
//   operand is the same halo tile read at a row-shifted start address
//   ((h+ky)*(8+KW-1)+kx rows; tcgen05 swizzles on absolute smem address bits,
//   verified by tools/umma_probe_mn.cu), so X is fetched once, not KH*KW times.
// * Up to 512/BN taps accumulate side by side in TMEM (one [128 x BN] fp32
//   block each); more taps = more passes over the pixel range.
// * Work item = (ci tile, co tile, tap pass, pixel split); persistent CTAs; the
//   epilogue adds the partial tile into dW with vector atomics (dW zeroed by
//   the caller).
// * Arithmetic (template parameter MATH, as in conv_tc.cu): 0 = kind::tf32 on the fp32 words as
//   they are; 1 = kind::f16 on bf16 pairs — the TMA boxes land as plain SWIZZLE_128B rows, four
//   converter warps fold every pair of adjacent 32-channel atoms in place into one 64-channel
//   bf16 hi atom and one mid atom (tc_common.cuh: split_rowpair_inplace), and each fp32 product
//   is issued as hi*hi + mid*hi + hi*mid (K = 16 pixels per MMA; nprod = 1: hi*hi only).
//   A ci tile with at most 64 channels (the 64 -> 64 layers, the tail of 288 = 128 + 128 + 32)
//   is "stacked": its one folded pair IS an M = 128 operand [hi rows ; mid rows], so two MMAs
//   ([hi;mid] x dY_hi, [hi;mid] x dY_mid) give all four partial products — 2/3 of the tensor time,
//   half of the loads and splits — and the epilogue adds rows r and r + 64 into the same dW row.
// Replaces cuDNN's backward-filter behind nn.Conv2d / nn.Linear
// (sg2im/crn.py:41-45,80-82; model.py:100; layers.py:221).
#include <cstdlib>
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int WG_THREADS = 192;
constexpr int WG_CONV_THREADS = 128;        // one converter group (4 warps) of the bf16 arithmetic
constexpr int WG_CONV_GROUPS = 2;           // the groups take alternate pipeline stages
constexpr int A_ATOM_BYTES = 8192;          // halo tile of one 32-channel atom, padded to 1 KB
constexpr int B_ATOM_BYTES = 4096;          // 32 pixel rows x 128 B
constexpr int A_STAGE = 4 * A_ATOM_BYTES;   // M = 128 channels = 4 atoms

struct WgParams {
  int Cin, Cout, KH, KW, P, taps;
  int RH, pitch;                 // tile rows, halo row pitch (8 + KW - 1) in pixels
  int tiles_w, tiles_h, total_ptiles;
  int ci_tiles, co_tiles, passes, T, splits, per_split;
  int a_bytes, b_atom_rows;      // bytes of one A atom box, pixel rows of a B atom (=8*RH)
  int nprod;                     // MATH 1: products per fp32 multiply (3 = bf16x3, 1 = bf16)
  int s2d_c;                     // > 0: x is the space-to-depth form (4 * s2d_c channels, 2x2 taps) of a
                                 // 4x4 stride-2 convolution; dw is that filter's [16][s2d_c][Cout] gradient
  float* dw;
};

// MN-major SWIZZLE_128B_BASE32B descriptors (lo, hi words): lo = start>>4 | (byte stride
// between 32-channel atoms >> 4)<<16, hi = SBO(512 B = 4 pixel rows)>>4 | version 1<<14 | layout 1<<29.

template <int BN>
struct WCfg {
  static constexpr int B_STAGE = (BN / 32) * B_ATOM_BYTES;
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int STAGES = BN == 256 ? 3 : (BN == 128 ? 4 : 5);
  static constexpr int SMEM_BYTES = STAGES * STAGE + 1024 + 256;
  // D=F32, A=B=TF32, both MN-major (bits 15,16), N>>3, M>>4
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                    ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  static constexpr uint32_t IDESC16 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                      ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};

template <int BN, int MATH>
__global__ void __launch_bounds__(WG_THREADS + (MATH ? WG_CONV_GROUPS * WG_CONV_THREADS : 0), 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                     const WgParams p) {
  using C = WCfg<BN>;
  SG_DYN_SMEM(uint8_t, smem_raw);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
  uint64_t* ready = tempty + 2;                                  // [STAGES] operands split (MATH 1)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = p.ci_tiles * p.co_tiles * p.passes * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1); mbar_init(&empty[i], 1);
      if (MATH) mbar_init(&ready[i], WG_CONV_THREADS / 32);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, 512u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int& ci0, int& co0, int& pass, int& t0, int& t1) {
    int ci = item % p.ci_tiles; item /= p.ci_tiles;
    int co = item % p.co_tiles; item /= p.co_tiles;
    pass = item % p.passes;
    int split = item / p.passes;
    ci0 = ci * 128; co0 = co * BN;
    t0 = split * p.per_split;
    t1 = t0 + p.per_split < p.total_ptiles ? t0 + p.per_split : p.total_ptiles;
  };

  // bf16x3 on a ci tile of <= 64 channels: stacked [hi ; mid] A operand (see the header)
  auto stacked = [&](int ci0) { return MATH == 1 && p.nprod == 3 && p.Cin - ci0 <= 64; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int ci0, co0, pass, t0, t1;
        decode(item, ci0, co0, pass, t0, t1);
        int na = (p.Cin - ci0 + 31) / 32; if (na > 4) na = 4;
        int nb = (p.Cout - co0 + 31) / 32; if (nb > BN / 32) nb = BN / 32;
        if (MATH) { na = stacked(ci0) ? 2 : 4; nb = BN / 32; }   // folded in pairs: fetch whole pairs (past the edge: zeros)
        const uint32_t bytes = (uint32_t)(na * p.a_bytes + nb * B_ATOM_BYTES);
        for (int pt = t0; pt < t1; ++pt) {
          int tw = pt % p.tiles_w;
          int r = pt / p.tiles_w;
          int th = r % p.tiles_h;
          int n = r / p.tiles_h;
          int x0 = tw * 8, y0 = th * p.RH;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], bytes);
          uint8_t* sa = smem + s * C::STAGE;
          uint8_t* sb = sa + A_STAGE;
          for (int a = 0; a < na; ++a)
            tma_load_4d(sa + a * A_ATOM_BYTES, &tmX, &full[s], ci0 + a * 32, x0 - p.P, y0 - p.P, n);
          for (int b = 0; b < nb; ++b)
            tma_load_4d(sb + b * B_ATOM_BYTES, &tmDY, &full[s], co0 + b * 32, x0, y0, n);
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, lane 0 issues) =====================
    {
      const uint32_t leader = lane == 0 ? 1u : 0u;
      // MN-major SWIZZLE_128B_BASE32B: lo = start>>4 | (atom stride>>4)<<16, hi = SBO(512)>>4 | v1 | type 1
      const uint32_t d_hi = 32u | (1u << 14) | (1u << 29);
      const uint32_t a_lo0 = (smem_u32(smem) >> 4) | ((uint32_t)(A_ATOM_BYTES >> 4) << 16);
      const uint32_t b_lo0 = ((smem_u32(smem) + A_STAGE) >> 4) | ((uint32_t)(B_ATOM_BYTES >> 4) << 16);
      const uint32_t pitch16 = (uint32_t)p.pitch * 8u;               // halo row pitch in 16-byte units
      const uint32_t row_wrap = (uint32_t)(p.pitch - p.KW) * 8u;
      int s = 0; uint32_t ph = 0;
      uint32_t acc_ph = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int ci0, co0, pass, t0, t1;
        decode(item, ci0, co0, pass, t0, t1);
        const int tap0 = pass * p.T;
        const int ntap = (p.taps - tap0) < p.T ? (p.taps - tap0) : p.T;
        const int ky0 = tap0 / p.KW, kx0 = tap0 - ky0 * p.KW;
        const uint32_t tap_off0 = (uint32_t)(ky0 * p.pitch + kx0) * 8u;
        const bool stk = stacked(ci0);
        mbar_wait(tempty, acc_ph ^ 1);
        tc_fence_after();
        for (int pt = t0; pt < t1; ++pt) {
          mbar_wait(MATH ? &ready[s] : &full[s], ph);
          tc_fence_after();
          const uint32_t first = (pt > t0) ? 1u : 0u;
          uint32_t at = a_lo0 + (uint32_t)s * (C::STAGE >> 4) + tap_off0;
          const uint32_t bt = b_lo0 + (uint32_t)s * (C::STAGE >> 4);
          uint32_t d_tmem = tmem_base;
          int kx = kx0;
          for (int tl = 0; tl < ntap; ++tl) {
            if constexpr (MATH == 1) {
              // 4 output rows of 8 pixels: K = 16 pixels (two rows) per MMA.  hi atoms are the even
              // 32-channel atoms of each folded pair, mid atoms the odd ones; 64-channel atoms are
              // one pair (LBO) apart; the second 8-pixel group is the next halo row (A) / 1 KB (B)
              // (stacked: rows 64..127 of A are the mid atom, one atom after the hi atom)
              const uint32_t a16 = (at & 0xffffu) | ((uint32_t)((stk ? 1 : 2) * A_ATOM_BYTES >> 4) << 16);
              const uint32_t b16 = (bt & 0xffffu) | ((uint32_t)(2 * B_ATOM_BYTES >> 4) << 16);
              const uint32_t a_hi16 = pitch16 | (1u << 14) | (2u << 29);
              const uint32_t b_hi16 = 64u | (1u << 14) | (2u << 29);
              const int nprod = stk ? 2 : p.nprod;
#pragma unroll
              for (int pr = 0; pr < 3; ++pr) {
                if (pr >= nprod) break;
                const uint32_t ao = (pr == 1 && !stk) ? (uint32_t)(A_ATOM_BYTES >> 4) : 0u;
                const uint32_t bo = (pr == 2 || (pr == 1 && stk)) ? (uint32_t)(B_ATOM_BYTES >> 4) : 0u;
                tc_mma_f16_lh(d_tmem, a16 + ao, a_hi16, b16 + bo, b_hi16, C::IDESC16,
                              pr ? 1u : first, leader);
                tc_mma_f16_lh(d_tmem, a16 + ao + 2 * pitch16, a_hi16, b16 + bo + 128, b_hi16, C::IDESC16,
                              1u, leader);
              }
            } else {
            // 4 output rows of 8 pixels (RH == 4): K = 8 pixels per MMA
            tc_mma_tf32_lh(d_tmem, at, d_hi, bt, d_hi, C::IDESC, first, leader);
            tc_mma_tf32_lh(d_tmem, at + pitch16, d_hi, bt + 64, d_hi, C::IDESC, 1u, leader);
            tc_mma_tf32_lh(d_tmem, at + 2 * pitch16, d_hi, bt + 128, d_hi, C::IDESC, 1u, leader);
            tc_mma_tf32_lh(d_tmem, at + 3 * pitch16, d_hi, bt + 192, d_hi, C::IDESC, 1u, leader);
            }
            d_tmem += BN;
            at += 8u;
            if (++kx == p.KW) { kx = 0; at += row_wrap; }
          }
          tc_commit(&empty[s], leader);
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit(tfull, leader);
        acc_ph ^= 1;
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue: TMEM -> vector atomics into dW =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;                           // ci row within the tile
    uint32_t acc_ph = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int ci0, co0, pass, t0, t1;
      decode(item, ci0, co0, pass, t0, t1);
      mbar_wait_relaxed(tfull, acc_ph, 200);
      tc_fence_after();
      const int ci = ci0 + (stacked(ci0) ? (row & 63) : row);   // stacked: rows r, r + 64 -> the same dW row
      const bool valid = ci < p.Cin;
      for (int tl = 0; tl < p.T; ++tl) {
        int tap = pass * p.T + tl;
        if (tap >= p.taps) break;
        long long wrow = (long long)tap * p.Cin + ci;
        if (p.s2d_c > 0) {                                   // (ty, tx), (py, px, c) -> (2 ty + py, 2 tx + px, c)
          const int blk = ci / p.s2d_c, c = ci - blk * p.s2d_c;
          wrow = (long long)((2 * (tap >> 1) + (blk >> 1)) * 4 + 2 * (tap & 1) + (blk & 1)) * p.s2d_c + c;
        }
        float* drow = p.dw + wrow * p.Cout + co0;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * BN);
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          if (co0 + ch * 32 >= p.Cout) break;
          float v[32];
          tc_ld32(taddr + ch * 32, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              atomicAdd(reinterpret_cast<float4*>(drow + ch * 32 + j),
                        make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      acc_ph ^= 1;
    }
  } else if constexpr (MATH == 1) {
    // ===================== operand converters (warps 6..13, two groups) =====================
    // per landed stage: the two atom pairs of the X halo tile and the BN/64 atom pairs of the dY
    // tile are folded in place into 64-channel bf16 hi / mid atoms.  Group g owns the pipeline slots
    // of parity g, so one group's wait -> split -> fence -> arrive latency overlaps the other's (a
    // slot is always served by the same group: skipping a use would alias the barrier's parity)
    const int ct = ((int)threadIdx.x - 6 * 32) & (WG_CONV_THREADS - 1);
    const int grp = ((int)threadIdx.x - 6 * 32) / WG_CONV_THREADS;
    const int a_rows = p.a_bytes >> 7;
    int s = 0; uint32_t ph = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int ci0, co0, pass, t0, t1;
      decode(item, ci0, co0, pass, t0, t1);
      const int n_a = (stacked(ci0) ? 1 : 2) * a_rows;       // stacked: only the first pair was fetched
      const int n_items = n_a + (BN / 64) * 32;
      for (int pt = t0; pt < t1; ++pt) {
        if ((s & 1) == grp) {
          mbar_wait_relaxed(&full[s], ph, 32);
          uint8_t* sa = smem + s * C::STAGE;
          uint8_t* sb = sa + A_STAGE;
          for (int i = ct; i < n_items; i += WG_CONV_THREADS) {
            if (i < n_a) {
              const int pair = i >= a_rows ? 1 : 0, r = i - pair * a_rows;
              uint8_t* r0 = sa + pair * 2 * A_ATOM_BYTES + r * 128;
              split_rowpair_inplace(r0, r0 + A_ATOM_BYTES);
            } else {
              const int j = i - n_a;
              uint8_t* r0 = sb + (j >> 5) * 2 * B_ATOM_BYTES + (j & 31) * 128;
              split_rowpair_inplace(r0, r0 + B_ATOM_BYTES);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ready[s]);
        }
        if (++s == C::STAGES) { s = 0; ph ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, 512u);
  }
}

template <int BN, int MATH>
int launch_wg(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WgParams& p, cudaStream_t st) {
  using C = WCfg<BN>;
#ifndef SG2IM_EMUL
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<BN, MATH>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      sg2im_set_error("conv_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
#endif
  int items = p.ci_tiles * p.co_tiles * p.passes * p.splits;
  int grid = items < num_sms() ? items : num_sms();
  SG_LAUNCH((conv_wgrad_tc_kernel<BN, MATH>), grid, WG_THREADS + (MATH ? WG_CONV_GROUPS * WG_CONV_THREADS : 0),
            C::SMEM_BYTES, st, tmX, tmDY, p);
  return 0;
}

// A 1x1 "conv" (Linear) has no spatial structure: its pixels are re-viewed as
// rows of 8 (needs N*H*W % 32 == 0).  K > 1: the real (N, H, W) grids; partial
// edge tiles are free because TMA zero-fills dY past the edge.
struct WgGeom { int64_t xN, xH, xW, yN, yH, yW; };

bool wg_geometry(int64_t N, int64_t Hin, int64_t Win, int64_t Hout, int64_t Wout, int KH, int KW,
                 WgGeom& g) {
  if (KH == 1 && KW == 1) {
    int64_t npix = N * Hin * Win;
    if (npix % 32 || Hout != Hin || Wout != Win) return false;
    g.xN = g.yN = 1; g.xH = g.yH = npix / 8; g.xW = g.yW = 8;
    return g.xH < (1ll << 31);
  }
  g.xN = g.yN = N; g.xH = Hin; g.xW = Win; g.yH = Hout; g.yW = Wout;
  return true;
}

}  // namespace

extern "C" int sg2im_conv_wgrad_tc_supported(int64_t N, int64_t Hin, int64_t Win, int64_t Cin,
                                             int64_t x_cstride, int KH, int KW, int S, int P,
                                             int64_t Hout, int64_t Wout, int64_t Cout) {
  if (S != 1 || KH < 1 || KW < 1 || KH > 3 || KW > 3 || P < 0) return 0;
  if (Hout < 1 || Wout < 1 || Hout > Hin + 2 * P - KH + 1 || Wout > Win + 2 * P - KW + 1) return 0;
  if (Cin % 4 || x_cstride % 4 || x_cstride < Cin || Cout % 32) return 0;
  if (KH == 1 && KW == 1 && (P != 0)) return 0;
  WgGeom g;
  return wg_geometry(N, Hin, Win, Hout, Wout, KH, KW, g) ? 1 : 0;
}

extern "C" int sg2im_conv_wgrad_tc(const float* x, int64_t x_cstride, int64_t N, int64_t Hin,
                                   int64_t Win, int64_t Cin, const float* dy, int KH, int KW,
                                   int P, int64_t Hout, int64_t Wout, int64_t Cout, float* dw,
                                   int math, int64_t s2d_channels, sg2im_stream_t stream) {
  SG_ARG(s2d_channels == 0 || (KH == 2 && KW == 2 && P == 0 && Cin == 4 * s2d_channels));
  SG_ARG(x && dy && dw);
  SG_ARG(math == SG2IM_MATH_TF32 || math == SG2IM_MATH_BF16X3 || math == SG2IM_MATH_BF16);
  const int bf = math != SG2IM_MATH_TF32;
  if (!sg2im_conv_wgrad_tc_supported(N, Hin, Win, Cin, x_cstride, KH, KW, 1, P, Hout, Wout, Cout)) {
    sg2im_set_error("sg2im_conv_wgrad_tc: unsupported shape (use sg2im_conv_wgrad)");
    return -2;
  }
  SG_ARG(aligned16(x) && aligned16(dy) && aligned16(dw));
  EncodeTiledFn enc = get_encode();
  if (!enc) { sg2im_set_error("sg2im_conv_wgrad_tc: cuTensorMapEncodeTiled unavailable"); return -3; }
  WgGeom g;
  wg_geometry(N, Hin, Win, Hout, Wout, KH, KW, g);
  const int RH = 4;

  WgParams p;
  p.Cin = (int)Cin; p.Cout = (int)Cout; p.KH = KH; p.KW = KW; p.P = P; p.taps = KH * KW;
  p.RH = RH; p.pitch = 8 + KW - 1;
  p.tiles_w = (int)ceil_div64(g.yW, 8); p.tiles_h = (int)ceil_div64(g.yH, RH);
  p.total_ptiles = (int)(g.yN * p.tiles_h * p.tiles_w);
  int BN = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
  if (const char* e = getenv("SG2IM_WG_BN")) {           // tuning aid: pins the Cout tile
    int f = atoi(e);
    if (f == 64 || f == 128 || f == 256) BN = f;
  }
  p.ci_tiles = (int)ceil_div64(Cin, 128);
  p.co_tiles = (int)ceil_div64(Cout, BN);
  p.passes = (int)ceil_div64((int64_t)p.taps * BN, 512);
  p.T = (int)ceil_div64(p.taps, p.passes);
  p.passes = (int)ceil_div64(p.taps, p.T);
  long long base = (long long)p.ci_tiles * p.co_tiles * p.passes;
  long long want = ceil_div64(3ll * num_sms(), base);
  long long max_split = p.total_ptiles / 8 > 0 ? p.total_ptiles / 8 : 1;
  if (want > max_split) want = max_split;
  if (want < 1) want = 1;
  p.per_split = (int)ceil_div64(p.total_ptiles, want);
  p.splits = (int)ceil_div64(p.total_ptiles, p.per_split);
  p.a_bytes = (RH + KH - 1) * p.pitch * 128;
  p.b_atom_rows = 8 * RH;
  p.nprod = math == SG2IM_MATH_BF16X3 ? 3 : 1;
  p.s2d_c = (int)s2d_channels;
  p.dw = dw;

  // bf16 arithmetic: the converter warps rewrite every tile and assume plain SWIZZLE_128B rows
  const CUtensorMapSwizzle sw = bf ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  CUtensorMap tmX, tmDY;
  {
    // x viewed as (C, gW, gH, gN) pixels-of-8 rows; for K>1 this is the real NHWC tensor
    cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)g.xW, (cuuint64_t)g.xH, (cuuint64_t)g.xN};
    cuuint64_t gstr[3] = {(cuuint64_t)x_cstride * 4, (cuuint64_t)g.xW * x_cstride * 4,
                          (cuuint64_t)g.xH * g.xW * x_cstride * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)p.pitch, (cuuint32_t)(RH + KH - 1), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sg2im_set_error("sg2im_conv_wgrad_tc: encode X failed (%d)", (int)r); return -4; }
  }
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Cout, (cuuint64_t)g.yW, (cuuint64_t)g.yH, (cuuint64_t)g.yN};
    cuuint64_t gstr[3] = {(cuuint64_t)Cout * 4, (cuuint64_t)g.yW * Cout * 4,
                          (cuuint64_t)g.yH * g.yW * Cout * 4};
    cuuint32_t box[4] = {32, 8, (cuuint32_t)RH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmDY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dy), gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sg2im_set_error("sg2im_conv_wgrad_tc: encode dY failed (%d)", (int)r); return -4; }
  }
  cudaStream_t st = as_stream(stream);
  int rc;
  if (bf) {
    rc = BN == 256 ? launch_wg<256, 1>(tmX, tmDY, p, st)
       : BN == 128 ? launch_wg<128, 1>(tmX, tmDY, p, st) : launch_wg<64, 1>(tmX, tmDY, p, st);
  } else {
    rc = BN == 256 ? launch_wg<256, 0>(tmX, tmDY, p, st)
       : BN == 128 ? launch_wg<128, 0>(tmX, tmDY, p, st) : launch_wg<64, 0>(tmX, tmDY, p, st);
  }
  if (rc) return rc;
  SG_LAUNCH_OK();
  return 0;
}
