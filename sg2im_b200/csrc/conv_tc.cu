// tcgen05 (5th-gen tensor core) implicit-GEMM convolution, stride 1, NHWC fp32
// in HBM, fp32 accumulate in TMEM.
//
//   D[pixel, co] = sum_{tap, ci} X[pixel + tap - P, ci] * Wt[tap][co][ci]
//
// * A operand: one TMA 4-D box (32 channels x BW x BH x BI pixels = 128 rows of
//   128 B) per (tap, 32-channel block), fetched at the tap-shifted coordinate;
//   the TMA unit zero-fills out-of-bounds pixels, which IS the conv padding, and
//   lays rows out in the 128B-swizzled K-major form tcgen05 reads directly.
// * B operand: weights [tap][Cout][Cin] (K-major, packed) or read in place from the
//   weight-gradient layout [tap][Cin][Cout] (WMODE 1 / 2, below), 3-D TMA box.
// * Arithmetic (template parameter MATH):
//     0  kind::tf32, M=128 x N=BN x K=8: fp32 data consumed as-is (the tensor core ignores the
//        low 13 mantissa bits) — 2^-11 operand precision, the fast labelled mode;
//     1  kind::f16 on bf16 pairs, M=128 x N=BN x K=16: four converter warps split every landed
//        tile in place into [bf16 hi | bf16 mid] (tc_common.cuh) and each fp32 product is
//        issued as hi*hi + mid*hi + hi*mid (p.nprod = 3, 'bf16x3': 2^-17 operand precision —
//        the mode that meets the 1e-3 parity bar against the fp32 reference) or as hi*hi only
//        (p.nprod = 1, plain bf16).  HBM / L2 / TMA traffic is identical to mode 0.
// * Accumulators double-buffered in TMEM (2 x BN columns) so the epilogue of tile
//   i overlaps the MMAs of tile i+1; persistent CTAs, static tile striding.
// * Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5
//   epilogue (tcgen05.ld -> bias/LeakyReLU -> 128-bit global stores into the
//   destination channel slice), warps 6-13 operand converters (MATH 1 only; two groups of four
//   warps that take alternate pipeline stages).
// Replaces cuDNN's conv behind nn.Conv2d (sg2im/crn.py:41-45,80-82;
// model.py:100) for the shapes that dominate the step.
#include <cstdlib>
#include "tc_common.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int KB_BYTES = 128;                 // 32 fp32 channels per k-block row
constexpr int A_STAGE_BYTES = TILE_M * KB_BYTES;
constexpr int NUM_THREADS = 192;
constexpr int CONV_THREADS = 128;             // one converter group (4 warps) of the bf16 arithmetic
constexpr int CONV_GROUPS = 2;                // per-tap kernel: groups take alternate pipeline stages

struct TcParams {
  int N, Hout, Wout, Cin, Cout;
  int KH, KW, P;
  int BW, BH, BI;                  // tile = BI images x BH rows x BW cols (=128 pixels)
  int tiles_w, tiles_h, tiles_n, n_tiles;   // n_tiles over Cout
  int num_kb;                      // taps * ceil(Cin/32)
  int cblocks;                     // ceil(Cin/32)
  const float* bias;
  int act; float slope;
  float* y; long long y_cstride, y_coff;
  double* stats;                   // optional [2*Cout] per-channel sum / sum of squares
  int round_out;                   // write RN-TF32 values (output feeds another tf32 tensor-core op)
  int nprod;                       // MATH 1: products per fp32 multiply (3 = bf16x3, 1 = bf16)
};

using namespace tc;

// K-major SWIZZLE_128B descriptors are assembled in the MMA warps as (lo, hi) words:
// lo = start>>4 | LBO(=1, ignored)<<16, hi = SBO>>4 | version 1<<14 | layout 2<<29.
// bf16 tiles after the in-place split: a 128-byte row = [hi: 2 K-steps of 32 B | mid: 2 K-steps],
// so hi sits at +0 / +2 and mid at +4 / +6 (16-byte units) of the same descriptor.

template <int BN>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * KB_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;                     // 128 / 256 / 512
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 + 8192 /*BN partials*/;
  // instruction descriptor: D=F32 (1<<4), A/B format at bits 7 / 10 (TF32 = 2, BF16 = 1), K-major
  // both, N>>3 at bit 17, M>>4 at bit 24
  static constexpr uint32_t IDESC =
      (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  static constexpr uint32_t IDESC16 =
      (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
};

// WMODE selects where the B operand (weights) comes from:
//   0  packed [tap][Cout][Cin] (K-major; sg2im_pack_weights)
//   1  FORWARD straight from the weight-gradient layout [tap][Cin][Cout]: B is MN-major
//      (tf32: 32-co atoms of 32 ci rows, SWIZZLE_128B_BASE32B like the wgrad kernel's operands;
//      bf16: the converters fold atom pairs into 64-co hi / mid atoms, plain SWIZZLE_128B)
//   2  DATA GRADIENT straight from the same layout [tap][Cin_w][Cout_w]: conv-Cin = Cout_w is
//      contiguous => K-major as in mode 0, only the tap index is flipped
// Modes 1/2 remove every pack / unpack pass when the master weights live in that layout
// (sg2im_conv_tc_kcc; validated on the B200 in round 2).
// PS (MATH 1, WMODE 0 only): the B operand arrives PRE-SPLIT from HBM (sg2im_split_weights wrote
// its rows as [hi | mid] blocks once per training step), so the converters touch only A.
template <int BN, int WMODE, int MATH, int PS = 0>
__global__ void __launch_bounds__(NUM_THREADS + (MATH ? CONV_GROUPS * CONV_THREADS : 0), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcParams p) {
  using C = Cfg<BN>;
  SG_DYN_SMEM(uint8_t, smem_raw);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* sA = smem;                                          // STAGES x 16 KB
  uint8_t* sB = smem + C::STAGES * A_STAGE_BYTES;              // STAGES x BN*128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;                                       // [STAGES]
  uint64_t* empty = bars + C::STAGES;                          // [STAGES]
  uint64_t* tfull = bars + 2 * C::STAGES;                      // [2]
  uint64_t* tempty = bars + 2 * C::STAGES + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  uint64_t* ready = bars + 2 * C::STAGES + 5;                  // [STAGES] operands split (MATH 1)
  float* s_part = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256);   // [2][1024]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_n * p.tiles_h * p.tiles_w * p.n_tiles;
  if (p.stats)
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s_part[i] = 0.f;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1); mbar_init(&empty[i], 1);
      if (MATH) mbar_init(&ready[i], CONV_THREADS / 32);
    }
    mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 4); mbar_init(&tempty[1], 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, (uint32_t)C::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (n-tile over Cout fastest, so CTAs running together share A in L2)
  auto decode = [&](int tile, int& nt, int& n0, int& y0, int& x0) {
    nt = tile % p.n_tiles;
    int m = tile / p.n_tiles;
    int tw = m % p.tiles_w; m /= p.tiles_w;
    int th = m % p.tiles_h; m /= p.tiles_h;
    n0 = m * p.BI; y0 = th * p.BH; x0 = tw * p.BW;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int nt, n0, y0, x0;
        decode(tile, nt, n0, y0, x0);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
          int ky = tap / p.KW, kx = tap - ky * p.KW;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], C::STAGE_BYTES);
          tma_load_4d(sA + s * A_STAGE_BYTES, &tmA, &full[s], cb * 32, x0 + kx - p.P,
                      y0 + ky - p.P, n0);
          if constexpr (WMODE == 1) {
#pragma unroll
            for (int a = 0; a < BN / 32; ++a)
              tma_load_3d(sB + s * C::B_STAGE_BYTES + a * 4096, &tmB, &full[s], nt * BN + a * 32,
                          cb * 32, tap);
          } else {
            tma_load_3d(sB + s * C::B_STAGE_BYTES, &tmB, &full[s], cb * 32, nt * BN,
                        WMODE == 2 ? p.KH * p.KW - 1 - tap : tap);
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, lane 0 issues) =====================
    {
      const uint32_t leader = lane == 0 ? 1u : 0u;
      const uint32_t d_hi = 64u | (1u << 14) | (2u << 29);          // SBO 1024 B, v1, SWIZZLE_128B
      const uint32_t sA16 = (smem_u32(sA) >> 4) | (1u << 16);
      const uint32_t sB16 = (smem_u32(sB) >> 4) | (1u << 16);
      int s = 0; uint32_t ph = 0;
      int acc = 0; uint32_t acc_ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(MATH ? &ready[s] : &full[s], ph);
          tc_fence_after();
          const uint32_t at = sA16 + (uint32_t)s * (A_STAGE_BYTES >> 4);
          const uint32_t bt = sB16 + (uint32_t)s * (C::B_STAGE_BYTES >> 4);
          if constexpr (MATH == 1) {
            // products hi*hi, mid*hi, hi*mid; two K = 16 steps (32 B) per 32-channel block
#pragma unroll
            for (int pr = 0; pr < 3; ++pr) {
              if (pr >= p.nprod) break;
              const uint32_t ao = pr == 1 ? 4u : 0u;
              if constexpr (WMODE == 1) {
                // B MN-major bf16: 64-co atoms (hi = even 4 KB atom of a pair, mid = odd), pairs
                // 8 KB apart (LBO), 8 ci rows = 1 KB per group (SBO), 16 ci rows per K step
                constexpr uint32_t IDESC_MN = C::IDESC16 | (1u << 16);
                const uint32_t bm_hi = 64u | (1u << 14) | (2u << 29);
                const uint32_t btm = ((bt + (pr == 2 ? 256u : 0u)) & 0xffffu) | ((8192u >> 4) << 16);
                tc_mma_f16_lh(d_tmem, at + ao, d_hi, btm, bm_hi, IDESC_MN, (kb | pr) ? 1u : 0u, leader);
                tc_mma_f16_lh(d_tmem, at + ao + 2, d_hi, btm + 128, bm_hi, IDESC_MN, 1u, leader);
              } else {
                const uint32_t bo = pr == 2 ? 4u : 0u;
                tc_mma_f16_lh(d_tmem, at + ao, d_hi, bt + bo, d_hi, C::IDESC16, (kb | pr) ? 1u : 0u, leader);
                tc_mma_f16_lh(d_tmem, at + ao + 2, d_hi, bt + bo + 2, d_hi, C::IDESC16, 1u, leader);
              }
            }
          } else if constexpr (WMODE == 1) {
            // B MN-major: atoms of 32 co (LBO = 4 KB apart), 8 ci rows = 1 KB per K step
            constexpr uint32_t IDESC_MN = C::IDESC | (1u << 16);
            const uint32_t bm_hi = 32u | (1u << 14) | (1u << 29);         // SBO 512 B, SWIZZLE_128B_BASE32B
            const uint32_t btm = (bt & 0xffffu) | ((4096u >> 4) << 16);
            tc_mma_tf32_lh(d_tmem, at, d_hi, btm, bm_hi, IDESC_MN, kb ? 1u : 0u, leader);
            tc_mma_tf32_lh(d_tmem, at + 2, d_hi, btm + 64, bm_hi, IDESC_MN, 1u, leader);
            tc_mma_tf32_lh(d_tmem, at + 4, d_hi, btm + 128, bm_hi, IDESC_MN, 1u, leader);
            tc_mma_tf32_lh(d_tmem, at + 6, d_hi, btm + 192, bm_hi, IDESC_MN, 1u, leader);
          } else {
            tc_mma_tf32_lh(d_tmem, at, d_hi, bt, d_hi, C::IDESC, kb ? 1u : 0u, leader);
            tc_mma_tf32_lh(d_tmem, at + 2, d_hi, bt + 2, d_hi, C::IDESC, 1u, leader);
            tc_mma_tf32_lh(d_tmem, at + 4, d_hi, bt + 4, d_hi, C::IDESC, 1u, leader);
            tc_mma_tf32_lh(d_tmem, at + 6, d_hi, bt + 6, d_hi, C::IDESC, 1u, leader);
          }
          tc_commit(&empty[s], leader);                    // frees the smem stage when the MMAs retire
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit(&tfull[acc], leader);                    // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                                // TMEM lane quadrant this warp may read
    const int r = q * 32 + lane;                           // tile row = TMEM lane
    const int img = r / (p.BH * p.BW);
    const int hh = (r / p.BW) % p.BH, ww = r % p.BW;
    int acc = 0; uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int nt, n0, y0, x0;
      decode(tile, nt, n0, y0, x0);
      mbar_wait_relaxed(&tfull[acc], acc_ph, 200);
      tc_fence_after();
      const int n = n0 + img;
      const bool valid = n < p.N && (y0 + hh) < p.Hout && (x0 + ww) < p.Wout;
      float* yrow = p.y + (((long long)n * p.Hout + (y0 + hh)) * p.Wout + (x0 + ww)) * p.y_cstride +
                    p.y_coff + (long long)nt * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        if (nt * BN + ch * 32 >= p.Cout) break;              // partial last N tile (warp-uniform)
        float v[32];
        tc_ld32(taddr + ch * 32, v);
        epilogue_chunk(v, valid, nt * BN + ch * 32, p.Cout, p.bias, p.act, p.slope, yrow + ch * 32,
                       p.stats ? s_part : nullptr, lane, p.round_out);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  } else if constexpr (MATH == 1) {
    // ===================== operand converters (warps 6..13, two groups) =====================
    // every landed stage: the 128 A rows and the B rows (WMODE 1: the row pairs of adjacent
    // 32-co atoms) are split in place, then each warp of the group signals `ready`.  Group g owns
    // the pipeline slots of parity g (STAGES is even: consecutive stages alternate), so the fixed
    // latency of one wait -> load -> split -> store -> fence -> arrive round trip overlaps with the
    // other group's.  A slot must always be served by the SAME group: a group that skipped one use
    // of a slot would find the full barrier's parity wait satisfied one lap early
    const int ct = ((int)threadIdx.x - 6 * 32) & (CONV_THREADS - 1);
    const int grp = ((int)threadIdx.x - 6 * 32) / CONV_THREADS;
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        if ((s & 1) == grp) {
          mbar_wait_relaxed(&full[s], ph, 32);
          uint8_t* a = sA + s * A_STAGE_BYTES;
          uint8_t* b = sB + s * C::B_STAGE_BYTES;
          split_row_inplace(a + ct * 128);
          if constexpr (PS) {
            (void)b;
          } else if constexpr (WMODE == 1) {
            for (int i = ct; i < BN / 2; i += CONV_THREADS) {
              uint8_t* r0 = b + (i >> 5) * 8192 + (i & 31) * 128;
              split_rowpair_inplace(r0, r0 + 4096);
            }
          } else {
            for (int i = ct; i < BN; i += CONV_THREADS) split_row_inplace(b + i * 128);
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
  if (p.stats) {
    for (int c = threadIdx.x; c < p.Cout; c += blockDim.x) {
      float a = s_part[c], b = s_part[1024 + c];
      if (a != 0.f || b != 0.f) { atomicAdd(p.stats + c, (double)a); atomicAdd(p.stats + p.Cout + c, (double)b); }
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, (uint32_t)C::TMEM_COLS);
  }
}

// =============================================================================
// Halo / weight-stationary variant for narrow outputs (Cout tile of 64) on real
// images (KH*KW > 1, H >= 16, W >= 8).  N = 64 MMAs are bound by the shared-memory
// operand feed (48.1 cycles per 128x64 MMA measured, tools/umma_rate_probe.cu),
// so the kernel minimises everything else: per 128x64 tile and tap the generic kernel
// refills 16 KB (A) + 8 KB (B) of shared memory from L2, which caps it well below that.
// Here
//   * ONE TMA halo box per (pixel tile, 32-channel block) serves all KH*KW taps:
//     the A tile of tap (ky,kx) is the same shared memory read with the matrix
//     descriptor's start address shifted by (ky*pitch + kx) rows and a stride of
//     `pitch` rows between 8-row groups (tile = 8 columns x 16 rows, so each
//     8-row core group is one image row of 8 pixels; tcgen05 applies the 128B
//     swizzle to the final address, so the shift is free — validated on the B200);
//   * the KH*KW weight tiles are loaded once and reused by T = 4 pixel tiles
//     whose accumulators sit side by side in TMEM (2 sets x 4 x 64 columns,
//     so the epilogue of one group overlaps the MMAs of the next).
// L2->SM bytes per MMA drop ~5x.  Warp roles: 0 = halo (A) producer, 1 = MMA
// issuer + TMEM owner, 2 = weight (B) producer, 4-7 = epilogue; bf16 arithmetic only:
// 3 = weight-tile converter, 8-11 = halo-box converters (every halo box and weight
// tile is split once and then read by all the taps / pixel tiles it serves).
// =============================================================================
constexpr int H_BW = 8, H_BH = 16;
// Cout tile HB = 64: T = 4 pixel tiles per weight set, two weight sets (double buffered per channel
// block).  HB = 128 (outputs >= 128 channels wide): T = 2, ONE weight set whose tap tiles are
// refilled for the next channel block as soon as the last pixel tile has used them (eight taps of
// MMA time ahead of their next use) — N = 128 MMAs run at the math rate (64.1 cycles measured)
// instead of the shared-memory-bound 48.1 per N = 64 half, and every activation tile is read,
// converted and fed once per 128 instead of once per 64 output channels.
template <int HB> struct HCfg {
  static constexpr int T = HB == 64 ? 4 : 2;
  static constexpr int SETS = HB == 64 ? 2 : 1;
  static constexpr int B_TILE = HB * KB_BYTES;          // 8 / 16 KB
};
constexpr int H_A_SLOT = 23 * 1024;           // 180 rows x 128 B = 23040, padded to 1 KB
constexpr int H_A_SLOTS = 3;
constexpr int H_MAX_TAPS = 9;
constexpr int H_THREADS = 256;
constexpr int H_SMEM = H_A_SLOTS * H_A_SLOT + 2 * H_MAX_TAPS * 64 * KB_BYTES + 1024 + 1024 + 8192;

struct HaloParams {
  int N, Hout, Wout, Cin, Cout;
  int KH, KW, P, taps, pitch;      // pitch = 8 + KW - 1 halo pixels per tile row
  int tiles_w, tiles_h, ptiles;    // pixel tiles of 8 x 16
  int groups, n_tiles, cblocks;    // groups of T pixel tiles; Cout tiles of HB
  uint32_t a_bytes;                // bytes of one halo box
  const float* bias;
  int act; float slope;
  float* y; long long y_cstride, y_coff;
  double* stats;
  int round_out;
  int nprod;                       // MATH 1: 3 = bf16x3, 1 = bf16
};

template <int WMODE, int MATH, int PS = 0, int HB = 64>   // weight source / arithmetic / pre-split B / Cout tile
__global__ void __launch_bounds__(H_THREADS + (MATH ? CONV_THREADS : 0), 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const HaloParams p) {
  constexpr int H_BN = HB, H_T = HCfg<HB>::T, H_B_TILE = HCfg<HB>::B_TILE, SETS = HCfg<HB>::SETS;
  SG_DYN_SMEM(uint8_t, smem_raw);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + H_A_SLOTS * H_A_SLOT;                    // [2][taps][8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + SETS * H_MAX_TAPS * H_B_TILE);
  uint64_t* a_full = bars;                    // [3]
  uint64_t* a_empty = bars + 3;               // [3]
  uint64_t* b_full = bars + 6;                // [2][9]
  uint64_t* b_empty = bars + 6 + 18;          // [2][9]
  uint64_t* tfull = bars + 6 + 36;            // [2]
  uint64_t* tempty = bars + 6 + 38;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 40);
  uint64_t* a_ready = bars + 48;              // [3]     operands split (MATH 1)
  uint64_t* b_ready = bars + 51;              // [2][9]
  float* s_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 1024);   // [2][1024]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = p.groups * p.n_tiles;
  if (p.stats)
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s_part[i] = 0.f;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1);
      if (MATH) mbar_init(&a_ready[i], CONV_THREADS / 32);
    }
    for (int i = 0; i < 18; ++i) {
      mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1);
      if (MATH) mbar_init(&b_ready[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, 512u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item -> (Cout tile, first pixel tile of the group)
  auto decode = [&](int item, int& nt, int& pt0) {
    nt = item % p.n_tiles;
    pt0 = (item / p.n_tiles) * H_T;
  };
  auto tile_xy = [&](int pt, int& n, int& y0, int& x0) {
    int tw = pt % p.tiles_w; int r = pt / p.tiles_w;
    int th = r % p.tiles_h; n = r / p.tiles_h;            // n >= N for padding tiles: TMA zero-fills
    y0 = th * H_BH; x0 = tw * H_BW;
  };

  if (warp == 0) {
    // ===================== halo (A) producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        int nt, pt0;
        decode(item, nt, pt0);
        for (int cb = 0; cb < p.cblocks; ++cb) {
          for (int t = 0; t < H_T; ++t) {
            int n, y0, x0;
            tile_xy(pt0 + t, n, y0, x0);
            mbar_wait(&a_empty[s], ph ^ 1);
            mbar_expect_tx(&a_full[s], p.a_bytes);
            tma_load_4d(sA + s * H_A_SLOT, &tmA, &a_full[s], cb * 32, x0 - p.P, y0 - p.P, n);
            if (++s == H_A_SLOTS) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== weight (B) producer =====================
    if (lane == 0) {
      uint32_t bcnt = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        int nt, pt0;
        decode(item, nt, pt0);
        for (int cb = 0; cb < p.cblocks; ++cb, ++bcnt) {
          const int set = (int)(bcnt % SETS); const uint32_t bph = (bcnt / SETS) & 1;
          for (int tap = 0; tap < p.taps; ++tap) {
            uint64_t* fb = &b_full[set * H_MAX_TAPS + tap];
            mbar_wait(&b_empty[set * H_MAX_TAPS + tap], bph ^ 1);
            mbar_expect_tx(fb, H_B_TILE);
            if constexpr (WMODE == 1) {
#pragma unroll
              for (int a = 0; a < H_BN / 32; ++a)
                tma_load_3d(sB + (set * H_MAX_TAPS + tap) * H_B_TILE + a * 4096, &tmB, fb,
                            nt * H_BN + a * 32, cb * 32, tap);
            } else {
              tma_load_3d(sB + (set * H_MAX_TAPS + tap) * H_B_TILE, &tmB, fb, cb * 32, nt * H_BN,
                          WMODE == 2 ? p.taps - 1 - tap : tap);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, lane 0 issues) =====================
    {
      const uint32_t leader = lane == 0 ? 1u : 0u;
      constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H_BN >> 3) << 17) |
                                 ((uint32_t)(TILE_M >> 4) << 24);
      constexpr uint32_t IDESC16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(H_BN >> 3) << 17) |
                                   ((uint32_t)(TILE_M >> 4) << 24);
      // descriptor words: lo = start>>4 | LBO(=1)<<16, hi = SBO>>4 | version 1<<14 | SWIZZLE_128B<<29
      // A: 8-row groups `pitch` rows apart (halo rows); B: dense (1024 B)
      const uint32_t a_hi = (uint32_t)((p.pitch * 128) >> 4) | (1u << 14) | (2u << 29);
      const uint32_t b_hi = 64u | (1u << 14) | (2u << 29);
      const uint32_t sA16 = (smem_u32(sA) >> 4) | (1u << 16);
      const uint32_t sB16 = (smem_u32(sB) >> 4) | (1u << 16);
      const uint32_t row_wrap = (uint32_t)(p.pitch - p.KW) * 8u;     // 16-byte units, row = 128 B
      int s = 0; uint32_t ph = 0;
      uint32_t bcnt = 0;
      int aset = 0; uint32_t acc_ph = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        mbar_wait(&tempty[aset], acc_ph ^ 1);
        tc_fence_after();
        for (int cb = 0; cb < p.cblocks; ++cb, ++bcnt) {
          const int set = (int)(bcnt % SETS); const uint32_t bph = (bcnt / SETS) & 1;
          const uint32_t b_set = sB16 + (uint32_t)(set * H_MAX_TAPS) * (H_B_TILE >> 4);
          uint64_t* bf = (MATH && !PS) ? &b_ready[set * H_MAX_TAPS] : &b_full[set * H_MAX_TAPS];
          uint64_t* be = &b_empty[set * H_MAX_TAPS];
          for (int t = 0; t < H_T; ++t) {
            mbar_wait(MATH ? &a_ready[s] : &a_full[s], ph);
            tc_fence_after();
            uint32_t at = sA16 + (uint32_t)s * (H_A_SLOT >> 4);
            uint32_t bt = b_set;
            const uint32_t d_tmem = tmem_base + (uint32_t)((aset * H_T + t) * H_BN);
            int kx = 0;
            for (int tap = 0; tap < p.taps; ++tap) {
              if (t == 0) { mbar_wait(&bf[tap], bph); tc_fence_after(); }
              if constexpr (MATH == 1) {
#pragma unroll
                for (int pr = 0; pr < 3; ++pr) {
                  if (pr >= p.nprod) break;
                  const uint32_t ao = pr == 1 ? 4u : 0u;
                  const uint32_t accf = (cb | tap | pr) ? 1u : 0u;
                  if constexpr (WMODE == 1) {
                    // B MN-major bf16: one 64-co atom (hi = first 4 KB, mid = second), 8 ci rows per
                    // 1 KB group, 16 ci rows per K step
                    constexpr uint32_t IDESC_MN = IDESC16 | (1u << 16);
                    const uint32_t btm = ((bt + (pr == 2 ? 256u : 0u)) & 0xffffu) | ((8192u >> 4) << 16);
                    tc_mma_f16_lh(d_tmem, at + ao, a_hi, btm, b_hi, IDESC_MN, accf, leader);
                    tc_mma_f16_lh(d_tmem, at + ao + 2, a_hi, btm + 128, b_hi, IDESC_MN, 1u, leader);
                  } else {
                    const uint32_t bo = pr == 2 ? 4u : 0u;
                    tc_mma_f16_lh(d_tmem, at + ao, a_hi, bt + bo, b_hi, IDESC16, accf, leader);
                    tc_mma_f16_lh(d_tmem, at + ao + 2, a_hi, bt + bo + 2, b_hi, IDESC16, 1u, leader);
                  }
                }
              } else if constexpr (WMODE == 1) {
                constexpr uint32_t IDESC_MN = IDESC | (1u << 16);
                const uint32_t bm_hi = 32u | (1u << 14) | (1u << 29);     // SBO 512 B, SWIZZLE_128B_BASE32B
                const uint32_t btm = (bt & 0xffffu) | ((4096u >> 4) << 16);
                tc_mma_tf32_lh(d_tmem, at, a_hi, btm, bm_hi, IDESC_MN, (cb | tap) ? 1u : 0u, leader);
                tc_mma_tf32_lh(d_tmem, at + 2, a_hi, btm + 64, bm_hi, IDESC_MN, 1u, leader);
                tc_mma_tf32_lh(d_tmem, at + 4, a_hi, btm + 128, bm_hi, IDESC_MN, 1u, leader);
                tc_mma_tf32_lh(d_tmem, at + 6, a_hi, btm + 192, bm_hi, IDESC_MN, 1u, leader);
              } else {
                tc_mma_tf32_lh(d_tmem, at, a_hi, bt, b_hi, IDESC, (cb | tap) ? 1u : 0u, leader);
                tc_mma_tf32_lh(d_tmem, at + 2, a_hi, bt + 2, b_hi, IDESC, 1u, leader);
                tc_mma_tf32_lh(d_tmem, at + 4, a_hi, bt + 4, b_hi, IDESC, 1u, leader);
                tc_mma_tf32_lh(d_tmem, at + 6, a_hi, bt + 6, b_hi, IDESC, 1u, leader);
              }
              if (t == H_T - 1) tc_commit(&be[tap], leader);
              at += 8u; bt += (H_B_TILE >> 4);
              if (++kx == p.KW) { kx = 0; at += row_wrap; }
            }
            tc_commit(&a_empty[s], leader);
            if (++s == H_A_SLOTS) { s = 0; ph ^= 1; }
          }
        }
        tc_commit(&tfull[aset], leader);
        if (++aset == 2) { aset = 0; acc_ph ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue (warps 4..7) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int hh = r >> 3, ww = r & 7;                      // 8-wide x 16-high tile
    int aset = 0; uint32_t acc_ph = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      int nt, pt0;
      decode(item, nt, pt0);
      mbar_wait_relaxed(&tfull[aset], acc_ph, 200);
      tc_fence_after();
      for (int t = 0; t < H_T; ++t) {
        int n, y0, x0;
        tile_xy(pt0 + t, n, y0, x0);
        const bool valid = (pt0 + t) < p.ptiles && n < p.N && (y0 + hh) < p.Hout && (x0 + ww) < p.Wout;
        float* yrow = p.y + (((long long)n * p.Hout + (y0 + hh)) * p.Wout + (x0 + ww)) * p.y_cstride +
                      p.y_coff + (long long)nt * H_BN;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((aset * H_T + t) * H_BN);
#pragma unroll 1
        for (int ch = 0; ch < H_BN / 32; ++ch) {
          if (nt * H_BN + ch * 32 >= p.Cout) break;
          float v[32];
          tc_ld32(taddr + ch * 32, v);
          epilogue_chunk(v, valid, nt * H_BN + ch * 32, p.Cout, p.bias, p.act, p.slope,
                         yrow + ch * 32, p.stats ? s_part : nullptr, lane, p.round_out);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[aset]);
      if (++aset == 2) { aset = 0; acc_ph ^= 1; }
    }
  } else if (warp == 3) {
    if constexpr (MATH == 1 && !PS) {
      // ===================== weight-tile converter (warp 3) =====================
      // each tap's 64 weight rows (WMODE 1: its 32 row pairs) as soon as the tile lands, in the
      // order the MMA issuer first touches them; independent of the halo converters below
      uint32_t bcnt = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        for (int cb = 0; cb < p.cblocks; ++cb, ++bcnt) {
          const int set = (int)(bcnt % SETS); const uint32_t bph = (bcnt / SETS) & 1;
          for (int tap = 0; tap < p.taps; ++tap) {
            mbar_wait_relaxed(&b_full[set * H_MAX_TAPS + tap], bph, 32);
            uint8_t* b = sB + (set * H_MAX_TAPS + tap) * H_B_TILE;
            if constexpr (WMODE == 1) {
#pragma unroll
              for (int q = 0; q < H_BN / 64; ++q)
                split_rowpair_inplace(b + q * 8192 + lane * 128, b + q * 8192 + 4096 + lane * 128);
            } else {
#pragma unroll
              for (int q = 0; q < H_BN / 32; ++q) split_row_inplace(b + (q * 32 + lane) * 128);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_ready[set * H_MAX_TAPS + tap]);
          }
        }
      }
    }
  } else if (warp >= 8) {
    if constexpr (MATH == 1) {
      // ===================== halo-box converters (warps 8..11) =====================
      const int ct = (int)threadIdx.x - 8 * 32;
      const int a_rows = (int)(p.a_bytes >> 7);
      int s = 0; uint32_t ph = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        for (int cb = 0; cb < p.cblocks; ++cb) {
          for (int t = 0; t < H_T; ++t) {
            mbar_wait_relaxed(&a_full[s], ph, 32);
            uint8_t* a = sA + s * H_A_SLOT;
            for (int i = ct; i < a_rows; i += CONV_THREADS) split_row_inplace(a + i * 128);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[s]);
            if (++s == H_A_SLOTS) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.stats) {
    for (int c = threadIdx.x; c < p.Cout; c += blockDim.x) {
      float a = s_part[c], b = s_part[1024 + c];
      if (a != 0.f || b != 0.f) { atomicAdd(p.stats + c, (double)a); atomicAdd(p.stats + p.Cout + c, (double)b); }
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, 512u);
  }
}

// ------------------------------------------------------------- host side ---
template <int BN, int WMODE, int MATH, int PS = 0>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t st) {
  using C = Cfg<BN>;
#ifndef SG2IM_EMUL
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, WMODE, MATH, PS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      sg2im_set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
#endif
  int total = p.tiles_n * p.tiles_h * p.tiles_w * p.n_tiles;
  int grid = total < num_sms() ? total : num_sms();
  SG_LAUNCH((conv_tc_kernel<BN, WMODE, MATH, PS>), grid, NUM_THREADS + (MATH ? CONV_GROUPS * CONV_THREADS : 0),
            C::SMEM_BYTES, st, tmA, tmB, p);
  return 0;
}

template <int BN, int MATH>
int launch_w(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, int wmode, cudaStream_t st) {
  return wmode == 1 ? launch<BN, 1, MATH>(tmA, tmB, p, st)
       : wmode == 2 ? launch<BN, 2, MATH>(tmA, tmB, p, st) : launch<BN, 0, MATH>(tmA, tmB, p, st);
}

template <int WMODE, int MATH, int PS = 0, int HB = 64>
int launch_halo(const CUtensorMap& hA, const CUtensorMap& hB, const HaloParams& h, cudaStream_t st) {
#ifndef SG2IM_EMUL
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_halo_kernel<WMODE, MATH, PS, HB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM);
    if (e != cudaSuccess) {
      sg2im_set_error("conv_tc_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
#endif
  int items = h.groups * h.n_tiles;
  int grid = items < num_sms() ? items : num_sms();
  SG_LAUNCH((conv_tc_halo_kernel<WMODE, MATH, PS, HB>), grid, H_THREADS + (MATH ? CONV_THREADS : 0), H_SMEM,
            st, hA, hB, h);
  return 0;
}

}  // namespace

// Tensor map of the B operand for the three weight sources (see conv_tc_kernel).  bf16 arithmetic:
// every tile is rewritten by the converter warps, which assume the plain SWIZZLE_128B pattern.
static int encode_weights(tc::EncodeTiledFn enc, CUtensorMap* map, const float* w, int64_t Cin,
                          int64_t Cout, int taps, int BN, int wmode, int64_t w_rows_full, int math,
                          int64_t ps_pitch = 0) {
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t box[3] = {32, (cuuint32_t)BN, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  if (wmode == 0 && ps_pitch) {           // pre-split [taps][w_rows_full][ps_pitch], Cout rows used
    gdim[0] = (cuuint64_t)ps_pitch; gdim[1] = (cuuint64_t)Cout;
    gstr[0] = (cuuint64_t)ps_pitch * 4; gstr[1] = (cuuint64_t)w_rows_full * ps_pitch * 4;
  } else if (wmode == 0) {                // [taps][Cout][Cin]
    gdim[0] = (cuuint64_t)Cin; gdim[1] = (cuuint64_t)Cout;
    gstr[0] = (cuuint64_t)Cin * 4; gstr[1] = (cuuint64_t)Cout * Cin * 4;
  } else if (wmode == 1) {                // [taps][rows = Cin][cols = Cout], cols contiguous, MN-major atoms
    gdim[0] = (cuuint64_t)Cout; gdim[1] = (cuuint64_t)Cin;
    gstr[0] = (cuuint64_t)Cout * 4; gstr[1] = (cuuint64_t)w_rows_full * Cout * 4;
    box[1] = 32;
    if (!math) sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  } else {                                // [taps][rows = conv Cout][cols = conv Cin], K-major
    gdim[0] = (cuuint64_t)Cin; gdim[1] = (cuuint64_t)Cout;
    gstr[0] = (cuuint64_t)Cin * 4; gstr[1] = (cuuint64_t)w_rows_full * Cin * 4;
  }
  gdim[2] = (cuuint64_t)taps;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(w), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { sg2im_set_error("sg2im_conv_tc: encode B failed (%d)", (int)r); return -4; }
  return 0;
}

// Tile geometry for an output of Hout x Wout: 128 pixels = BI images x BH rows x
// BW cols, all powers of two; the last tile in each direction may be partial
// (TMA reads past the edge are zero-filled or unused, the epilogue masks).
static void tc_geometry(long long Hout, long long Wout, int& BW, int& BH, int& BI) {
  BW = Wout > 8 ? 16 : (Wout > 4 ? 8 : (Wout > 2 ? 4 : (Wout > 1 ? 2 : 1)));
  int rem = TILE_M / BW;
  BH = 1;
  while (BH < rem && BH < Hout) BH <<= 1;
  BI = rem / BH;
}

extern "C" int sg2im_conv_tc_supported(int64_t N, int64_t Hin, int64_t Win, int64_t Cin,
                                       int64_t x_cstride, int KH, int KW, int S, int P,
                                       int64_t Hout, int64_t Wout, int64_t Cout,
                                       int64_t y_cstride, int64_t y_coff) {
  if (S != 1 || KH < 1 || KW < 1 || KH * KW > 64 || P < 0) return 0;
  // any Hout/Wout: reads outside the input are zero-filled by TMA
  if (Hout < 1 || Wout < 1 || Hout > 65536 || Wout > 65536) return 0;
  if (Cin % 4 || x_cstride % 4 || x_cstride < Cin || Cout % 4) return 0;
  if (y_cstride % 4 || y_coff % 4) return 0;
  if (N > (1 << 24) || Hin > 32768 || Win > 32768) return 0;
  return 1;
}

// wmode 0: w_tc packed [taps][Cout][Cin].  wmode 1 / 2: w_tc is the weight-gradient layout
// [taps][w_rows][w_cols] of the conv's weight (rows = its input channels, cols = its output
// channels; w_rows_full >= rows actually used = row pitch of a tap): 1 = forward (Cin rows used,
// Cout == w_cols), 2 = data gradient (conv-Cin == w_cols, conv-Cout rows used).
// math: SG2IM_MATH_TF32 / SG2IM_MATH_BF16X3 / SG2IM_MATH_BF16 (include/sg2im_b200.h).
static int conv_tc_impl(const float* x, int64_t x_cstride, int64_t N, int64_t Hin,
                        int64_t Win, int64_t Cin, const float* w_tc, const float* bias,
                        int KH, int KW, int P, int64_t Hout, int64_t Wout, int64_t Cout,
                        int act, float slope, float* y, int64_t y_cstride, int64_t y_coff,
                        double* stats, int round_out, sg2im_stream_t stream, int wmode,
                        int64_t w_rows_full, int math, int64_t ps_pitch = 0) {
  SG_ARG(x && w_tc && y);
  SG_ARG(ps_pitch == 0 || (wmode == 0 && math != SG2IM_MATH_TF32 && ps_pitch % 32 == 0 &&
                           ps_pitch >= Cin && w_rows_full >= Cout));
  SG_ARG(math == SG2IM_MATH_TF32 || math == SG2IM_MATH_BF16X3 || math == SG2IM_MATH_BF16);
  if (!sg2im_conv_tc_supported(N, Hin, Win, Cin, x_cstride, KH, KW, 1, P, Hout, Wout, Cout,
                               y_cstride, y_coff)) {
    sg2im_set_error("sg2im_conv_tc: unsupported shape (use sg2im_conv_igemm)");
    return -2;
  }
  SG_ARG(aligned16(x) && aligned16(w_tc) && aligned16(y) && (!bias || aligned16(bias)));
  SG_ARG(stats == nullptr || (Cout <= 1024 && act == 0));
  EncodeTiledFn enc = get_encode();
  if (!enc) { sg2im_set_error("sg2im_conv_tc: cuTensorMapEncodeTiled unavailable"); return -3; }
  const int bf = math != SG2IM_MATH_TF32;
  const int nprod = math == SG2IM_MATH_BF16X3 ? 3 : 1;
  if (bf) round_out = 0;                  // RN-TF32 hand-over is a tf32-mode contract

  TcParams p;
  p.stats = stats; p.round_out = round_out; p.nprod = nprod;
  p.N = (int)N; p.Hout = (int)Hout; p.Wout = (int)Wout;
  p.Cin = (int)Cin; p.Cout = (int)Cout; p.KH = KH; p.KW = KW; p.P = P;
  tc_geometry(p.Hout, p.Wout, p.BW, p.BH, p.BI);
  p.tiles_w = (int)ceil_div64(p.Wout, p.BW); p.tiles_h = (int)ceil_div64(p.Hout, p.BH);
  p.tiles_n = (int)ceil_div64(N, p.BI);
  // N tile: the widest of 256/128/64 whose padding waste (last tile may be
  // partial: TMA zero-fills the missing weight rows, the epilogue skips the
  // columns) stays under 1/8, then narrowed while there is less than a wave.
  int BN = 64;
  for (int cand = 256; cand >= 64; cand >>= 1) {
    long long padded = ceil_div64(Cout, cand) * cand;
    if ((padded - Cout) * 8 <= padded) { BN = cand; break; }
  }
  long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
  if (bf) {
    // bf16 arithmetic: a pipeline stage costs ~700 cycles of L2 -> SM latency / converter round trip
    // plus ~1.3 cycles per output column whatever the N tile (measured on 1024 -> 1024 at 8x8:
    // 806 / 890 / 1041 cycles per stage for N = 64 / 128 / 256), so the widest tile wins unless it
    // leaves SMs idle: minimise waves x stage cost over the admissible tiles (8x8 maps: N = 128,
    // 133 us instead of 242 us with N = 64)
    const int bn_max = BN;
    double best = 1e30;
    for (int cand = bn_max; cand >= 64; cand >>= 1) {
      const long long tiles = m_tiles * ceil_div64(Cout, cand);
      const double cost = (double)ceil_div64(tiles, num_sms()) * (700.0 + 1.3 * cand);
      if (cost < best) { best = cost; BN = cand; }
    }
  } else {
    while (BN > 64 && m_tiles * ceil_div64(Cout, BN) < num_sms()) BN >>= 1;
  }
  // tuning aid (tools/prof_conv.py): SG2IM_TC_BN=64|128|256 pins the N tile
  if (const char* e = getenv("SG2IM_TC_BN")) {
    int forced = atoi(e);
    if (forced == 64 || forced == 128 || forced == 256) BN = forced;
  }
  p.n_tiles = (int)ceil_div64(Cout, BN);
  p.cblocks = (int)ceil_div64(Cin, 32);
  p.num_kb = KH * KW * p.cblocks;
  p.bias = bias; p.act = act; p.slope = slope;
  p.y = y; p.y_cstride = y_cstride; p.y_coff = y_coff;
  cudaStream_t st = as_stream(stream);

  // narrow outputs on real images: halo + weight-stationary kernel
  // (tf32: only where the per-tap kernel would use N <= 128 tiles; bf16 arithmetic: every such shape —
  // the per-tap kernel moves 2.3x the L2 -> SM bytes per FLOP even with N = 256 tiles and sits on the
  // ~6.6 TB/s the L2 fabric delivers, profiles/r02_call3_conv_kernels_bf16x3_ncu.txt)
  if (KH * KW > 1 && KH <= 3 && KW <= 3 && Hout >= H_BH && Wout >= H_BW && (BN <= 128 || bf) &&
      getenv("SG2IM_NO_HALO") == nullptr) {
    HaloParams h;
    h.N = (int)N; h.Hout = (int)Hout; h.Wout = (int)Wout; h.Cin = (int)Cin; h.Cout = (int)Cout;
    h.KH = KH; h.KW = KW; h.P = P; h.taps = KH * KW; h.pitch = H_BW + KW - 1;
    h.tiles_w = (int)ceil_div64(Wout, H_BW); h.tiles_h = (int)ceil_div64(Hout, H_BH);
    h.ptiles = (int)(N * h.tiles_h * h.tiles_w);
    // Cout tile: 128 for outputs at least 128 channels wide when the weights arrive pre-split (its single
    // weight set leaves no slack for an in-kernel split of 9 x 128 weight rows per channel block:
    // measured slower than the 64-wide tile then); SG2IM_HALO_BN=64|128 pins it
    int HB = (ps_pitch && Cout >= 128) ? 128 : 64;
    if (const char* e = getenv("SG2IM_HALO_BN")) { int f = atoi(e); if (f == 64 || (f == 128 && bf)) HB = f; }
    const int HT = HB == 64 ? 4 : 2;
    h.groups = (int)ceil_div64(h.ptiles, HT);
    h.n_tiles = (int)ceil_div64(Cout, HB);
    h.cblocks = (int)ceil_div64(Cin, 32);
    h.a_bytes = (uint32_t)((H_BH + KH - 1) * h.pitch * 128);
    h.bias = bias; h.act = act; h.slope = slope;
    h.y = y; h.y_cstride = y_cstride; h.y_coff = y_coff;
    h.stats = stats; h.round_out = round_out; h.nprod = nprod;
    CUtensorMap hA, hB;
    {
      cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)N};
      cuuint64_t gstr[3] = {(cuuint64_t)x_cstride * 4, (cuuint64_t)Win * x_cstride * 4,
                            (cuuint64_t)Hin * Win * x_cstride * 4};
      cuuint32_t box[4] = {32, (cuuint32_t)h.pitch, (cuuint32_t)(H_BH + KH - 1), 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = enc(&hA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gdim, gstr,
                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { sg2im_set_error("sg2im_conv_tc: encode halo A failed (%d)", (int)r); return -4; }
    }
    if (int rc = encode_weights(enc, &hB, w_tc, Cin, Cout, KH * KW, HB, wmode, w_rows_full, bf, ps_pitch)) return rc;
    int rc;
    if (HB == 128) {
      if (ps_pitch) rc = launch_halo<0, 1, 1, 128>(hA, hB, h, st);
      else rc = wmode == 1 ? launch_halo<1, 1, 0, 128>(hA, hB, h, st)
              : wmode == 2 ? launch_halo<2, 1, 0, 128>(hA, hB, h, st) : launch_halo<0, 1, 0, 128>(hA, hB, h, st);
    } else if (ps_pitch) rc = launch_halo<0, 1, 1>(hA, hB, h, st);
    else if (bf) rc = wmode == 1 ? launch_halo<1, 1>(hA, hB, h, st) : wmode == 2 ? launch_halo<2, 1>(hA, hB, h, st)
                                                                           : launch_halo<0, 1>(hA, hB, h, st);
    else rc = wmode == 1 ? launch_halo<1, 0>(hA, hB, h, st) : wmode == 2 ? launch_halo<2, 0>(hA, hB, h, st)
                                                                         : launch_halo<0, 0>(hA, hB, h, st);
    if (rc) return rc;
    SG_LAUNCH_OK();
    return 0;
  }

  CUtensorMap tmA, tmB;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)x_cstride * 4, (cuuint64_t)Win * x_cstride * 4,
                          (cuuint64_t)Hin * Win * x_cstride * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)p.BW, (cuuint32_t)p.BH, (cuuint32_t)p.BI};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sg2im_set_error("sg2im_conv_tc: encode A failed (%d)", (int)r); return -4; }
  }
  if (int rc = encode_weights(enc, &tmB, w_tc, Cin, Cout, KH * KW, BN, wmode, w_rows_full, bf, ps_pitch)) return rc;
  int rc = 0;
  if (ps_pitch) {
    rc = BN == 256 ? launch<256, 0, 1, 1>(tmA, tmB, p, st)
       : BN == 128 ? launch<128, 0, 1, 1>(tmA, tmB, p, st) : launch<64, 0, 1, 1>(tmA, tmB, p, st);
  } else if (bf) {
    rc = BN == 256 ? launch_w<256, 1>(tmA, tmB, p, wmode, st)
       : BN == 128 ? launch_w<128, 1>(tmA, tmB, p, wmode, st) : launch_w<64, 1>(tmA, tmB, p, wmode, st);
  } else {
    rc = BN == 256 ? launch_w<256, 0>(tmA, tmB, p, wmode, st)
       : BN == 128 ? launch_w<128, 0>(tmA, tmB, p, wmode, st) : launch_w<64, 0>(tmA, tmB, p, wmode, st);
  }
  if (rc) return rc;
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_conv_tc(const float* x, int64_t x_cstride, int64_t N, int64_t Hin,
                             int64_t Win, int64_t Cin, const float* w_tc, const float* bias,
                             int KH, int KW, int P, int64_t Hout, int64_t Wout, int64_t Cout,
                             int act, float slope, float* y, int64_t y_cstride, int64_t y_coff,
                             double* stats, int round_out, int math, sg2im_stream_t stream) {
  return conv_tc_impl(x, x_cstride, N, Hin, Win, Cin, w_tc, bias, KH, KW, P, Hout, Wout, Cout, act,
                      slope, y, y_cstride, y_coff, stats, round_out, stream, 0, 0, math);
}

extern "C" int sg2im_conv_tc_kcc(const float* x, int64_t x_cstride, int64_t N, int64_t Hin,
                                 int64_t Win, int64_t Cin, const float* w_kcc, int64_t w_rows_full,
                                 int dgrad, const float* bias, int KH, int KW, int P, int64_t Hout,
                                 int64_t Wout, int64_t Cout, int act, float slope, float* y,
                                 int64_t y_cstride, int64_t y_coff, double* stats, int round_out,
                                 int math, sg2im_stream_t stream) {
  SG_ARG(w_rows_full >= (dgrad ? Cout : Cin));
  return conv_tc_impl(x, x_cstride, N, Hin, Win, Cin, w_kcc, bias, KH, KW, P, Hout, Wout, Cout, act,
                      slope, y, y_cstride, y_coff, stats, round_out, stream, dgrad ? 2 : 1,
                      w_rows_full, math);
}

extern "C" int sg2im_conv_tc_presplit(const float* x, int64_t x_cstride, int64_t N, int64_t Hin,
                                      int64_t Win, int64_t Cin, const float* w_split, int64_t w_pitch,
                                      int64_t w_rows_per_tap, const float* bias, int KH, int KW, int P,
                                      int64_t Hout, int64_t Wout, int64_t Cout, int act, float slope,
                                      float* y, int64_t y_cstride, int64_t y_coff, double* stats,
                                      int math, sg2im_stream_t stream) {
  SG_ARG(math == SG2IM_MATH_BF16X3 || math == SG2IM_MATH_BF16);
  SG_ARG(w_pitch > 0);
  return conv_tc_impl(x, x_cstride, N, Hin, Win, Cin, w_split, bias, KH, KW, P, Hout, Wout, Cout, act,
                      slope, y, y_cstride, y_coff, stats, 0, stream, 0, w_rows_per_tap, math, w_pitch);
}
