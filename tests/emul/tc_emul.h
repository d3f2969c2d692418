// TEST INFRASTRUCTURE ONLY — a FUNCTIONAL model of the Blackwell primitives the tensor-core
// kernels use (csrc/tc_common.cuh wrappers): TMA tiled loads with shared-memory swizzle and
// out-of-bounds zero fill, mbarriers (arrive / expect_tx / complete_tx / parity wait),
// tcgen05.alloc / mma kind::tf32 (K-major and MN-major shared-memory descriptors) / commit /
// ld, TMEM, and thread-block clusters (rank, barrier, multicast TMA, multicast commit).
// The point is to execute the kernels' OWN control flow, pipeline protocol, descriptor arithmetic
// and indexing on the CPU.  No timing; asynchrony is modelled (see `AsyncState` below): loads and
// tensor-pipe work complete some scheduler sweeps after they are issued, in-flight TMA destinations
// are poisoned, so missing waits and early slot releases produce wrong numbers.
//
// The model encodes this repository's understanding of the hardware (probed on the B200 with
// tools/umma_probe*.cu, see profiles/README.md) and is CALIBRATED by the kernels that are proven
// on hardware: conv_tc / conv_tc_halo / conv_wgrad_tc must reproduce correct convolutions under
// it (tests/test_tc_kernels_emulated_cpu.py) before it is used to exercise kernels that have not
// run on hardware yet (conv_wgrad_tc_mc).  Facts the model relies on:
//   * the 128-byte swizzles are XORs on ABSOLUTE shared-memory address bits (so row-shifted
//     operand start addresses stay consistent with what TMA wrote);
//   * K-major SWIZZLE_128B operand, tf32: element (row r, k) at start + (r/8)*SBO + (r%8)*128 + 4k;
//   * MN-major SWIZZLE_128B_BASE32B operand, tf32: element (mn, k) at
//     start + (mn/32)*LBO + (k/4)*SBO + (k%4)*128 + 4*(mn%32);
//   * tcgen05.mma kind::tf32 ignores the 13 low mantissa bits of its operands, accumulates fp32;
//   * kind::f16 with bf16 operands (K = 16): K-major SWIZZLE_128B element (row r, k) at
//     start + (r/8)*SBO + (r%8)*128 + 2k; MN-major SWIZZLE_128B element (mn, k) at
//     start + (mn/64)*LBO + (k/8)*SBO + (k%8)*128 + 2*(mn%64) (CUTLASS's canonical layouts,
//     cute/atom/mma_traits_sm100.hpp; the 16-bit MN-major form is pinned on hardware by
//     tools/umma_probe_bf16.cu), both through the 16-byte-chunk XOR of SWIZZLE_128B;
//   * accumulator D[row][col] lives at TMEM lane `row`, column `d_tmem.col + col`.
// Only the exact XOR pattern of SWIZZLE_128B_ATOM_32B is an assumption; it cancels out because
// producer (TMA) and consumer (MMA) use the same function.
#pragma once
#ifdef SG2IM_EMUL_THREADS
#error "the tensor-core model needs the fiber execution model (no -DSG2IM_EMUL_THREADS)"
#endif
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <functional>
#include <vector>

#define __grid_constant__

static inline long long clock64() { static long long t = 0; return t += 1000; }
[[noreturn]] static inline void __trap() {
  std::fprintf(stderr, "emulated kernel: __trap() (an mbarrier wait never completed)\n");
  std::abort();
}
static inline float4 atomicAdd(float4* p, float4 v) {
  float4 old = *p;
  p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w;
  return old;
}

namespace emul {
constexpr uint32_t SMEM_VA = 0x400;                      // shared window base (1 KB aligned)
struct TMap {                                            // lives inside the 128-byte CUtensorMap
  void* gptr;
  uint32_t rank, swizzle;
  uint64_t dim[5];
  uint64_t stride[5];                                    // bytes; stride[0] = element size
  uint32_t box[5];
};
static_assert(sizeof(TMap) <= sizeof(CUtensorMap), "descriptor does not fit");

inline uint32_t va_of(const Block* b, const void* p) {
  return (uint32_t)(reinterpret_cast<const unsigned char*>(p) - b->smem) + SMEM_VA;
}
inline unsigned char* host_of(Block* b, uint32_t va) { return b->smem + (va - SMEM_VA); }
inline uint32_t swz128(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }
inline uint32_t swz128_atom32(uint32_t a) { return a ^ (((a >> 7) & 3u) << 5); }
inline float tf32_trunc(float f) {
  uint32_t u; std::memcpy(&u, &f, 4); u &= ~0x1fffu; std::memcpy(&f, &u, 4); return f;
}
inline MBar& bar_at(Block* b, uint32_t va) { return b->mbar[va]; }
inline void bar_check_complete(MBar& m) {
  if (m.inited && m.pending == 0 && m.tx == 0) { m.phase ^= 1u; m.pending = m.init_count; }
}
inline void bar_arrive(Block* b, uint32_t va) {
  MBar& m = bar_at(b, va);
  if (!m.inited || m.pending == 0) { std::fprintf(stderr, "emul: arrive on bad mbarrier %x\n", va); std::abort(); }
  --m.pending;
  bar_check_complete(m);
}
inline void bar_complete_tx(Block* b, uint32_t va, long long bytes) {
  MBar& m = bar_at(b, va);
  m.tx -= bytes;
  bar_check_complete(m);
}
// ---- asynchrony (SG2IM_EMUL_ASYNC=<sweeps>, default 3; 0 = everything completes at issue).
// A TMA load poisons its destination with NaNs when it is ISSUED and delivers data + complete_tx a
// few scheduler sweeps later (pseudo-random per load, so loads also complete out of order); MMAs and
// commits go through an in-order queue per CTA and read their shared-memory operands when they
// EXECUTE.  A consumer that does not wait for `full`, a producer that refills a slot before the
// MMAs reading it have been committed (`empty`), an epilogue that reads TMEM before the commit —
// all of them now compute with NaNs / stale data and fail the comparison instead of passing by luck
// of the cooperative schedule.
struct AsyncOp { unsigned long long ready; std::function<void()> fn; };
struct AsyncState {
  unsigned long long now = 0;
  int delay = -1;
  uint32_t lcg = 12345u;
  std::vector<AsyncOp> tma;                            // complete individually, in any order
  std::deque<AsyncOp> pipe[16];                        // tensor pipe of CTA `rank`: in order
};
inline AsyncState& async() { static AsyncState s; return s; }
inline void async_tick() {
  AsyncState& A = async();
  ++A.now;
  for (size_t i = 0; i < A.tma.size();) {
    if (A.tma[i].ready <= A.now) { auto fn = std::move(A.tma[i].fn); A.tma.erase(A.tma.begin() + (long)i); fn(); }
    else ++i;
  }
  for (auto& q : A.pipe)
    while (!q.empty() && q.front().ready <= A.now) { auto fn = std::move(q.front().fn); q.pop_front(); fn(); }
}
inline void async_drain() {
  AsyncState& A = async();
  if (!A.tma.empty()) { std::fprintf(stderr, "emul: a CTA exited with a TMA load still in flight\n"); std::abort(); }
  for (auto& q : A.pipe) { while (!q.empty()) { auto fn = std::move(q.front().fn); q.pop_front(); fn(); } }
}
inline int async_delay() {
  AsyncState& A = async();
  if (A.delay < 0) {
    const char* e = std::getenv("SG2IM_EMUL_ASYNC");
    A.delay = e ? std::atoi(e) : 3;
    if (A.delay > 0) { sweep_hook() = &async_tick; drain_hook() = &async_drain; }
  }
  return A.delay;
}
inline unsigned long long async_when(int spread) {      // now + delay + pseudo-random 0..spread-1
  AsyncState& A = async();
  A.lcg = A.lcg * 1664525u + 1013904223u;
  return A.now + (unsigned long long)A.delay + (spread > 1 ? (A.lcg >> 16) % (unsigned)spread : 0u);
}
inline void pipe_push(unsigned rank, std::function<void()> fn) {
  if (async_delay() <= 0) { fn(); return; }
  // adversarial schedule: SG2IM_EMUL_SLOW_PIPE=<sweeps> makes the tensor pipes of CTAs of rank >= 1 lag (a
  // multicasting producer that does not collect every peer's release then overwrites operands in use)
  const char* sp = std::getenv("SG2IM_EMUL_SLOW_PIPE");
  const unsigned long long lag = (sp && rank >= 1) ? (unsigned long long)std::atoi(sp) : 0ull;
  async().pipe[rank & 15].push_back(AsyncOp{async_when(1) + lag, std::move(fn)});
}

// copy one TMA box into CTA `b` at shared address dst_va (swizzled), signal its barrier
inline void tma_box_now(Block* b, uint32_t dst_va, const TMap& t, uint32_t bar_va, const int* c, bool poison) {
  if (t.stride[0] != 4 || t.box[0] * 4 != 128) { std::fprintf(stderr, "emul: TMA box must be 32 floats wide\n"); std::abort(); }
  long long bytes = 0;
  const uint32_t b1 = t.rank > 1 ? t.box[1] : 1, b2 = t.rank > 2 ? t.box[2] : 1, b3 = t.rank > 3 ? t.box[3] : 1;
  for (uint32_t i3 = 0; i3 < b3; ++i3)
    for (uint32_t i2 = 0; i2 < b2; ++i2)
      for (uint32_t i1 = 0; i1 < b1; ++i1) {
        const uint32_t row = (i3 * b2 + i2) * b1 + i1;
        long long g1 = t.rank > 1 ? (long long)c[1] + i1 : 0, g2 = t.rank > 2 ? (long long)c[2] + i2 : 0,
                  g3 = t.rank > 3 ? (long long)c[3] + i3 : 0;
        const bool row_in = g1 >= 0 && (t.rank < 2 || g1 < (long long)t.dim[1]) && g2 >= 0 &&
                            (t.rank < 3 || g2 < (long long)t.dim[2]) && g3 >= 0 &&
                            (t.rank < 4 || g3 < (long long)t.dim[3]);
        for (uint32_t i0 = 0; i0 < t.box[0]; ++i0) {
          const long long g0 = (long long)c[0] + i0;
          float v = 0.f;
          if (row_in && g0 >= 0 && g0 < (long long)t.dim[0]) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(t.gptr) + g0 * 4 +
                                       (t.rank > 1 ? g1 * (long long)t.stride[1] : 0) +
                                       (t.rank > 2 ? g2 * (long long)t.stride[2] : 0) +
                                       (t.rank > 3 ? g3 * (long long)t.stride[3] : 0);
            std::memcpy(&v, src, 4);
          }
          uint32_t a = dst_va + row * 128u + i0 * 4u;
          a = t.swizzle == (uint32_t)CU_TENSOR_MAP_SWIZZLE_128B ? swz128(a)
              : t.swizzle == (uint32_t)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B ? swz128_atom32(a) : a;
          if (a - SMEM_VA + 4 > b->smem_bytes) { std::fprintf(stderr, "emul: TMA write outside shared memory\n"); std::abort(); }
          if (poison) { const uint32_t nan = 0x7fc0deadu; std::memcpy(host_of(b, a), &nan, 4); }
          else std::memcpy(host_of(b, a), &v, 4);
          bytes += 4;
        }
      }
  if (!poison) bar_complete_tx(b, bar_va, bytes);
}
inline void tma_box(Block* b, uint32_t dst_va, const TMap& t, uint32_t bar_va, const int* c) {
  if (async_delay() <= 0) { tma_box_now(b, dst_va, t, bar_va, c, false); return; }
  tma_box_now(b, dst_va, t, bar_va, c, true);           // the destination is undefined from now on ...
  const int c4[4] = {c[0], c[1], c[2], c[3]};
  // loads into the higher-ranked CTAs of a cluster land later (SG2IM_EMUL_ASYNC_SKEW sweeps per rank,
  // default 6): a CTA that consumes a peer's shared memory without having been told that the
  // peer's load completed reads the poison
  static const int skew = std::getenv("SG2IM_EMUL_ASYNC_SKEW") ? std::atoi(std::getenv("SG2IM_EMUL_ASYNC_SKEW")) : 6;
  // adversarial schedules for cluster kernels (tests sweep them): SG2IM_EMUL_ASYNC_SLOW3D=<sweeps> delays
  // the rank-3 tensor-map loads (the weight tiles) into CTAs of rank >= 1 further
  const char* s3 = std::getenv("SG2IM_EMUL_ASYNC_SLOW3D");
  const unsigned long long slow3d = (s3 && t.rank == 3 && b->rank >= 1) ? (unsigned long long)std::atoi(s3) : 0ull;
  async().tma.push_back(AsyncOp{async_when(3) + (unsigned long long)(skew * (int)b->rank) + slow3d,
                                [=]() { tma_box_now(b, dst_va, t, bar_va, c4, false); }});
}
}  // namespace emul

// cuTensorMapEncodeTiled stand-in (tc::get_encode under SG2IM_EMUL)
static inline CUresult emul_tensor_map_encode_tiled(
    CUtensorMap* map, CUtensorMapDataType dtype, cuuint32_t rank, void* gaddr, const cuuint64_t* gdim,
    const cuuint64_t* gstride, const cuuint32_t* box, const cuuint32_t* estride, CUtensorMapInterleave,
    CUtensorMapSwizzle swizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  if (dtype != CU_TENSOR_MAP_DATA_TYPE_FLOAT32 || rank < 1 || rank > 5) return CUDA_ERROR_INVALID_VALUE;
  if ((reinterpret_cast<uintptr_t>(gaddr) & 15) != 0) return CUDA_ERROR_INVALID_VALUE;
  emul::TMap t{};
  t.gptr = gaddr; t.rank = rank; t.swizzle = (uint32_t)swizzle;
  t.stride[0] = 4;
  for (uint32_t i = 0; i < rank; ++i) {
    t.dim[i] = gdim[i]; t.box[i] = box[i];
    if (estride[i] != 1 || box[i] == 0 || box[i] > 256) return CUDA_ERROR_INVALID_VALUE;
    if (i > 0) {
      t.stride[i] = gstride[i - 1];
      if (gstride[i - 1] % 16) return CUDA_ERROR_INVALID_VALUE;     // the driver's rule
    }
  }
  std::memset(map, 0, sizeof(*map));
  std::memcpy(map, &t, sizeof(t));
  return CUDA_SUCCESS;
}

namespace tc {
using emul::Block;

static inline uint32_t smem_u32(const void* p) { return emul::va_of(emul::current(), p); }
static inline void mbar_init(uint64_t* bar, uint32_t count) {
  emul::MBar& m = emul::bar_at(emul::current(), smem_u32(bar));
  m = emul::MBar{};
  m.init_count = m.pending = count; m.inited = true;
}
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {          // arrive.expect_tx
  Block* b = emul::current();
  emul::MBar& m = emul::bar_at(b, smem_u32(bar));
  m.tx += bytes;
  emul::bar_arrive(b, smem_u32(bar));
}
static inline void mbar_arrive(uint64_t* bar) { emul::bar_arrive(emul::current(), smem_u32(bar)); }
static inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  emul::MBar& m = emul::bar_at(emul::current(), smem_u32(bar));
  if (!m.inited) { std::fprintf(stderr, "emul: wait on uninitialised mbarrier\n"); std::abort(); }
  if (m.phase != parity) return true;                   // the phase with this parity has completed
  emul::yield();
  return false;
}
static inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > 50000000LL) __trap();                 // a protocol bug must fail, never hang
}
static inline void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned) { mbar_wait(bar, parity); }
static inline void tma_load_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  emul::TMap t; std::memcpy(&t, map, sizeof(t));
  const int c[4] = {c0, c1, c2, c3};
  emul::tma_box(emul::current(), smem_u32(smem), t, smem_u32(bar), c);
}
static inline void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  emul::TMap t; std::memcpy(&t, map, sizeof(t));
  const int c[4] = {c0, c1, c2, 0};
  emul::tma_box(emul::current(), smem_u32(smem), t, smem_u32(bar), c);
}
static inline void tma_load_4d_mc(void* smem, const CUtensorMap* map, uint64_t* bar, uint16_t mask,
                                  int c0, int c1, int c2, int c3) {
  emul::TMap t; std::memcpy(&t, map, sizeof(t));
  const int c[4] = {c0, c1, c2, c3};
  emul::Gang* G = emul::gang();
  const uint32_t dst = smem_u32(smem), bv = smem_u32(bar);
  for (unsigned r = 0; r < G->blocks.size(); ++r)
    if (mask & (1u << r)) emul::tma_box(&G->blocks[r], dst, t, bv, c);
}
static inline void tma_prefetch_desc(const CUtensorMap*) {}
static inline void mbar_fence_init() {}
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}

static inline void tc_alloc(uint32_t* slot, uint32_t ncols) {                // whole warp calls it
  Block* b = emul::current();
  if (ncols < 32 || ncols > 512 || (ncols & (ncols - 1))) { std::fprintf(stderr, "emul: bad TMEM column count\n"); std::abort(); }
  if (b->tmem.empty()) b->tmem.assign(128 * 512, 0.f);
  *slot = 0;                                             // lane 0, column 0
}
static inline void tc_dealloc(uint32_t, uint32_t) {}
// commits arrive once every MMA issued before them by this CTA has executed (in-order pipe)
static inline void tc_commit(uint64_t* bar, uint32_t = 1u) {                 // converged warp, one arrive
  if (emul_lane() != 0) return;
  Block* b = emul::current();
  const uint32_t bv = smem_u32(bar);
  emul::pipe_push(b->rank, [=]() { emul::bar_arrive(b, bv); });
}
static inline void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  if (emul_lane() != 0) return;
  emul::Gang* G = emul::gang();
  const uint32_t bv = smem_u32(bar);
  emul::pipe_push(emul::current()->rank, [=]() {
    for (unsigned r = 0; r < G->blocks.size(); ++r)
      if (mask & (1u << r)) emul::bar_arrive(&G->blocks[r], bv);
  });
}

// one operand element through a shared-memory matrix descriptor, read from `blk`'s shared memory
static inline float emul_desc_elem(Block* blk, uint32_t lo, uint32_t hi, bool mn_major, int idx, int k) {
  const uint32_t start = (lo & 0x3fffu) << 4, lbo = ((lo >> 16) & 0x3fffu) << 4, sbo = (hi & 0x3fffu) << 4;
  const uint32_t layout = (hi >> 29) & 7u;
  uint32_t a;
  if (!mn_major) {
    if (layout != 2u) { std::fprintf(stderr, "emul: K-major operand must be SWIZZLE_128B\n"); std::abort(); }
    a = emul::swz128(start + (uint32_t)(idx >> 3) * sbo + (uint32_t)(idx & 7) * 128u + (uint32_t)k * 4u);
  } else {
    if (layout != 1u) { std::fprintf(stderr, "emul: MN-major tf32 operand must be SWIZZLE_128B_BASE32B\n"); std::abort(); }
    a = emul::swz128_atom32(start + (uint32_t)(idx >> 5) * lbo + (uint32_t)(k >> 2) * sbo +
                            (uint32_t)(k & 3) * 128u + (uint32_t)(idx & 31) * 4u);
  }
  if (a < emul::SMEM_VA || a - emul::SMEM_VA + 4 > blk->smem_bytes) { std::fprintf(stderr, "emul: MMA operand read outside shared memory\n"); std::abort(); }
  float v; std::memcpy(&v, emul::host_of(blk, a), 4);
  return emul::tf32_trunc(v);
}

// one bf16 operand element (kind::f16) through a shared-memory matrix descriptor
static inline float emul_desc_elem16(Block* blk, uint32_t lo, uint32_t hi, bool mn_major, int idx, int k) {
  const uint32_t start = (lo & 0x3fffu) << 4, lbo = ((lo >> 16) & 0x3fffu) << 4, sbo = (hi & 0x3fffu) << 4;
  const uint32_t layout = (hi >> 29) & 7u;
  if (layout != 2u) { std::fprintf(stderr, "emul: bf16 operand must be SWIZZLE_128B\n"); std::abort(); }
  uint32_t a;
  if (!mn_major)
    a = emul::swz128(start + (uint32_t)(idx >> 3) * sbo + (uint32_t)(idx & 7) * 128u + (uint32_t)k * 2u);
  else
    a = emul::swz128(start + (uint32_t)(idx >> 6) * lbo + (uint32_t)(k >> 3) * sbo +
                     (uint32_t)(k & 7) * 128u + (uint32_t)(idx & 63) * 2u);
  if (a < emul::SMEM_VA || a - emul::SMEM_VA + 2 > blk->smem_bytes) { std::fprintf(stderr, "emul: MMA operand read outside shared memory\n"); std::abort(); }
  uint16_t h; std::memcpy(&h, emul::host_of(blk, a), 2);
  uint32_t u = (uint32_t)h << 16;
  float v; std::memcpy(&v, &u, 4);
  return v;
}
static inline void emul_mma1_f16(Block* blk, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                 uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  const int N = (int)((idesc >> 17) & 0x3f) << 3, M = (int)((idesc >> 24) & 0x1f) << 4;
  const bool a_mn = (idesc >> 15) & 1u, b_mn = (idesc >> 16) & 1u;
  if (M != 128 || N < 16 || N > 256 || (N & 15) || ((idesc >> 7) & 7u) != 1u || ((idesc >> 10) & 7u) != 1u ||
      ((idesc >> 4) & 3u) != 1u) {
    std::fprintf(stderr, "emul: unsupported kind::f16 instruction descriptor %x\n", idesc); std::abort();
  }
  const uint32_t lane0 = d_tmem >> 16, col0 = d_tmem & 0xffffu;
  if (lane0 != 0 || col0 + (uint32_t)N > 512u) { std::fprintf(stderr, "emul: accumulator outside TMEM\n"); std::abort(); }
  static float A[128][16], B[256][16];
  for (int m = 0; m < M; ++m) for (int k = 0; k < 16; ++k) A[m][k] = emul_desc_elem16(blk, a_lo, a_hi, a_mn, m, k);
  for (int n = 0; n < N; ++n) for (int k = 0; k < 16; ++k) B[n][k] = emul_desc_elem16(blk, b_lo, b_hi, b_mn, n, k);
  for (int m = 0; m < M; ++m) {
    float* drow = &blk->tmem[(size_t)m * 512 + col0];
    for (int n = 0; n < N; ++n) {
      float acc = accumulate ? drow[n] : 0.f;
      for (int k = 0; k < 16; ++k) acc += A[m][k] * B[n][k];
      drow[n] = acc;
    }
  }
}
static inline void tc_mma_f16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                 uint32_t b_hi, uint32_t idesc, uint32_t accumulate, uint32_t) {
  if (emul_lane() != 0) return;                           // elect.sync: one lane issues
  Block* blk = emul::current();
  emul::pipe_push(blk->rank, [=]() { emul_mma1_f16(blk, d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate); });
}
static inline float4 lds128(uint32_t addr) {             // ld.shared.v4.f32 at a shared-window address
  Block* b = emul::current();
  if (addr < emul::SMEM_VA || addr - emul::SMEM_VA + 16 > b->smem_bytes || (addr & 15u)) { std::fprintf(stderr, "emul: ld.shared outside shared memory / misaligned\n"); std::abort(); }
  float4 v; std::memcpy(&v, emul::host_of(b, addr), 16);
  return v;
}
static inline void sts128(uint32_t addr, uint32_t a, uint32_t c, uint32_t d, uint32_t e) {
  Block* b = emul::current();
  if (addr < emul::SMEM_VA || addr - emul::SMEM_VA + 16 > b->smem_bytes || (addr & 15u)) { std::fprintf(stderr, "emul: st.shared outside shared memory / misaligned\n"); std::abort(); }
  const uint32_t w[4] = {a, c, d, e};
  std::memcpy(emul::host_of(b, addr), w, 16);
}
static inline void fence_proxy_async() {}
// tcgen05.mma.cta_group::1.kind::tf32, descriptors given as (lo, hi) words: what the tensor pipe does
// when the instruction EXECUTES (operands are read from shared memory then, not at issue)
static inline void emul_mma1(Block* blk, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                             uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  const int N = (int)((idesc >> 17) & 0x3f) << 3, M = (int)((idesc >> 24) & 0x1f) << 4;
  const bool a_mn = (idesc >> 15) & 1u, b_mn = (idesc >> 16) & 1u;
  if (M != 128 || N < 8 || N > 256 || ((idesc >> 7) & 7u) != 2u || ((idesc >> 10) & 7u) != 2u) {
    std::fprintf(stderr, "emul: unsupported instruction descriptor %x\n", idesc); std::abort();
  }
  const uint32_t lane0 = d_tmem >> 16, col0 = d_tmem & 0xffffu;
  if (lane0 != 0 || col0 + (uint32_t)N > 512u) { std::fprintf(stderr, "emul: accumulator outside TMEM\n"); std::abort(); }
  float A[128][8], B[256][8];
  for (int m = 0; m < M; ++m) for (int k = 0; k < 8; ++k) A[m][k] = emul_desc_elem(blk, a_lo, a_hi, a_mn, m, k);
  for (int n = 0; n < N; ++n) for (int k = 0; k < 8; ++k) B[n][k] = emul_desc_elem(blk, b_lo, b_hi, b_mn, n, k);
  for (int m = 0; m < M; ++m) {
    float* drow = &blk->tmem[(size_t)m * 512 + col0];
    for (int n = 0; n < N; ++n) {
      float acc = accumulate ? drow[n] : 0.f;
      for (int k = 0; k < 8; ++k) acc += A[m][k] * B[n][k];
      drow[n] = acc;
    }
  }
}

// ---- CTA pair (cta_group::2).  ASSUMED semantics (CUTLASS's 2x1SM atoms; to be pinned on hardware
// by tools/umma_2cta_probe.cu before any kernel built on them is trusted): M = 256, the SAME
// descriptors are applied to both CTAs' shared memory; CTA r of the pair supplies rows
// 128r..128r+127 of A and columns (N/2)r..(N/2)(r+1)-1 of B (its descriptor addresses an N/2-wide
// tile); the accumulator rows 128r.. land in CTA r's TMEM at the same lane / column address.
// Issued by the even-ranked CTA only.
static inline void tc_mma_tf32_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                  uint32_t b_hi, uint32_t idesc, uint32_t accumulate, uint32_t) {
  if (emul_lane() != 0) return;                           // elect.sync: one lane issues
  Block* blk = emul::current();
  emul::pipe_push(blk->rank, [=]() { emul_mma1(blk, d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate); });
}
static inline void emul_mma2(emul::Gang* G, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                             uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  if (G->blocks.size() != 2) { std::fprintf(stderr, "emul: cta_group::2 MMA needs a cluster of 2, issued by rank 0\n"); std::abort(); }
  const int N = (int)((idesc >> 17) & 0x3f) << 3, M = (int)((idesc >> 24) & 0x1f) << 4;
  const bool a_mn = (idesc >> 15) & 1u, b_mn = (idesc >> 16) & 1u;
  if (M != 256 || N < 16 || N > 256 || (N & 15) || ((idesc >> 7) & 7u) != 2u || ((idesc >> 10) & 7u) != 2u) {
    std::fprintf(stderr, "emul: unsupported pair instruction descriptor %x\n", idesc); std::abort();
  }
  const uint32_t lane0 = d_tmem >> 16, col0 = d_tmem & 0xffffu;
  if (lane0 != 0 || col0 + (uint32_t)N > 512u) { std::fprintf(stderr, "emul: accumulator outside TMEM\n"); std::abort(); }
  static float B[256][8];
  for (int n = 0; n < N; ++n) {
    Block* src = &G->blocks[n < N / 2 ? 0 : 1];
    for (int k = 0; k < 8; ++k) B[n][k] = emul_desc_elem(src, b_lo, b_hi, b_mn, n % (N / 2), k);
  }
  for (int r = 0; r < 2; ++r) {
    Block* blk = &G->blocks[r];
    if (blk->tmem.empty()) { std::fprintf(stderr, "emul: pair MMA before both CTAs allocated TMEM\n"); std::abort(); }
    for (int m = 0; m < 128; ++m) {
      float a[8];
      for (int k = 0; k < 8; ++k) a[k] = emul_desc_elem(blk, a_lo, a_hi, a_mn, m, k);
      float* drow = &blk->tmem[(size_t)m * 512 + col0];
      for (int n = 0; n < N; ++n) {
        float acc = accumulate ? drow[n] : 0.f;
        for (int k = 0; k < 8; ++k) acc += a[k] * B[n][k];
        drow[n] = acc;
      }
    }
  }
}
static inline void tc_mma2_tf32_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                   uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  if (emul_lane() != 0) return;
  emul::Gang* G = emul::gang();
  if (emul::current()->rank != 0) { std::fprintf(stderr, "emul: cta_group::2 MMA issued by rank 1\n"); std::abort(); }
  emul::pipe_push(0, [=]() { emul_mma2(G, d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate); });
}
static inline void tc_alloc2(uint32_t* slot, uint32_t ncols) { tc_alloc(slot, ncols); }   // warp 0 of both CTAs
static inline void tc_dealloc2(uint32_t, uint32_t) {}
static inline void tc_commit2_mc(uint64_t* bar, uint16_t mask) { tc_commit_mc(bar, mask); }
// mbarrier.arrive on the barrier at this offset in CTA `rank` of the cluster (mapa + shared::cluster)
static inline void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  emul::Gang* G = emul::gang();
  if (rank >= G->blocks.size()) { std::fprintf(stderr, "emul: remote arrive outside the cluster\n"); std::abort(); }
  emul::bar_arrive(&G->blocks[rank], smem_u32(bar));
}
static inline void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  tc_mma_tf32_lh(d_tmem, (uint32_t)adesc, (uint32_t)(adesc >> 32), (uint32_t)bdesc, (uint32_t)(bdesc >> 32), idesc, accumulate, 1u);
}
// tcgen05.ld.32x32b.x32: lane i of the warp reads TMEM lane (base + i), 32 consecutive columns
static inline void tc_ld32(uint32_t taddr, float* v) {
  Block* b = emul::current();
  // adversarial schedule: SG2IM_EMUL_SLOW_EPILOGUE=<sweeps> makes the epilogue warps of CTAs of rank >= 1
  // lag (an MMA issuer that does not wait for a peer's epilogue then overwrites live accumulators)
  if (b->rank >= 1 && emul::async_delay() > 0) {
    static int lag = -1;
    const char* e = std::getenv("SG2IM_EMUL_SLOW_EPILOGUE");
    lag = e ? std::atoi(e) : 0;
    for (int i = 0; i < lag; ++i) emul::yield();
  }
  const uint32_t lane = (taddr >> 16) + emul_lane(), col = taddr & 0xffffu;
  if (lane >= 128 || col + 32 > 512 || (emul_warp() & 3u) != ((taddr >> 16) >> 5)) {
    std::fprintf(stderr, "emul: tcgen05.ld outside the warp's TMEM lane quadrant\n"); std::abort();
  }
  for (int i = 0; i < 32; ++i) v[i] = b->tmem[(size_t)lane * 512 + col + i];
}

static inline uint32_t cluster_rank() { return emul::current()->rank; }
static inline void cluster_sync_all() { emul::cluster_sync(); }
}  // namespace tc
