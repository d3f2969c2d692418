// Hardware probe (not part of the library): tcgen05.mma.cta_group::2 (CTA pair, M = 256) for TF32.
//
// Why: the N = 64 tensor-core kernels run at ~63 cycles per 128 x 64 x 8 MMA against 32 cycles of
// math (profiles/r01_prof_conv_tc_halo.txt, r01_prof_conv_wgrad_tc_v2.txt) — consistent with a
// shared-memory operand feed of ~100 B/clk (4 KB of A + 2 KB of B per MMA).  A CTA pair shares B:
// each SM reads its own 128 rows of A and only HALF of the B tile, 5 KB instead of 6 KB at N = 64,
// 6 KB instead of 8 KB at N = 128.  Before the convolution kernels are rebuilt around pairs, this
// probe pins the semantics the rebuild depends on and measures what a pair buys:
//   part 1 (one cluster): D[256 x N] = A[256 x 32] * B[N x 32]^T with
//       - CTA r holding rows 128r..128r+127 of A and rows (N/2)r..(N/2)(r+1)-1 of B (K-major,
//         SWIZZLE_128B, written with plain stores in the swizzled order),
//       - TMEM allocated with alloc.cta_group::2 by warp 0 of BOTH CTAs,
//       - the MMAs issued by rank 0 only, completion multicast to both CTAs' mbarriers,
//       - every CTA reading its own 128 TMEM lanes;
//     the host checks the assumed placement (CTA r = rows 128r.., all N columns, B columns in rank
//     order) and, if it does not hold, prints where each CTA's values sit in the reference product;
//   part 2 (all SMs, 74 pairs): cycles per pair-MMA for N = 64 / 128 / 256 next to the single-CTA
//     numbers of tools/umma_rate_probe.cu.
// Every wait is bounded (trap after ~4 s); run under `timeout`.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I sg2im_b200/csrc \
//        -o /tmp/umma_2cta_probe tools/umma_2cta_probe.cu -lcuda && timeout 60 /tmp/umma_2cta_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include "tc_common.cuh"

using namespace tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

namespace {

__device__ __forceinline__ void alloc2(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)),
               "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dealloc2(uint32_t base, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
// converged warp, one elected lane issues
__device__ __forceinline__ void mma2_tf32(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                          uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;\n\t}" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// K-major SWIZZLE_128B tile of `rows` x 32 floats at a 1024-byte aligned base: row r = 128 bytes at
// r * 128, its eight 16-byte chunks permuted by r % 8 (XOR on address bits 4..6 with bits 7..9)
__device__ __forceinline__ void store_swizzled(uint8_t* base, int r, int k, float v) {
  uint32_t off = (uint32_t)r * 128u + (uint32_t)k * 4u;
  off ^= ((off >> 7) & 7u) << 4;
  *reinterpret_cast<float*>(base + off) = v;
}

struct RateCase { int n, accs, iters; };

// part 1 when rate == nullptr (one cluster), part 2 otherwise
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const float* __restrict__ A, const float* __restrict__ B, int N, float* __restrict__ D,
            const RateCase* rate, int nrate, long long* cycles) {
  extern __shared__ __align__(16) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                      // 128 rows x 128 B
  uint8_t* sB = smem + 16 * 1024;          // up to 128 rows x 128 B (N / 2 <= 128)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int half = N / 2;
  if (!rate) {
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x)
      store_swizzled(sA, i / 32, i % 32, A[((size_t)rank * 128 + i / 32) * 32 + i % 32]);
    for (int i = threadIdx.x; i < half * 32; i += blockDim.x)
      store_swizzled(sB, i / 32, i % 32, B[((size_t)rank * half + i / 32) * 32 + i % 32]);
  } else {
    for (int i = threadIdx.x; i < 32 * 1024 / 4; i += blockDim.x)
      reinterpret_cast<float*>(smem)[i] = 1.0f / (float)(1 + (i & 63));
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) alloc2(slot, 512u);                         // warp 0 of BOTH CTAs (collective)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                          // both CTAs' operands and barriers are in place
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t a_lo = (smem_u32(sA) >> 4) | (1u << 16), b_lo = (smem_u32(sB) >> 4) | (1u << 16);
  const uint32_t hi = 64u | (1u << 14) | (2u << 29);           // SBO 1024 B, SWIZZLE_128B
  uint32_t phase = 0;
  if (!rate) {
    if (warp == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      for (int k = 0; k < 4; ++k) mma2_tf32(tmem, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, k ? 1u : 0u);
      commit2(bar, (uint16_t)3);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int ch = 0; ch < N / 32; ++ch) {
      float v[32];
      tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ch * 32), v);
      for (int j = 0; j < 32; ++j) D[((size_t)rank * 128 + row) * N + ch * 32 + j] = v[j];
    }
  } else {
    for (int c = 0; c < nrate; ++c) {
      const RateCase rc = rate[c];
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(rc.n >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      cluster_sync_all();
      const long long t0 = clock64();
      if (warp == 0 && rank == 0) {
        int acc = 0;
        for (int i = 0; i < rc.iters; i += 4) {
          const uint32_t d = tmem + (uint32_t)(acc * rc.n);
          for (int k = 0; k < 4; ++k)
            mma2_tf32(d, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, (i >= 4 * rc.accs || k) ? 1u : 0u);
          if (++acc == rc.accs) acc = 0;
        }
        commit2(bar, (uint16_t)3);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      const long long t1 = clock64();
      if (threadIdx.x == 0) cycles[(size_t)blockIdx.x * nrate + c] = t1 - t0;
      tc_fence_after();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                          // the peer may still read / signal until here
  if (warp == 0) { tc_fence_after(); dealloc2(tmem, 512u); }
}

float tf32_exact(int i) { return (float)((i * 37 + 11) % 61 - 30) / 16.0f; }   // exact in TF32

}  // namespace

int main() {
  int sms = 0;
  CK(cudaSetDevice(0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int smem = 32 * 1024 + 1024 + 256;
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));

  // ---- part 1: placement of A rows / B columns / accumulator lanes
  for (int N : {64, 128, 256}) {
    std::vector<float> A(256 * 32), B((size_t)N * 32), ref((size_t)256 * N), D((size_t)256 * N, -1.f);
    for (int i = 0; i < 256 * 32; ++i) A[i] = tf32_exact(i * 7 + 3);
    for (int i = 0; i < N * 32; ++i) B[i] = tf32_exact(i * 13 + 5);
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < 32; ++k) s += (double)A[m * 32 + k] * B[n * 32 + k];
        ref[(size_t)m * N + n] = (float)s;
      }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
    pair_kernel<<<2, 128, smem>>>(dA, dB, N, dD, nullptr, 0, nullptr);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (size_t i = 0; i < D.size(); ++i) worst = std::max(worst, (double)std::fabs(D[i] - ref[i]));
    printf("N=%3d  assumed placement (CTA r: rows 128r.., all columns; B columns in rank order): max |err| = %.3g  %s\n",
           N, worst, worst < 1e-3 ? "HOLDS" : "DOES NOT HOLD");
    if (worst >= 1e-3) {
      // where do the first values of each CTA's tile sit in the reference product?
      for (int r = 0; r < 2; ++r)
        for (int probe : {0, 1, 33, N + 5}) {
          size_t at = (size_t)r * 128 * N + probe;
          int hits = 0, hm = -1, hn = -1;
          for (int m = 0; m < 256; ++m)
            for (int n = 0; n < N; ++n)
              if (std::fabs(ref[(size_t)m * N + n] - D[at]) < 1e-4) { if (!hits) { hm = m; hn = n; } ++hits; }
          printf("   CTA %d element (%d, %d) = %g: %d reference matches, first at (row %d, col %d)\n", r,
                 probe / N, probe % N, D[at], hits, hm, hn);
        }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }

  // ---- part 2: issue rate of pair MMAs on all SMs
  std::vector<RateCase> rc = {{64, 1, 4096}, {64, 2, 4096}, {64, 4, 4096}, {128, 1, 4096}, {128, 2, 4096},
                              {256, 1, 4096}, {256, 2, 4096}};
  const int grid = sms & ~1;
  RateCase* d_rc; long long* d_cyc;
  CK(cudaMalloc(&d_rc, rc.size() * sizeof(RateCase)));
  CK(cudaMalloc(&d_cyc, (size_t)grid * rc.size() * sizeof(long long)));
  CK(cudaMemcpy(d_rc, rc.data(), rc.size() * sizeof(RateCase), cudaMemcpyHostToDevice));
  for (int rep = 0; rep < 3; ++rep) {
    pair_kernel<<<grid, 128, smem>>>(nullptr, nullptr, 64, nullptr, d_rc, (int)rc.size(), d_cyc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> cyc((size_t)grid * rc.size());
  CK(cudaMemcpy(cyc.data(), d_cyc, cyc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  printf("%d CTA pairs; cycles per 256 x N x 8 pair-MMA (median over CTAs [min, max]); one pair-MMA = two\n"
         "single-CTA 128 x N x 8 MMAs of tools/umma_rate_probe.cu; math floor N/2 cycles\n", grid / 2);
  for (size_t c = 0; c < rc.size(); ++c) {
    std::vector<double> v;
    for (int b = 0; b < grid; ++b) v.push_back((double)cyc[(size_t)b * rc.size() + c] / rc[c].iters);
    std::sort(v.begin(), v.end());
    printf("N=%3d accumulators %d : %6.1f  [%6.1f, %6.1f]\n", rc[c].n, rc[c].accs, v[v.size() / 2], v.front(),
           v.back());
  }
  return 0;
}
