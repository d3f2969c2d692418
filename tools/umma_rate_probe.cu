// Hardware probe (not part of the library): sustained issue rate of tcgen05.mma kind::tf32 for the
// operand layouts and tile shapes the convolution kernels use, with the operands resident in shared
// memory (no TMA, no epilogue) — the ceiling each kernel's main loop can reach, measured instead of
// assumed.  One CTA per SM on all SMs (clocks / power as in a real launch), a converged warp issues
// `iters` MMAs with lane-0 election exactly like the kernels (tc_common.cuh wrappers), one
// tcgen05.commit at the end, clock64 around the whole sequence.
//
// Questions it answers:
//   * cycles per 128 x N x 8 MMA for N = 64 / 128 / 256 — is N = 64 bound by the shared-memory
//     operand feed (6 KB per MMA) rather than by the 32 math cycles?
//   * does the halo kernel's A addressing (8-row groups 1280 B apart, row-shifted starts) cost
//     anything against dense 1024 B groups?
//   * does alternating between 2 / 4 accumulators beat a chain on one accumulator?
//   * MN-major B (forward straight from the weight-gradient layout) and MN-major A + B (weight
//     gradient) against K-major operands.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I sg2im_b200/csrc \
//        -o gpurun_out/umma_rate_probe tools/umma_rate_probe.cu -lcuda && gpurun_out/umma_rate_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "tc_common.cuh"

using namespace tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Case {
  int n;            // N tile: 64 / 128 / 256
  int a_mode;       // 0: K-major dense (SBO 1024)   1: K-major halo (SBO 1280, start shifted per tap)
                    // 2: MN-major (weight gradient: 4 atoms of 8 KB, SBO 512)
  int b_mode;       // 0: K-major dense   1: MN-major (SBO 512, atoms 4 KB apart)
  int accs;         // accumulators cycled through (1, 2, 4)
  int iters;        // MMAs issued
  int m;            // M: 128 (the kernels) or 64 (weights as the A operand: D^T = W X^T)
};

__global__ void __launch_bounds__(128, 1)
rate_kernel(const Case* cases, int ncases, long long* cycles) {
  extern __shared__ __align__(16) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                         // 64 KB: room for every A layout
  uint8_t* sB = smem + 64 * 1024;             // 64 KB: 256 rows x 128 B x 2
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 128 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<float*>(smem)[i] = 1.0f / (float)(1 + (i & 63));      // finite operands
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) tc_alloc(slot, 512u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic writes -> async proxy reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    const uint32_t leader = lane == 0;
    uint32_t phase = 0;
    for (int c = 0; c < ncases; ++c) {
      const Case cs = cases[c];
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(cs.n >> 3) << 17) |
                             ((uint32_t)(cs.m >> 4) << 24) | (cs.a_mode == 2 ? (1u << 15) : 0u) |
                             (cs.b_mode == 1 ? (1u << 16) : 0u);
      // A descriptor (lo, hi)
      uint32_t a_lo = (smem_u32(sA) >> 4), a_hi;
      if (cs.a_mode == 0) { a_lo |= 1u << 16; a_hi = 64u | (1u << 14) | (2u << 29); }
      else if (cs.a_mode == 1) { a_lo |= 1u << 16; a_hi = (1280u >> 4) | (1u << 14) | (2u << 29); }
      else { a_lo |= (8192u >> 4) << 16; a_hi = 32u | (1u << 14) | (1u << 29); }
      uint32_t b_lo = (smem_u32(sB) >> 4), b_hi;
      if (cs.b_mode == 0) { b_lo |= 1u << 16; b_hi = 64u | (1u << 14) | (2u << 29); }
      else { b_lo |= (4096u >> 4) << 16; b_hi = 32u | (1u << 14) | (1u << 29); }
      const int acc_cols = cs.n;
      __syncwarp();
      const long long t0 = clock64();
      int tap = 0, acc = 0;
      for (int i = 0; i < cs.iters; i += 4) {
        // one "tap": 4 k-steps of 8 (32 channels), as in the kernels
        uint32_t at = a_lo, bt = b_lo;
        if (cs.a_mode == 1) at += (uint32_t)((tap / 3) * 10 + tap % 3) * 8u;      // row-shifted start
        const uint32_t d = tmem + (uint32_t)(acc * acc_cols);
        const uint32_t ka = cs.a_mode == 2 ? 64u : 2u, kb = cs.b_mode == 1 ? 64u : 2u;
        tc_mma_tf32_lh(d, at, a_hi, bt, b_hi, idesc, i >= 4 * cs.accs ? 1u : 0u, leader);
        tc_mma_tf32_lh(d, at + ka, a_hi, bt + kb, b_hi, idesc, 1u, leader);
        tc_mma_tf32_lh(d, at + 2 * ka, a_hi, bt + 2 * kb, b_hi, idesc, 1u, leader);
        tc_mma_tf32_lh(d, at + 3 * ka, a_hi, bt + 3 * kb, b_hi, idesc, 1u, leader);
        if (++tap == 9) tap = 0;
        if (++acc == cs.accs) acc = 0;
      }
      tc_commit(bar, leader);
      mbar_wait(bar, phase);
      phase ^= 1;
      const long long t1 = clock64();
      if (lane == 0) cycles[(size_t)blockIdx.x * ncases + c] = t1 - t0;
      tc_fence_after();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tc_dealloc(tmem, 512u); }
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  std::vector<Case> cases;
  const int IT = 4096;
  for (int n : {64, 128, 256}) cases.push_back({n, 0, 0, 1, IT, 128});          // K-major dense, one accumulator
  cases.push_back({64, 1, 0, 1, IT, 128});                                      // halo A addressing
  cases.push_back({64, 0, 0, 2, IT, 128});                                      // alternate accumulators
  cases.push_back({64, 0, 0, 4, IT, 128});
  cases.push_back({64, 1, 0, 4, IT, 128});
  cases.push_back({128, 0, 0, 2, IT, 128});
  for (int n : {64, 128, 256}) cases.push_back({n, 0, 1, 1, IT, 128});          // MN-major B (kcc forward)
  cases.push_back({64, 1, 1, 1, IT, 128});                                      // halo A + MN-major B
  for (int n : {64, 128, 256}) cases.push_back({n, 2, 1, 1, IT, 128});          // weight gradient operands
  cases.push_back({256, 2, 1, 2, IT, 128});
  for (int n : {64, 128, 256}) cases.push_back({n, 0, 0, 1, IT, 64});      // M = 64: half the rows per MMA
  cases.push_back({256, 0, 0, 2, IT, 64});
  const int nc = (int)cases.size();
  Case* d_cases; long long* d_cyc;
  CK(cudaMalloc(&d_cases, nc * sizeof(Case)));
  CK(cudaMalloc(&d_cyc, (size_t)sms * nc * sizeof(long long)));
  CK(cudaMemcpy(d_cases, cases.data(), nc * sizeof(Case), cudaMemcpyHostToDevice));
  const int smem = 128 * 1024 + 1024 + 256;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 3; ++rep) {                                       // first rep warms clocks
    rate_kernel<<<sms, 128, smem>>>(d_cases, nc, d_cyc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> cyc((size_t)sms * nc);
  CK(cudaMemcpy(cyc.data(), d_cyc, cyc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  const char* am[] = {"K-major dense", "K-major halo (SBO 1280, shifted)", "MN-major"};
  const char* bm[] = {"K-major", "MN-major"};
  printf("%d SMs, %d MMAs (M x N x 8, tf32) per case; cycles per MMA: median over SMs [min, max]; "
         "math floor at M = 128: N/2 cycles\n", sms, IT);
  for (int c = 0; c < nc; ++c) {
    std::vector<double> v;
    for (int b = 0; b < sms; ++b) v.push_back((double)cyc[(size_t)b * nc + c] / cases[c].iters);
    std::sort(v.begin(), v.end());
    printf("M=%3d N=%3d  A: %-34s B: %-9s accumulators %d : %6.1f  [%6.1f, %6.1f]\n", cases[c].m, cases[c].n,
           am[cases[c].a_mode], bm[cases[c].b_mode], cases[c].accs, v[v.size() / 2], v.front(), v.back());
  }
  return 0;
}
