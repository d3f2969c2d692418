// Hardware probe #2 (not part of the library), MN-major operands for the weight
// gradient: A = [K pixel rows][M channels], B = [K pixel rows][N channels] as TMA
// writes NHWC boxes; one MMA consumes 8 K-rows starting at an arbitrary row.
// (derived from umma_probe.cu)  does tcgen05.mma accept a K-major
// SWIZZLE_128B A-operand whose start address is shifted by whole 128-byte rows
// (not 1024-B aligned), and 8-row-group strides (SBO) that are not multiples of
// 1024 B?  That is what an in-smem halo tile needs so that the 9 taps of a 3x3
// convolution read ONE TMA-loaded tile instead of 9.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_probe tools/umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) { if (clock64() - t0 > 4000000000LL) __trap(); }
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

constexpr int KROWS = 48;        // pixel rows staged per operand
constexpr int BN = 64;

struct Case { int a_row; int b_row; int swap_lbo_sbo; };

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
      const Case* cases, int ncases, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // 176 x 128 B = 22528 -> pad to 23552
  uint8_t* sB = smem + 24 * 1024;           // 64 x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 40 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], KROWS * 128 * (4 + 2));
    for (int a = 0; a < 4; ++a) tma_load_2d(sA + a * KROWS * 128, &tmA, &bars[0], a * 32, 0);
    for (int a = 0; a < 2; ++a) tma_load_2d(sB + a * KROWS * 128, &tmB, &bars[0], a * 32, 0);
  }
  mbar_wait(&bars[0], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t parity = 0;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  for (int c = 0; c < ncases; ++c) {
    Case cs = cases[c];
    if (threadIdx.x == 0) {
      uint32_t a_addr = smem_u32(sA) + cs.a_row * 128;
      uint32_t b_addr = smem_u32(sB) + cs.b_row * 128;
      uint64_t atom = (uint64_t)((KROWS * 128) >> 4), grp = 32ull;   // K groups of 4 rows = 512 B
      uint64_t lbo = cs.swap_lbo_sbo ? grp : atom, sbo = cs.swap_lbo_sbo ? atom : grp;
      uint64_t adesc = (uint64_t)((a_addr >> 4) & 0x3FFF) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (1ull << 61);   // layout SWIZZLE_128B_BASE32B
      uint64_t bdesc = (uint64_t)((b_addr >> 4) & 0x3FFF) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (1ull << 61);   // layout SWIZZLE_128B_BASE32B
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
    }
    mbar_wait(&bars[1], parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // each warp reads its lane quadrant: 32 lanes x 64 columns
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t r[32];
      uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + ch * 32;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                     "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j)
        out[((size_t)c * 128 + warp * 32 + lane) * BN + ch * 32 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= ~0x1FFFu; memcpy(&x, &u, 4); return x; }

int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)f;
  std::vector<float> hA(KROWS * 128), hB(KROWS * BN);
  srand(1);
  for (auto& v : hA) v = tf32_trunc((rand() % 2001 - 1000) / 500.f);
  for (auto& v : hB) v = tf32_trunc((rand() % 2001 - 1000) / 500.f);
  float *dA, *dB, *dOut; Case* dC;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  std::vector<Case> cases;
  int rows[7] = {0, 8, 1, 3, 10, 13, 24};
  for (int sw = 0; sw < 2; ++sw) {
    for (int i = 0; i < 7; ++i) cases.push_back({rows[i], rows[i], sw});
    for (int i = 0; i < 7; ++i) cases.push_back({0, rows[i], sw});
    for (int i = 0; i < 7; ++i) cases.push_back({rows[i], 16, sw});
  }
  CK(cudaMalloc(&dC, cases.size() * sizeof(Case)));
  CK(cudaMemcpy(dC, cases.data(), cases.size() * sizeof(Case), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dOut, cases.size() * 128 * BN * 4));
  CUtensorMap tmA, tmB;
  {
    cuuint64_t gd[2] = {128, KROWS}; cuuint64_t gs[1] = {128 * 4}; cuuint32_t box[2] = {32, KROWS}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A %d\n", (int)r); return 1; }
    cuuint64_t gd2[2] = {BN, KROWS}; cuuint64_t gs2[1] = {BN * 4};
    r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, gd2, gs2, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B %d\n", (int)r); return 1; }
  }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
  probe<<<1, 128, 48 * 1024>>>(tmA, tmB, dC, (int)cases.size(), dOut);
  CK(cudaDeviceSynchronize());
  std::vector<float> out(cases.size() * 128 * BN);
  CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
  for (size_t c = 0; c < cases.size(); ++c) {
    Case cs = cases[c];
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < BN; ++n) {
        double ref = 0;
        for (int k = 0; k < 8; ++k) ref += (double)hA[(cs.a_row + k) * 128 + m] * hB[(cs.b_row + k) * BN + n];
        double e = fabs(ref - out[(c * 128 + m) * BN + n]);
        if (e > maxerr) maxerr = e;
        if (fabs(ref) > maxref) maxref = fabs(ref);
      }
    printf("MN-major  A rows %2d..  B rows %2d..  %s : max abs err %.3e (max |ref| %.2f)  %s\n", cs.a_row, cs.b_row,
           cs.swap_lbo_sbo ? "LBO=512,SBO=atom" : "LBO=atom,SBO=512", maxerr, maxref, maxerr < 1e-3 ? "MATCH" : "differs");
  }
  return 0;
}
