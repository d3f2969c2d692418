// Shared helpers for libsg2im_b200 (sm_100a).  Internal — the public surface is
// include/sg2im_b200.h.
#pragma once
#ifdef SG2IM_EMUL
// tests/emul/cuda_emul.h: host-side SIMT emulation of the CUDA builtins — lets the CPU test
// suite execute the non-tensor-core kernel sources themselves (test infrastructure only)
#include "cuda_emul.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include "sg2im_b200.h"

// Kernel launch and dynamic shared memory spelled so that the same source also builds for the
// emulation harness; under nvcc these expand to the plain CUDA forms.
#ifdef SG2IM_EMUL
#define SG_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emul_launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); })
#define SG_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emul_dynamic_smem())
#else
#define SG_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define SG_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

void sg2im_set_error(const char* fmt, ...);

#define SG_ARG(cond)                                                          \
  do {                                                                        \
    if (!(cond)) {                                                            \
      sg2im_set_error("%s: invalid argument: %s", __func__, #cond);           \
      return -1;                                                              \
    }                                                                         \
  } while (0)

#ifdef SG2IM_EMUL
#define SG_LAUNCH_OK() do { } while (0)
#else
#define SG_LAUNCH_OK()                                                        \
  do {                                                                        \
    cudaError_t e_ = cudaGetLastError();                                      \
    if (e_ != cudaSuccess) {                                                  \
      sg2im_set_error("%s: launch failed: %s", __func__, cudaGetErrorString(e_)); \
      return (int)e_;                                                         \
    }                                                                         \
  } while (0)
#endif

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }
static inline cudaStream_t as_stream(sg2im_stream_t s) { return (cudaStream_t)s; }

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }

// Round-to-nearest TF32 (10-bit mantissa, low 13 bits cleared).  The tensor core
// TRUNCATES the fp32 words it is fed, which biases every product toward zero;
// operands written through this are consumed exactly, so the only error left is
// this unbiased rounding (what cuBLAS/cuDNN TF32 paths do with cvt.rna).
__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t u;
#ifdef SG2IM_EMUL
  u = __float_as_uint(v);                     // cvt.rna: nearest, ties away from zero (magnitude)
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x1000u;
  u &= ~0x1fffu;
#else
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
#endif
  return __uint_as_float(u);
}

// (x0, x1) -> packed bf16x2 of the round-to-nearest-even bf16 values (x0 in the low half) and of
// the bf16-rounded remainders: x = hi + mid up to 2^-17 |x| (the subtraction is exact).  The
// operand split of the 'bf16x3' tensor-core arithmetic (tc_common.cuh, pack.cu).
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& mid) {
#ifdef SG2IM_EMUL
  auto rn = [](float f) -> uint32_t {                    // what cvt.rn.bf16x2.f32 does per element
    uint32_t u = __float_as_uint(f);
    if ((u & 0x7f800000u) == 0x7f800000u) return u >> 16;
    return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
  };
  const uint32_t h0 = rn(x0), h1 = rn(x1);
  hi = h0 | (h1 << 16);
  mid = rn(x0 - __uint_as_float(h0 << 16)) | (rn(x1 - __uint_as_float(h1 << 16)) << 16);
#else
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(r1), "f"(r0));
#endif
}

// Bilinear sampling footprint of one normalised coordinate g in [-1,1] on an
// axis of `size` texels (torch grid_sample, zeros padding): lower texel index
// and the weight of the upper texel.
__device__ __forceinline__ void bilinear_axis(float g, int size, int align_corners,
                                              int& lo, float& w_hi) {
  float pix = align_corners ? (g + 1.f) * 0.5f * (float)(size - 1)
                            : ((g + 1.f) * (float)size - 1.f) * 0.5f;
  float fl = floorf(pix);
  // clamp far-out coordinates so the int conversion is defined; anything
  // beyond [-2, size+1] samples only padding anyway
  fl = fminf(fmaxf(fl, -2.f), (float)size + 1.f);
  lo = (int)fl;
  w_hi = pix - fl;
  if (!(pix >= -2.f && pix <= (float)size + 1.f)) { lo = -2; w_hi = 0.f; }   // also NaN
}
