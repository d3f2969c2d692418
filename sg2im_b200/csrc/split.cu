// TF32 hi / lo split of an fp32 activation for the error-compensated tensor-core mode
// (ops.set_conv_math('tf32x3')):  hi = round-to-nearest-TF32(x),  lo = x - hi  (exact in fp32,
// |lo| <= 2^-11 |x|).  A product x*w evaluated as hi_x*hi_w + lo_x*hi_w + hi_x*lo_w on the TF32
// tensor core (fp32 accumulate) drops only lo_x*lo_w and the truncation of the lo operands,
// ~2^-21 relative: fp32-grade results from the tcgen05 kernels.  The three products are laid
// side by side along the GEMM's reduction dimension (input channels), so ONE launch of the
// unchanged convolution kernel accumulates them in TMEM:
//     x3 = [hi_x | lo_x | hi_x]  (3C channels)      w3 = [hi_w | hi_w | lo_w]
// Rows are pixels (row stride in floats, so a channel-prefix view of a wider NHWC buffer works);
// every destination is optional and has its own row stride, so the same kernel writes the
// concatenated form (three pointers into one buffer) or separate hi / lo tensors.
// HBM-bound: reads x once, writes 2-3x its size.
#include "common.cuh"

namespace {

template <int VEC>
__global__ void split_tf32_kernel(const float* __restrict__ x, uint32_t rows, uint32_t C,
                                  int64_t xs, float* __restrict__ hi, int64_t his,
                                  float* __restrict__ lo, int64_t los,
                                  float* __restrict__ hi2, int64_t hi2s) {
  uint32_t cg = C / VEC;
  uint32_t ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= rows * cg) return;
  uint32_t r = ii / cg, c = (ii % cg) * VEC;
  float v[VEC], h[VEC], l[VEC];
  if (VEC == 4) {
    float4 t = *reinterpret_cast<const float4*>(x + (int64_t)r * xs + c);
    v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
  } else {
    v[0] = x[(int64_t)r * xs + c];
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    h[k] = tf32_rn(v[k]);
    l[k] = v[k] - h[k];
    if (!(fabsf(v[k]) <= 3.0e38f)) l[k] = 0.f;        // inf / nan stay in the hi part only
  }
  if (VEC == 4) {
    if (hi)  *reinterpret_cast<float4*>(hi + (int64_t)r * his + c) = make_float4(h[0], h[1 % VEC], h[2 % VEC], h[3 % VEC]);
    if (lo)  *reinterpret_cast<float4*>(lo + (int64_t)r * los + c) = make_float4(l[0], l[1 % VEC], l[2 % VEC], l[3 % VEC]);
    if (hi2) *reinterpret_cast<float4*>(hi2 + (int64_t)r * hi2s + c) = make_float4(h[0], h[1 % VEC], h[2 % VEC], h[3 % VEC]);
  } else {
    if (hi)  hi[(int64_t)r * his + c] = h[0];
    if (lo)  lo[(int64_t)r * los + c] = l[0];
    if (hi2) hi2[(int64_t)r * hi2s + c] = h[0];
  }
}

bool vec_ok(const float* p, int64_t stride) { return p == nullptr || (aligned16(p) && stride % 4 == 0); }

}  // namespace

extern "C" int sg2im_split_tf32(const float* x, int64_t rows, int64_t C, int64_t x_stride,
                                float* hi, int64_t hi_stride, float* lo, int64_t lo_stride,
                                float* hi2, int64_t hi2_stride, sg2im_stream_t stream) {
  SG_ARG(x && rows >= 0 && C >= 1 && x_stride >= C);
  SG_ARG((hi || lo || hi2) && (!hi || hi_stride >= C) && (!lo || lo_stride >= C) &&
         (!hi2 || hi2_stride >= C));
  SG_ARG(rows * C < (1ll << 32));
  if (rows == 0) return 0;
  bool vec = C % 4 == 0 && vec_ok(x, x_stride) && vec_ok(hi, hi_stride) && vec_ok(lo, lo_stride) &&
             vec_ok(hi2, hi2_stride);
  int64_t total = rows * (C / (vec ? 4 : 1));
  unsigned grid = (unsigned)ceil_div64(total, 256);
  cudaStream_t st = as_stream(stream);
  if (vec) SG_LAUNCH(split_tf32_kernel<4>, grid, 256, 0, st, x, (uint32_t)rows, (uint32_t)C, x_stride, hi, hi_stride, lo, lo_stride, hi2, hi2_stride);
  else     SG_LAUNCH(split_tf32_kernel<1>, grid, 256, 0, st, x, (uint32_t)rows, (uint32_t)C, x_stride, hi, hi_stride, lo, lo_stride, hi2, hi2_stride);
  SG_LAUNCH_OK();
  return 0;
}
