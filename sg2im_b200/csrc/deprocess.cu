// imagenet_deprocess_batch (sg2im/data/utils.py:32-67) on the device: undo the
// ImageNet normalisation, rescale every image to [0,1] by its own min/max over
// all channels, x255, clamp, truncate to uint8.  The reference does this on the
// CPU, one image at a time, after a full-precision device->host copy; here the
// float image never leaves HBM and the host receives bytes (4x less PCIe).
// HBM-bound: reads the image twice (min/max pass, map pass), writes N*C*H*W bytes.
//
// Arithmetic order is the reference's, in fp32 with IEEE division (nvcc default
// -prec-div=true), so the bytes are identical:
//   v = (x - 0) / inv_std[c];  v = (v - neg_mean[c]) / 1;        (T.Normalize x2)
//   v = (v - lo) / (hi - lo)                       if rescale     (utils.py:27-29)
//   byte = trunc(clamp(v * 255, 0, 255))                          (utils.py:62)
#include "common.cuh"

namespace {

// order-preserving map float -> uint32 so min/max can use integer atomics
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float denorm(float x, float inv_std, float neg_mean) {
  float v = x / inv_std;
  return v - neg_mean;
}

__global__ void deprocess_init_kernel(uint32_t* mm, int64_t N) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { mm[2 * i] = 0xffffffffu; mm[2 * i + 1] = 0u; }
}

// grid (blocks per image, N)
__global__ void deprocess_minmax_kernel(const float* __restrict__ x, int64_t sn, int64_t sc,
                                        int64_t sh, int64_t sw, int C, int H, int W,
                                        const float* __restrict__ inv_std,
                                        const float* __restrict__ neg_mean,
                                        uint32_t* __restrict__ mm) {
  const int64_t n = blockIdx.y;
  const int64_t per = (int64_t)C * H * W;
  const float* xn = x + n * sn;
  float lo = INFINITY, hi = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per;
       i += (int64_t)gridDim.x * blockDim.x) {
    // iterate in (h, w, c) order: unit stride for the NHWC buffers the generator writes
    int c = (int)(i % C);
    int64_t t = i / C;
    int w = (int)(t % W);
    int h = (int)(t / W);
    float v = denorm(xn[c * sc + h * sh + w * sw], inv_std[c], neg_mean[c]);
    lo = fminf(lo, v); hi = fmaxf(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float s_lo[32], s_hi[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    lo = lane < nw ? s_lo[lane] : INFINITY;
    hi = lane < nw ? s_hi[lane] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
      atomicMin(&mm[2 * n], f2ord(lo));
      atomicMax(&mm[2 * n + 1], f2ord(hi));
    }
  }
}

__global__ void deprocess_map_kernel(const float* __restrict__ x, int64_t sn, int64_t sc,
                                     int64_t sh, int64_t sw, int64_t N, int C, int H, int W,
                                     const float* __restrict__ inv_std,
                                     const float* __restrict__ neg_mean,
                                     const uint32_t* __restrict__ mm, int rescale,
                                     uint8_t* __restrict__ out, int64_t on, int64_t oc,
                                     int64_t oh, int64_t ow) {
  const int64_t per = (int64_t)C * H * W;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * per) return;
  const int64_t n = i / per;
  int64_t r = i - n * per;
  int c = (int)(r % C);
  int64_t t = r / C;
  int w = (int)(t % W);
  int h = (int)(t / W);
  float v = denorm(x[n * sn + c * sc + h * sh + w * sw], inv_std[c], neg_mean[c]);
  if (rescale) {
    float lo = ord2f(mm[2 * n]), hi = ord2f(mm[2 * n + 1]);
    v = (v - lo) / (hi - lo);
  }
  v = v * 255.f;
  v = fminf(fmaxf(v, 0.f), 255.f);                      // NaN -> 0 (fmaxf returns the number)
  out[n * on + c * oc + h * oh + w * ow] = (uint8_t)v;   // truncation, like Tensor.byte()
}

}  // namespace

extern "C" int sg2im_deprocess(const float* imgs, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                               int64_t N, int64_t C, int64_t H, int64_t W,
                               const float* inv_std, const float* neg_mean, int rescale,
                               uint32_t* minmax, uint8_t* out, int64_t on, int64_t oc,
                               int64_t oh, int64_t ow, sg2im_stream_t stream) {
  SG_ARG(imgs && inv_std && neg_mean && out);
  SG_ARG(N >= 0 && C >= 1 && C <= 1024 && H >= 1 && W >= 1 && H <= 65536 && W <= 65536);
  SG_ARG(!rescale || minmax);
  if (N == 0) return 0;
  SG_ARG(N <= 65535);                                    // gridDim.y of the min/max pass
  cudaStream_t st = as_stream(stream);
  const int64_t per = C * H * W;
  if (rescale) {
    SG_LAUNCH(deprocess_init_kernel, (unsigned)ceil_div64(N, 128), 128, 0, st, minmax, N);
    int64_t want = ceil_div64(148 * 8, N);               // ~8 CTAs per SM over the whole batch
    int64_t fit = ceil_div64(per, 256 * 4);
    unsigned bx = (unsigned)(want < fit ? want : fit);
    if (bx < 1) bx = 1;
    dim3 grid(bx, (unsigned)N);
    SG_LAUNCH(deprocess_minmax_kernel, grid, 256, 0, st, imgs, sn, sc, sh, sw, (int)C, (int)H, (int)W,
                                                  inv_std, neg_mean, minmax);
  }
  const int64_t total = N * per;
  SG_ARG(ceil_div64(total, 256) <= 0x7fffffff);
  SG_LAUNCH(deprocess_map_kernel, (unsigned)ceil_div64(total, 256), 256, 0, st, 
      imgs, sn, sc, sh, sw, N, (int)C, (int)H, (int)W, inv_std, neg_mean, minmax, rescale, out, on,
      oc, oh, ow);
  SG_LAUNCH_OK();
  return 0;
}
