// Mean binary cross-entropy with logits against a CONSTANT target (the GAN losses of the step:
// sg2im/losses.py:39-57, bce_loss(scores, ones / zeros)), forward and backward as one pass each.
// The reference composes it from ~10 elementwise ATen ops forward and as many backward; the step
// evaluates it six times, which made it the largest group of tiny launches left in the iteration.
//   loss = mean_i( max(x_i, 0) - x_i * t + log(1 + exp(-|x_i|)) )        (the reference's formula)
//   dx_i = (sigmoid(x_i) - t) * g / n        (x_i == 0: (1 - t) * g / n, the reference's subgradient)
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
bce_partial_kernel(const float* __restrict__ x, int64_t n, float t, double* __restrict__ acc) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    s += (double)(fmaxf(v, 0.f) - v * t + log1pf(expf(-fabsf(v))));
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(acc, sh[0]);
}

__global__ void bce_finalize_kernel(const double* __restrict__ acc, int64_t n, float* __restrict__ out) {
  out[0] = (float)(acc[0] / (double)n);
}

__global__ void __launch_bounds__(256)
bce_bwd_kernel(const float* __restrict__ x, int64_t n, float t, const float* __restrict__ gout,
               float* __restrict__ dx) {
  const float g = gout[0] / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float e = expf(-fabsf(v));
    // x == 0 exactly: the reference's composition differentiates clamp(x, min=0) as 1 and |x| as 0
    // there, i.e. d = 1 - t (not the analytic 0.5 - t); kept so that autograd results are identical
    const float sig = v > 0.f ? 1.f / (1.f + e) : (v < 0.f ? e / (1.f + e) : 1.f);
    dx[i] = (sig - t) * g;
  }
}

unsigned grid_for(int64_t n) {
  int64_t b = ceil_div64(n, 256 * 4);
  if (b < 1) b = 1;
  if (b > 148 * 4) b = 148 * 4;
  return (unsigned)b;
}

}  // namespace

extern "C" int sg2im_bce_logits_mean_fwd(const float* x, int64_t n, float target, double* scratch,
                                         float* out, sg2im_stream_t stream) {
  SG_ARG(x && scratch && out && n >= 1);
  cudaStream_t st = as_stream(stream);
  SG_LAUNCH(bce_partial_kernel, grid_for(n), 256, 0, st, x, n, target, scratch);
  SG_LAUNCH(bce_finalize_kernel, 1, 1, 0, st, scratch, n, out);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_bce_logits_mean_bwd(const float* x, int64_t n, float target, const float* gout,
                                         float* dx, sg2im_stream_t stream) {
  SG_ARG(x && gout && dx && n >= 1);
  SG_LAUNCH(bce_bwd_kernel, grid_for(n), 256, 0, as_stream(stream), x, n, target, gout, dx);
  SG_LAUNCH_OK();
  return 0;
}
