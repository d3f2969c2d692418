// Adam over one FLAT fp32 bucket per network (parameters, gradients and both
// moments are four equally long contiguous arrays; sg2im_b200/train_step.py lays
// every parameter / .grad out as a view into them).  Replaces the six
// multi-tensor launches of torch.optim.Adam(fused=True) per optimiser
// (scripts/train.py:426,436,443 + :560,579,592) with one streaming pass:
// 4 reads + 3 writes of 16 bytes per thread-iteration, HBM-bound.
//
// Capturable semantics: the step count lives on the device and `found_inf`
// (device float, nonzero = skip) suppresses the whole update including the
// step increment — what the CUDA-graph path uses for the collective
// non-finite-loss skip of train.py:552-555.
//
// Arithmetic (torch/optim/adam.py, single-tensor path, amsgrad=False,
// maximize=False):  m += (g - m)(1 - b1);  v = b2 v + (1 - b2) g g;
// p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps), weight decay
// (L2, added to g) optional.
#include "common.cuh"

namespace {

__global__ void adam_tick_kernel(float* step, const float* found_inf) {
  if (found_inf && *found_inf != 0.f) return;
  *step += 1.f;
}

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                 float wd, const float* __restrict__ step, const float* __restrict__ found_inf,
                 float* __restrict__ rounded, float gscale) {
  if (found_inf && *found_inf != 0.f) return;
  const float t = *step;
  // bias corrections in double like the host-side reference implementation
  const float bc1 = (float)(1.0 - pow((double)b1, (double)t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)t));
  const float step_size = lr / bc1;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ga[4] = {gg.x, gg.y, gg.z, gg.w};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gj = ga[j] * gscale;                          // 1 / world of a SUM all-reduce (1 = exact no-op)
      gj = wd != 0.f ? fmaf(wd, pa[j], gj) : gj;
      ma[j] = ma[j] + (gj - ma[j]) * (1.f - b1);
      va[j] = b2 * va[j] + (1.f - b2) * gj * gj;
      float denom = sqrtf(va[j]) / bc2_sqrt + eps;
      pa[j] = pa[j] - step_size * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    if (rounded)                                       // RN-TF32 shadow the tensor-core kernels read
      reinterpret_cast<float4*>(rounded)[i] =
          make_float4(tf32_rn(pa[0]), tf32_rn(pa[1]), tf32_rn(pa[2]), tf32_rn(pa[3]));
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
  }
  // tail (n % 4 elements), one thread
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = n4 << 2; i < n; ++i) {
      float gj = g[i] * gscale;
      gj = wd != 0.f ? fmaf(wd, p[i], gj) : gj;
      float mj = m[i] + (gj - m[i]) * (1.f - b1);
      float vj = b2 * v[i] + (1.f - b2) * gj * gj;
      m[i] = mj; v[i] = vj;
      p[i] = p[i] - step_size * (mj / (sqrtf(vj) / bc2_sqrt + eps));
      if (rounded) rounded[i] = tf32_rn(p[i]);
    }
  }
}

__global__ void __launch_bounds__(256)
round_tf32_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = tf32_rn(x[i]);
}

}  // namespace

extern "C" int sg2im_round_tf32(const float* x, int64_t n, float* y, sg2im_stream_t stream) {
  SG_ARG(x && y && n >= 0);
  if (n == 0) return 0;
  int64_t blocks = ceil_div64(n, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  SG_LAUNCH(round_tf32_kernel, (unsigned)blocks, 256, 0, as_stream(stream), x, n, y);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                               int64_t n, float lr, float beta1, float beta2, float eps,
                               float weight_decay, float* step, const float* found_inf,
                               float* rounded_out, float grad_scale, sg2im_stream_t stream) {
  SG_ARG(params && grads && exp_avg && exp_avg_sq && step && n >= 0);
  SG_ARG(aligned16(params) && aligned16(grads) && aligned16(exp_avg) && aligned16(exp_avg_sq));
  SG_ARG(rounded_out == nullptr || aligned16(rounded_out));
  SG_ARG(lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f);
  cudaStream_t st = as_stream(stream);
  SG_LAUNCH(adam_tick_kernel, 1, 1, 0, st, step, found_inf);
  if (n > 0) {
    int64_t blocks = ceil_div64(ceil_div64(n, 4), 256);
    const int64_t cap = 148 * 16;                       // grid-stride beyond ~2 waves of 8 CTAs/SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    SG_LAUNCH(adam_flat_kernel, (unsigned)blocks, 256, 0, st, params, grads, exp_avg, exp_avg_sq, n, lr,
                                                       beta1, beta2, eps, weight_decay, step,
                                                       found_inf, rounded_out, grad_scale);
  }
  SG_LAUNCH_OK();
  return 0;
}
