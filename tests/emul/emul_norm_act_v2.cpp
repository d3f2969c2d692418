// TEST INFRASTRUCTURE: csrc/norm_act_v2.cu compiled for the host + C entry points.
#define SG2IM_EMUL 1
#include "../../sg2im_b200/csrc/norm_act_v2.cu"

extern "C" int emul_bn_bwd_reduce_v2(const float* dy, int64_t dcs, int64_t dco, const float* x,
                                     int64_t N, int64_t H, int64_t W, int64_t C, const float* scale,
                                     const float* shift, const float* save, float slope, int up,
                                     double* sums) {
  return sg2im_bn_bwd_reduce_v2(dy, dcs, dco, x, N, H, W, C, scale, shift, save, slope, up, sums, nullptr);
}
extern "C" int emul_bn_bwd_apply_v2(const float* dy, int64_t dcs, int64_t dco, const float* x,
                                    int64_t N, int64_t H, int64_t W, int64_t C, const float* scale,
                                    const float* shift, const float* save, float slope, int up,
                                    const double* sums, float* dx) {
  return sg2im_bn_bwd_apply_v2(dy, dcs, dco, x, N, H, W, C, scale, shift, save, slope, up, sums, dx,
                               nullptr);
}
extern "C" int emul_scale_act_fwd_v2(const float* x, int64_t N, int64_t H, int64_t W, int64_t C,
                                     const float* scale, const float* shift, float slope, int up,
                                     float* y, int64_t ycs, int64_t yco, int rnd) {
  return sg2im_scale_act_fwd_v2(x, N, H, W, C, scale, shift, slope, up, y, ycs, yco, rnd, nullptr);
}
extern "C" int emul_colsum_small(const float* x, int64_t M, int64_t C, float* out) {
  return sg2im_colsum_small(x, M, C, out, nullptr);
}
