// Gradient of the fused scene layout w.r.t. the object BOXES (sg2im/layout.py:94-128 through
// F.grid_sample's grid gradient): needed only when the generator is trained on its own
// predicted boxes (Sg2ImModel.forward without boxes_gt, sg2im/model.py:151-160);
// scripts/train.py always passes boxes_gt, so this kernel is off the benchmarked step.
//
//   out[n,h,w,d] = sum_o vec[o,d] * S_o(h,w),   S_o = bilinear(mask_o; px(gx), py(gy)),
//   gx = 2 (X_w - x0) / (x1 - x0) - 1,  X_w = w / (W-1)         (likewise gy with y0, y1, Y_h)
//   dL/dbox[o] = sum_hw G_o(h,w) * ( dS/dpx * dpx/dgx * dgx/dbox  +  the y terms ),
//   G_o(h,w) = sum_d dout[n,h,w,d] * vec[o,d]
// dS/dpx is the bilinear derivative on the sample's cell with zero padding (taps outside the
// mask count as 0), exactly what ATen's grid_sampler_2d backward computes; dpx/dgx = M/2
// (align_corners=False) or (M-1)/2.
//
// One CTA per object, one warp per pixel (strided over the image): the geometry is warp-uniform,
// pixels whose sample cell lies outside the mask are skipped before any memory is touched, the
// channel dot product is a coalesced row read + shuffle reduction.  Latency / L2-bound and tiny
// next to the convolutions (O * H * W cell tests, O * box-area row reads).
#include "common.cuh"

namespace {

constexpr int LBX_THREADS = 128;

__global__ void __launch_bounds__(LBX_THREADS)
layout_bwd_boxes_kernel(const float* __restrict__ dout, int64_t dcs, const float* __restrict__ vecs,
                        const float* __restrict__ boxes, const float* __restrict__ masks, int M,
                        const int64_t* __restrict__ obj_to_img, int D, int H, int W, int align,
                        float* __restrict__ dboxes) {
  __shared__ double red[LBX_THREADS / 32][4];
  const int o = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n = obj_to_img[o];
  const float x0 = boxes[o * 4 + 0], y0 = boxes[o * 4 + 1];
  const float inv_w = 1.f / (boxes[o * 4 + 2] - x0), inv_h = 1.f / (boxes[o * 4 + 3] - y0);
  const float* mk = masks ? masks + (int64_t)o * M * M : nullptr;
  const float* vec = vecs + (int64_t)o * D;
  const float kpix = align ? 0.5f * (float)(M - 1) : 0.5f * (float)M;      // dpx/dgx
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;                            // d/dx0, d/dy0, d/dx1, d/dy1
  for (int p = warp; p < H * W; p += LBX_THREADS / 32) {
    const int h = p / W, w = p - h * W;
    const float lx = W > 1 ? (float)w / (float)(W - 1) : 0.f;
    const float ly = H > 1 ? (float)h / (float)(H - 1) : 0.f;
    const float tx = (lx - x0) * inv_w, ty = (ly - y0) * inv_h;             // box-frame coordinate
    int xl, yl;
    float wx, wy;
    bilinear_axis(tx * 2.f - 1.f, M, align, xl, wx);
    bilinear_axis(ty * 2.f - 1.f, M, align, yl, wy);
    const bool x0ok = xl >= 0 && xl < M, x1ok = xl + 1 >= 0 && xl + 1 < M;
    const bool y0ok = yl >= 0 && yl < M, y1ok = yl + 1 >= 0 && yl + 1 < M;
    float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
    if (y0ok && x0ok) v00 = mk ? mk[yl * M + xl] : 1.f;
    if (y0ok && x1ok) v01 = mk ? mk[yl * M + xl + 1] : 1.f;
    if (y1ok && x0ok) v10 = mk ? mk[(yl + 1) * M + xl] : 1.f;
    if (y1ok && x1ok) v11 = mk ? mk[(yl + 1) * M + xl + 1] : 1.f;
    const float dsx = ((v01 - v00) * (1.f - wy) + (v11 - v10) * wy) * kpix;   // dS/dgx
    const float dsy = ((v10 - v00) * (1.f - wx) + (v11 - v01) * wx) * kpix;   // dS/dgy
    if (dsx == 0.f && dsy == 0.f) continue;                                   // warp-uniform
    const float* row = dout + ((n * H + h) * (int64_t)W + w) * dcs;
    float g = 0.f;
    for (int d = lane; d < D; d += 32) g += row[d] * vec[d];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) g += __shfl_xor_sync(0xffffffffu, g, s);
    // gx = 2 tx - 1, tx = (X - x0) / (x1 - x0):  dtx/dx0 = (tx - 1) / (x1 - x0), dtx/dx1 = -tx / (x1 - x0)
    const double gx = (double)g * dsx * 2.0 * inv_w, gy = (double)g * dsy * 2.0 * inv_h;
    a0 += gx * ((double)tx - 1.0);
    a2 -= gx * (double)tx;
    a1 += gy * ((double)ty - 1.0);
    a3 -= gy * (double)ty;
  }
  if (lane == 0) { red[warp][0] = a0; red[warp][1] = a1; red[warp][2] = a2; red[warp][3] = a3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int k = 0; k < LBX_THREADS / 32; ++k) s += red[k][threadIdx.x];
    dboxes[o * 4 + threadIdx.x] = (float)s;
  }
}

}  // namespace

extern "C" int sg2im_layout_bwd_boxes(const float* dout, int64_t dout_cstride, const float* vecs,
                                      const float* boxes, const float* masks, int64_t M,
                                      const int64_t* obj_to_img, int64_t N, int64_t O, int64_t D,
                                      int64_t H, int64_t W, int align_corners, float* dboxes,
                                      sg2im_stream_t stream) {
  SG_ARG(dout && vecs && boxes && obj_to_img && dboxes);
  SG_ARG(N >= 1 && O >= 0 && D >= 1 && H >= 1 && W >= 1 && dout_cstride >= D);
  SG_ARG(masks ? (M >= 1 && M <= 4096) : true);
  SG_ARG(H * W < (1ll << 31) && O < (1ll << 31));
  if (O == 0) return 0;
  SG_LAUNCH(layout_bwd_boxes_kernel, (unsigned)O, LBX_THREADS, 0, as_stream(stream), dout,
            dout_cstride, vecs, boxes, masks, masks ? (int)M : 8, obj_to_img, (int)D, (int)H,
            (int)W, align_corners, dboxes);
  SG_LAUNCH_OK();
  return 0;
}
