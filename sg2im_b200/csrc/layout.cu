// Fused scene layout: out[n,h,w,:] = sum_{o in image n} vecs[o,:] * S_o(h,w),
// S_o = bilinear sample of the object's MxM mask in its box frame.
// Replaces _boxes_to_grid + F.grid_sample + _pool_samples of
// sg2im/layout.py:30-162; the reference's (O,D,H,W) temporary (2.7 GB at
// VG-128) is never formed: because img_in = vec (x) mask, grid_sample is linear
// per channel and the warp factorises.
// HBM-bound: forward writes N*H*W*(D+noise) floats once, backward reads the
// output gradient inside each object's box once.
#include "common.cuh"

namespace {

struct BoxFrame {
  float x0, y0, inv_w, inv_h;      // 1/(x1-x0), 1/(y1-y0)
};

// weight of pixel (h,w) for object with mask `mk` (NULL => ones, M=8)
__device__ __forceinline__ float sample_mask(const float* __restrict__ mk, int M, int align,
                                             float gx, float gy, int& xl, int& yl, float& wx,
                                             float& wy) {
  bilinear_axis(gx, M, align, xl, wx);
  bilinear_axis(gy, M, align, yl, wy);
  float s = 0.f;
  bool x0ok = xl >= 0 && xl < M, x1ok = xl + 1 >= 0 && xl + 1 < M;
  bool y0ok = yl >= 0 && yl < M, y1ok = yl + 1 >= 0 && yl + 1 < M;
  if (mk) {
    if (y0ok && x0ok) s += mk[yl * M + xl] * (1.f - wx) * (1.f - wy);
    if (y0ok && x1ok) s += mk[yl * M + xl + 1] * wx * (1.f - wy);
    if (y1ok && x0ok) s += mk[(yl + 1) * M + xl] * (1.f - wx) * wy;
    if (y1ok && x1ok) s += mk[(yl + 1) * M + xl + 1] * wx * wy;
  } else {
    if (y0ok && x0ok) s += (1.f - wx) * (1.f - wy);
    if (y0ok && x1ok) s += wx * (1.f - wy);
    if (y1ok && x0ok) s += (1.f - wx) * wy;
    if (y1ok && x1ok) s += wx * wy;
  }
  return s;
}

// normalised sampling coordinate of output pixel index i on an axis of `size`
// pixels for a box [b0, b0 + 1/inv): linspace(0,1,size)[i] mapped to the box
// frame, then to [-1,1]   (sg2im/layout.py:115-126)
__device__ __forceinline__ float grid_coord(int i, int size, float b0, float inv) {
  float lin = size > 1 ? (float)i / (float)(size - 1) : 0.f;
  return (lin - b0) * inv * 2.f - 1.f;
}

template <int VEC>
__global__ void __launch_bounds__(256)
layout_fwd_kernel(const float* __restrict__ vecs, const float* __restrict__ boxes,
                  const float* __restrict__ masks, int M, const int32_t* __restrict__ img_ptr,
                  const int32_t* __restrict__ img_ent, int64_t N, int64_t D, int64_t H, int64_t W,
                  int align, const float* __restrict__ noise, int64_t noise_c, int64_t nsn,
                  int64_t nsc, int64_t nsh, int64_t nsw, float* __restrict__ out, int64_t ocs) {
  int64_t G = (D + (noise ? noise_c : 0)) / VEC;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H * W * G) return;
  int64_t g = i % G;
  int64_t pix = i / G;
  int w = (int)(pix % W);
  int64_t t = pix / W;
  int h = (int)(t % H);
  int64_t n = t / H;
  int64_t c = g * VEC;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (c < D) {
    int32_t b = img_ptr[n], e = img_ptr[n + 1];
    for (int32_t k = b; k < e; ++k) {
      int o = img_ent[k] >> 1;
      float4 bx = *reinterpret_cast<const float4*>(boxes + (int64_t)o * 4);
      float gx = grid_coord(w, (int)W, bx.x, 1.f / (bx.z - bx.x));
      float gy = grid_coord(h, (int)H, bx.y, 1.f / (bx.w - bx.y));
      int xl, yl; float wx, wy;
      float s = sample_mask(masks ? masks + (int64_t)o * M * M : nullptr, M, align, gx, gy, xl, yl,
                            wx, wy);
      if (s != 0.f) {
        const float* v = vecs + (int64_t)o * D + c;
        if (VEC == 4) {
          float4 vv = *reinterpret_cast<const float4*>(v);
          acc[0] += vv.x * s; acc[1] += vv.y * s; acc[2] += vv.z * s; acc[3] += vv.w * s;
        } else {
          acc[0] += v[0] * s;
        }
      }
    }
  } else {
    int64_t nc = c - D;
    const float* np = noise + n * nsn + (int64_t)h * nsh + (int64_t)w * nsw;
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = np[(nc + j) * nsc];
  }
  float* op = out + pix * ocs + c;
  if (VEC == 4) *reinterpret_cast<float4*>(op) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  else op[0] = acc[0];
}

// One CTA per (object, band of rows).  Each warp owns pixels of the band that
// fall in the object's footprint: lanes hold 4 channels each (D = 128*J),
// dot(dout, vec) is a warp-shuffle reduction, dvec accumulates in registers.
constexpr int LB_WARPS = 8, LB_ROWS = 8, LB_MAXJ = 4;

__global__ void __launch_bounds__(LB_WARPS * 32)
layout_bwd_kernel(const float* __restrict__ dout, int64_t dcs, const float* __restrict__ vecs,
                  const float* __restrict__ boxes, const float* __restrict__ masks, int M,
                  const int64_t* __restrict__ obj_to_img, int64_t N, int64_t D, int64_t H,
                  int64_t W, int align, float* __restrict__ dvecs, float* __restrict__ dmasks) {
  extern __shared__ __align__(16) float sm[];   // [LB_WARPS][D] dvec partials + [M*M] dmask
  float* sdv = sm;
  float* sdm = sm + LB_WARPS * D;
  const int o = blockIdx.x;
  const int r0 = blockIdx.y * LB_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = obj_to_img[o];
  if (n < 0 || n >= N) return;
  float4 bx = *reinterpret_cast<const float4*>(boxes + (int64_t)o * 4);
  float inv_w = 1.f / (bx.z - bx.x), inv_h = 1.f / (bx.w - bx.y);
  for (int i = threadIdx.x; i < M * M; i += blockDim.x) sdm[i] = 0.f;
  __syncthreads();
  const float* mk = masks ? masks + (int64_t)o * M * M : nullptr;
  // lane owns channels j*128 + lane*4 .. +3 (D % 4 == 0, D <= 128*LB_MAXJ: host-checked)
  bool cv[LB_MAXJ];
  float4 vv[LB_MAXJ], acc[LB_MAXJ];
#pragma unroll
  for (int j = 0; j < LB_MAXJ; ++j) {
    cv[j] = j * 128 + lane * 4 < (int)D;
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    vv[j] = cv[j] ? *reinterpret_cast<const float4*>(vecs + (int64_t)o * D + j * 128 + lane * 4)
                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int r1 = r0 + LB_ROWS < (int)H ? r0 + LB_ROWS : (int)H;
  int npix = (r1 - r0) * (int)W;
  for (int p = warp; p < npix; p += LB_WARPS) {
    int h = r0 + p / (int)W, w = p % (int)W;
    float gx = grid_coord(w, (int)W, bx.x, inv_w);
    float gy = grid_coord(h, (int)H, bx.y, inv_h);
    int xl, yl; float wx, wy;
    float s = sample_mask(mk, M, align, gx, gy, xl, yl, wx, wy);
    bool xin = (xl >= -1 && xl < M), yin = (yl >= -1 && yl < M);
    if (!(xin && yin)) continue;                 // footprint entirely in the padding
    const float* dp = dout + ((n * H + h) * W + w) * dcs + lane * 4;
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < LB_MAXJ; ++j) {
      if (cv[j]) {
        float4 d = *reinterpret_cast<const float4*>(dp + j * 128);
        dot += d.x * vv[j].x + d.y * vv[j].y + d.z * vv[j].z + d.w * vv[j].w;
        acc[j].x += d.x * s; acc[j].y += d.y * s; acc[j].z += d.z * s; acc[j].w += d.w * s;
      }
    }
    if (dmasks) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
      if (lane < 4) {
        int cx = xl + (lane & 1), cy = yl + (lane >> 1);
        float wgt = ((lane & 1) ? wx : 1.f - wx) * ((lane >> 1) ? wy : 1.f - wy);
        if (cx >= 0 && cx < M && cy >= 0 && cy < M) atomicAdd(&sdm[cy * M + cx], wgt * dot);
      }
    }
  }
  // cross-warp reduction of dvec, then one atomic per channel per CTA
#pragma unroll
  for (int j = 0; j < LB_MAXJ; ++j)
    if (cv[j]) *reinterpret_cast<float4*>(&sdv[warp * D + j * 128 + lane * 4]) = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < (int)D; c += blockDim.x) {
    float s = 0.f;
    for (int wv = 0; wv < LB_WARPS; ++wv) s += sdv[wv * D + c];
    if (s != 0.f) atomicAdd(dvecs + (int64_t)o * D + c, s);
  }
  if (dmasks) {
    for (int i = threadIdx.x; i < M * M; i += blockDim.x) {
      float s = sdm[i];
      if (s != 0.f) atomicAdd(dmasks + (int64_t)o * M * M + i, s);
    }
  }
}

}  // namespace

extern "C" int sg2im_layout_fwd(const float* vecs, const float* boxes, const float* masks, int64_t M,
                                const int32_t* img_row_ptr, const int32_t* img_entries,
                                int64_t N, int64_t O, int64_t D, int64_t H, int64_t W,
                                int align_corners, const float* noise, int64_t noise_c,
                                int64_t nsn, int64_t nsc, int64_t nsh, int64_t nsw,
                                float* out, int64_t out_cstride, sg2im_stream_t stream) {
  SG_ARG(vecs && boxes && img_row_ptr && img_entries && out);
  SG_ARG(N >= 1 && O >= 1 && D >= 1 && H >= 1 && W >= 1);
  if (!masks) M = 8;
  SG_ARG(M >= 1 && M <= 1024);
  SG_ARG(noise == nullptr || noise_c >= 1);
  int64_t ctot = D + (noise ? noise_c : 0);
  SG_ARG(out_cstride >= ctot);
  SG_ARG(aligned16(boxes));
  bool vec = (D % 4 == 0) && (ctot % 4 == 0) && (out_cstride % 4 == 0) && aligned16(vecs) &&
             aligned16(out);
  int64_t total = N * H * W * (ctot / (vec ? 4 : 1));
  unsigned grid = (unsigned)ceil_div64(total, 256);
  cudaStream_t st = as_stream(stream);
  if (vec)
    layout_fwd_kernel<4><<<grid, 256, 0, st>>>(vecs, boxes, masks, (int)M, img_row_ptr, img_entries,
                                               N, D, H, W, align_corners, noise, noise_c, nsn, nsc,
                                               nsh, nsw, out, out_cstride);
  else
    layout_fwd_kernel<1><<<grid, 256, 0, st>>>(vecs, boxes, masks, (int)M, img_row_ptr, img_entries,
                                               N, D, H, W, align_corners, noise, noise_c, nsn, nsc,
                                               nsh, nsw, out, out_cstride);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_layout_bwd(const float* dout, int64_t dout_cstride, const float* vecs,
                                const float* boxes, const float* masks, int64_t M,
                                const int64_t* obj_to_img, int64_t N, int64_t O, int64_t D,
                                int64_t H, int64_t W, int align_corners, float* dvecs,
                                float* dmasks, sg2im_stream_t stream) {
  SG_ARG(dout && vecs && boxes && obj_to_img && dvecs);
  SG_ARG(N >= 1 && O >= 1 && H >= 1 && W >= 1);
  if (!masks) { M = 8; SG_ARG(dmasks == nullptr); }
  SG_ARG(M >= 1 && M <= 64);
  if (D % 4 != 0 || D > 128 * LB_MAXJ) {
    sg2im_set_error("sg2im_layout_bwd: unsupported D=%lld (need a multiple of 4, <= %d)",
                    (long long)D, 128 * LB_MAXJ);
    return -2;
  }
  SG_ARG(dout_cstride % 4 == 0 && dout_cstride >= D && aligned16(dout) && aligned16(vecs) &&
         aligned16(boxes));
  dim3 grid((unsigned)O, (unsigned)ceil_div64(H, LB_ROWS));
  size_t smem = (size_t)(M * M + LB_WARPS * D) * sizeof(float);
  layout_bwd_kernel<<<grid, LB_WARPS * 32, smem, as_stream(stream)>>>(
      dout, dout_cstride, vecs, boxes, masks, (int)M, obj_to_img, N, D, H, W, align_corners, dvecs,
      dmasks);
  SG_LAUNCH_OK();
  return 0;
}
