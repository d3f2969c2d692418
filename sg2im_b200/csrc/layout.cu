// Fused scene layout: out[n,h,w,:] = sum_{o in image n} vecs[o,:] * S_o(h,w),
// S_o = bilinear sample of the object's MxM mask in its box frame.
// Replaces _boxes_to_grid + F.grid_sample + _pool_samples of
// sg2im/layout.py:30-162; the reference's (O,D,H,W) temporary (2.7 GB at
// VG-128) is never formed: because img_in = vec (x) mask, grid_sample is linear
// per channel and the warp factorises.
// HBM-bound: forward writes N*H*W*(D+noise) floats once, backward reads the
// output gradient inside each object's box once.
#include <cstdlib>
#include "common.cuh"

int sg2im_layout_bwd_v2(const float* dout, int64_t dcs, const float* vecs, const float* boxes,
                        const float* masks, int64_t M, const int64_t* obj_to_img, int64_t N,
                        int64_t O, int64_t D, int64_t H, int64_t W, int align, float* dvecs,
                        float* dmasks, cudaStream_t st);
int sg2im_layout_fwd_v2(const float* vecs, const float* boxes, const float* masks, int64_t M,
                        const int32_t* img_ptr, const int32_t* img_ent, int64_t N, int64_t D,
                        int64_t H, int64_t W, int align, const float* noise, int64_t noise_c,
                        int64_t nsn, int64_t nsc, int64_t nsh, int64_t nsw, float* out, int64_t ocs,
                        int rnd, cudaStream_t st);

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

// Forward, one CTA per (image, row, 32-pixel segment):
//   phase 1  S[o][p] for the image's objects (chunks of LF_OBJ) and the 32 pixels
//            — one bilinear mask sample per (object, pixel), not per channel;
//   phase 2  out[p][c..c+3] += S[o][p] * vec[o][c..c+3] with vec staged in smem;
//            threads walk (pixel, channel-group) so each pixel row is written as
//            one contiguous run of float4 (coalesced 16 B stores).
// Noise channels are transposed NCHW -> NHWC through a padded smem tile.
constexpr int LF_PIX = 32, LF_OBJ = 16, LF_THREADS = 256, LF_MAXACC = 8;

__global__ void __launch_bounds__(LF_THREADS)
layout_fwd_kernel(const float* __restrict__ vecs, const float* __restrict__ boxes,
                  const float* __restrict__ masks, int M, const int32_t* __restrict__ img_ptr,
                  const int32_t* __restrict__ img_ent, int64_t N, int D, int H, int W,
                  int align, const float* __restrict__ noise, int noise_c, int64_t nsn,
                  int64_t nsc, int64_t nsh, int64_t nsw, float* __restrict__ out, int64_t ocs,
                  int rnd) {
  SG_DYN_SMEM(float, lsm);
  float* sS = lsm;                               // [LF_OBJ][LF_PIX]
  float* sV = lsm + LF_OBJ * LF_PIX;             // [LF_OBJ][D]
  const int segs = (W + LF_PIX - 1) / LF_PIX;
  int bid = blockIdx.x;
  const int seg = bid % segs; bid /= segs;
  const int h = bid % H;
  const int64_t n = bid / H;
  const int w0 = seg * LF_PIX;
  const int npx = W - w0 < LF_PIX ? W - w0 : LF_PIX;
  const int t = threadIdx.x;
  const int G = D >> 2;                           // float4 groups per pixel (D % 4 == 0)
  const int nout = npx * G;
  float4 acc[LF_MAXACC];
#pragma unroll
  for (int k = 0; k < LF_MAXACC; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int32_t ob = img_ptr[n], oe = img_ptr[n + 1];
  for (int32_t c0 = ob; c0 < oe; c0 += LF_OBJ) {
    const int nobj = oe - c0 < LF_OBJ ? oe - c0 : LF_OBJ;
    __syncthreads();                              // previous chunk fully consumed
    for (int i = t; i < nobj * LF_PIX; i += LF_THREADS) {
      int ol = i / LF_PIX, pp = i - ol * LF_PIX;
      float sv = 0.f;
      if (pp < npx) {
        int o = img_ent[c0 + ol] >> 1;
        float4 bx = *reinterpret_cast<const float4*>(boxes + (int64_t)o * 4);
        float gx = grid_coord(w0 + pp, W, bx.x, 1.f / (bx.z - bx.x));
        float gy = grid_coord(h, H, bx.y, 1.f / (bx.w - bx.y));
        int xl, yl; float wx, wy;
        sv = sample_mask(masks ? masks + (int64_t)o * M * M : nullptr, M, align, gx, gy, xl, yl,
                         wx, wy);
      }
      sS[i] = sv;
    }
    for (int i = t; i < nobj * G; i += LF_THREADS) {
      int ol = i / G, g = i - ol * G;
      int o = img_ent[c0 + ol] >> 1;
      reinterpret_cast<float4*>(sV)[i] = *reinterpret_cast<const float4*>(vecs + (int64_t)o * D + g * 4);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LF_MAXACC; ++k) {
      int idx = t + k * LF_THREADS;
      if (idx < nout) {
        int pp = idx / G, g = idx - pp * G;
        float4 a = acc[k];
        for (int ol = 0; ol < nobj; ++ol) {
          float sv = sS[ol * LF_PIX + pp];
          if (sv != 0.f) {
            float4 v = reinterpret_cast<const float4*>(sV)[ol * G + g];
            a.x += v.x * sv; a.y += v.y * sv; a.z += v.z * sv; a.w += v.w * sv;
          }
        }
        acc[k] = a;
      }
    }
  }
  float* orow = out + ((n * H + h) * (int64_t)W + w0) * ocs;
#pragma unroll
  for (int k = 0; k < LF_MAXACC; ++k) {
    int idx = t + k * LF_THREADS;
    if (idx < nout) {
      int pp = idx / G, g = idx - pp * G;
      float4 o = acc[k];
      if (rnd) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
      *reinterpret_cast<float4*>(orow + (int64_t)pp * ocs + g * 4) = o;
    }
  }
  if (noise) {
    // coalesced read along w (NCHW), padded smem transpose, contiguous write per pixel
    float* sN = lsm;                              // reuse: [LF_PIX][33] per 32-channel slab
    for (int cb = 0; cb < noise_c; cb += 32) {
      __syncthreads();
      for (int i = t; i < 32 * LF_PIX; i += LF_THREADS) {
        int c = i / LF_PIX, pp = i - c * LF_PIX;
        if (cb + c < noise_c && pp < npx)
          sN[pp * 33 + c] = noise[n * nsn + (int64_t)(cb + c) * nsc + (int64_t)h * nsh + (int64_t)(w0 + pp) * nsw];
      }
      __syncthreads();
      for (int i = t; i < 32 * LF_PIX; i += LF_THREADS) {
        int pp = i / 32, c = i - pp * 32;
        if (cb + c < noise_c && pp < npx) {
          float v = sN[pp * 33 + c];
          orow[(int64_t)pp * ocs + D + cb + c] = rnd ? tf32_rn(v) : v;
        }
      }
    }
  }
}

// Generic scalar fallback (D % 4 != 0): one thread per output element.
__global__ void __launch_bounds__(256)
layout_fwd_scalar_kernel(const float* __restrict__ vecs, const float* __restrict__ boxes,
                         const float* __restrict__ masks, int M,
                         const int32_t* __restrict__ img_ptr, const int32_t* __restrict__ img_ent,
                         int64_t N, int64_t D, int64_t H, int64_t W, int align,
                         const float* __restrict__ noise, int64_t noise_c, int64_t nsn, int64_t nsc,
                         int64_t nsh, int64_t nsw, float* __restrict__ out, int64_t ocs, int rnd) {
  int64_t G = D + (noise ? noise_c : 0);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H * W * G) return;
  int64_t c = i % G;
  int64_t pix = i / G;
  int w = (int)(pix % W);
  int64_t t = pix / W;
  int h = (int)(t % H);
  int64_t n = t / H;
  float acc = 0.f;
  if (c < D) {
    for (int32_t k = img_ptr[n]; k < img_ptr[n + 1]; ++k) {
      int o = img_ent[k] >> 1;
      float4 bx = *reinterpret_cast<const float4*>(boxes + (int64_t)o * 4);
      float gx = grid_coord(w, (int)W, bx.x, 1.f / (bx.z - bx.x));
      float gy = grid_coord(h, (int)H, bx.y, 1.f / (bx.w - bx.y));
      int xl, yl; float wx, wy;
      float s = sample_mask(masks ? masks + (int64_t)o * M * M : nullptr, M, align, gx, gy, xl, yl,
                            wx, wy);
      if (s != 0.f) acc += vecs[(int64_t)o * D + c] * s;
    }
  } else {
    acc = noise[n * nsn + (c - D) * nsc + (int64_t)h * nsh + (int64_t)w * nsw];
  }
  out[pix * ocs + c] = rnd ? tf32_rn(acc) : acc;
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
  SG_DYN_SMEM(float, sm);   // [LB_WARPS][D] dvec partials + [M*M] dmask
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
  // conservative pixel bounding box of the object's footprint (the bilinear
  // support reaches at most one mask texel beyond the box); everything outside
  // samples only padding and is skipped without touching memory
  int w_lo = 0, w_hi = (int)W - 1, h_lo = r0, h_hi = r1 - 1;
  {
    float ww = bx.z - bx.x, hh = bx.w - bx.y;
    float mg = 1.f / (float)(M > 1 ? M - 1 : 1);
    if (ww > 0.f && hh > 0.f) {
      float a = floorf((bx.x - mg * ww) * (float)(W - 1)) - 1.f, b = ceilf((bx.z + mg * ww) * (float)(W - 1)) + 1.f;
      float c = floorf((bx.y - mg * hh) * (float)(H - 1)) - 1.f, d = ceilf((bx.w + mg * hh) * (float)(H - 1)) + 1.f;
      if (a > (float)w_lo) w_lo = (int)fminf(a, (float)W);
      if (b < (float)w_hi) w_hi = (int)fmaxf(b, -1.f);
      if (c > (float)h_lo) h_lo = (int)fminf(c, (float)H);
      if (d < (float)h_hi) h_hi = (int)fmaxf(d, -1.f);
    }
  }
  const int bw = w_hi - w_lo + 1, bh = h_hi - h_lo + 1;
  int npix = (bw > 0 && bh > 0) ? bw * bh : 0;
  for (int p = warp; p < npix; p += LB_WARPS) {
    int h = h_lo + p / bw, w = w_lo + p % bw;
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
                                float* out, int64_t out_cstride, int round_tf32,
                                sg2im_stream_t stream) {
  SG_ARG(vecs && boxes && img_row_ptr && img_entries && out);
  SG_ARG(N >= 1 && O >= 1 && D >= 1 && H >= 1 && W >= 1);
  if (!masks) M = 8;
  SG_ARG(M >= 1 && M <= 1024);
  SG_ARG(noise == nullptr || noise_c >= 1);
  int64_t ctot = D + (noise ? noise_c : 0);
  SG_ARG(out_cstride >= ctot);
  SG_ARG(aligned16(boxes));
  bool vec = (D % 4 == 0) && (out_cstride % 4 == 0) && aligned16(vecs) && aligned16(out) &&
             (D / 4) * LF_PIX <= LF_MAXACC * LF_THREADS;
  cudaStream_t st = as_stream(stream);
  const char* v2 = getenv("SG2IM_LAYOUT_V2");            // read per call: tests toggle it in-process
  if (vec && !(v2 && v2[0] == '0') && D <= 1024 && N * ceil_div64(H, 4) < (1ll << 31)) {
    sg2im_layout_fwd_v2(vecs, boxes, masks, M, img_row_ptr, img_entries, N, D, H, W, align_corners,
                        noise, noise_c, nsn, nsc, nsh, nsw, out, out_cstride, round_tf32, st);
    SG_LAUNCH_OK();
    return 0;
  }
  if (vec) {
    int segs = (int)ceil_div64(W, LF_PIX);
    size_t smem = (size_t)(LF_OBJ * LF_PIX + LF_OBJ * D) * sizeof(float);
    size_t need_noise = (size_t)LF_PIX * 33 * sizeof(float);
    if (smem < need_noise) smem = need_noise;
    unsigned grid = (unsigned)(N * H * segs);
    SG_LAUNCH(layout_fwd_kernel, grid, LF_THREADS, smem, st, vecs, boxes, masks, (int)M, img_row_ptr,
                                                      img_entries, N, (int)D, (int)H, (int)W,
                                                      align_corners, noise, (int)noise_c, nsn, nsc,
                                                      nsh, nsw, out, out_cstride, round_tf32);
  } else {
    int64_t total = N * H * W * ctot;
    SG_LAUNCH(layout_fwd_scalar_kernel, (unsigned)ceil_div64(total, 256), 256, 0, st, 
        vecs, boxes, masks, (int)M, img_row_ptr, img_entries, N, D, H, W, align_corners, noise,
        noise_c, nsn, nsc, nsh, nsw, out, out_cstride, round_tf32);
  }
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
  const char* v2 = getenv("SG2IM_LAYOUT_V2");           // read per call: tests toggle it in-process
  if (!(v2 && v2[0] == '0') && ceil_div64(H, 8) <= 65535 && ceil_div64(W, 32) <= 65535 &&
      H * W < (1ll << 31)) {
    sg2im_layout_bwd_v2(dout, dout_cstride, vecs, boxes, masks, M, obj_to_img, N, O, D, H, W,
                        align_corners, dvecs, dmasks, as_stream(stream));
    SG_LAUNCH_OK();
    return 0;
  }
  dim3 grid((unsigned)O, (unsigned)ceil_div64(H, LB_ROWS));
  size_t smem = (size_t)(M * M + LB_WARPS * D) * sizeof(float);
  SG_LAUNCH(layout_bwd_kernel, grid, LB_WARPS * 32, smem, as_stream(stream), 
      dout, dout_cstride, vecs, boxes, masks, (int)M, obj_to_img, N, D, H, W, align_corners, dvecs,
      dmasks);
  SG_LAUNCH_OK();
  return 0;
}

// =============================================================================
// Second-generation layout backward (default since round 2, 447 -> 280 us on the B200;
// SG2IM_LAYOUT_V2=0 selects the first generation).  Same mathematics as layout_bwd_kernel; restructured because the
// first generation is latency-bound and unbalanced (0.44 ms for ~0.45 GB of
// reads, profiles/r01_kernel_table_tf32.txt): there, one CTA owns an object x
// 8-row band however wide the object is (the `__image__` object of every image
// makes 10 % of the CTAs carry >90 % of the pixels), and every warp walks its
// pixels one at a time (one 512 B load in flight, then a 5-step shuffle chain).
// Here
//   * a CTA owns an object x (8 rows x 32 columns) tile, so work per CTA is
//     bounded and the big objects spread over 4x more CTAs;
//   * a warp owns one row of the tile: each LANE evaluates the bilinear mask
//     sample of one of the 32 pixels once (instead of every lane recomputing
//     it for every pixel), a ballot marks the pixels whose footprint touches
//     the mask, and the row is consumed four pixels per iteration with their
//     dout loads issued together.
// =============================================================================
namespace {

constexpr int L2_ROWS = 8, L2_COLS = 32, L2_GROUP = 4;

__global__ void __launch_bounds__(L2_ROWS * 32)
layout_bwd_v2_kernel(const float* __restrict__ dout, int64_t dcs, const float* __restrict__ vecs,
                     const float* __restrict__ boxes, const float* __restrict__ masks, int M,
                     const int64_t* __restrict__ obj_to_img, int64_t N, int D, int H, int W,
                     int align, float* __restrict__ dvecs, float* __restrict__ dmasks) {
  SG_DYN_SMEM(float, sm);   // [L2_ROWS][D] dvec partials + [M*M] dmask
  float* sdv = sm;
  float* sdm = sm + L2_ROWS * D;
  const int o = blockIdx.x;
  const int h = blockIdx.y * L2_ROWS + (threadIdx.x >> 5);      // this warp's image row
  const int w = blockIdx.z * L2_COLS + (threadIdx.x & 31);      // this lane's pixel (sampling phase)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = obj_to_img[o];
  if (n < 0 || n >= N) return;                                   // CTA-uniform
  const float4 bx = *reinterpret_cast<const float4*>(boxes + (int64_t)o * 4);
  const float* mk = masks ? masks + (int64_t)o * M * M : nullptr;
  for (int i = threadIdx.x; i < M * M; i += blockDim.x) sdm[i] = 0.f;
  __syncthreads();

  // lane owns channels j*128 + lane*4 .. +3 (D % 4 == 0, D <= 128*LB_MAXJ: host-checked)
  bool cv[LB_MAXJ];
  float4 vv[LB_MAXJ], acc[LB_MAXJ];
#pragma unroll
  for (int j = 0; j < LB_MAXJ; ++j) {
    cv[j] = j * 128 + lane * 4 < D;
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    vv[j] = cv[j] ? *reinterpret_cast<const float4*>(vecs + (int64_t)o * D + j * 128 + lane * 4)
                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  // phase 1: lane = pixel.  Bilinear footprint of (h, w) in the object's mask.
  int xl = -2, yl = -2; float wx = 0.f, wy = 0.f, s = 0.f;
  bool touch = false;
  if (h < H && w < W) {
    float gx = grid_coord(w, W, bx.x, 1.f / (bx.z - bx.x));
    float gy = grid_coord(h, H, bx.y, 1.f / (bx.w - bx.y));
    s = sample_mask(mk, M, align, gx, gy, xl, yl, wx, wy);
    touch = (xl >= -1 && xl < M) && (yl >= -1 && yl < M);        // else: padding only
  }
  const unsigned active = __ballot_sync(0xffffffffu, touch);

  // phase 2: lane = 4 channels.  Four pixels per iteration.
  if (active) {
    const float* drow = dout + ((n * H + h) * (int64_t)W + (int64_t)blockIdx.z * L2_COLS) * dcs +
                        lane * 4;
    for (int p0 = 0; p0 < L2_COLS; p0 += L2_GROUP) {
      const unsigned grp = (active >> p0) & ((1u << L2_GROUP) - 1u);
      if (!grp) continue;                                        // warp-uniform
      float4 d[L2_GROUP][LB_MAXJ];
#pragma unroll
      for (int q = 0; q < L2_GROUP; ++q) {
        const bool on = (grp >> q) & 1u;
#pragma unroll
        for (int j = 0; j < LB_MAXJ; ++j)
          d[q][j] = (on && cv[j]) ? *reinterpret_cast<const float4*>(drow + (int64_t)(p0 + q) * dcs + j * 128)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float dot[L2_GROUP];
#pragma unroll
      for (int q = 0; q < L2_GROUP; ++q) {
        const float sq = __shfl_sync(0xffffffffu, s, p0 + q);
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < LB_MAXJ; ++j) {
          t += d[q][j].x * vv[j].x + d[q][j].y * vv[j].y + d[q][j].z * vv[j].z + d[q][j].w * vv[j].w;
          acc[j].x += d[q][j].x * sq; acc[j].y += d[q][j].y * sq;
          acc[j].z += d[q][j].z * sq; acc[j].w += d[q][j].w * sq;
        }
        dot[q] = t;
      }
      if (dmasks) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
          for (int q = 0; q < L2_GROUP; ++q) dot[q] += __shfl_xor_sync(0xffffffffu, dot[q], off);
        }
        // lanes 0..15: pixel q = lane/4, mask corner = lane%4
        const int q = (lane >> 2) & (L2_GROUP - 1), corner = lane & 3;
        const int src = p0 + q;
        const int pxl = __shfl_sync(0xffffffffu, xl, src), pyl = __shfl_sync(0xffffffffu, yl, src);
        const float pwx = __shfl_sync(0xffffffffu, wx, src), pwy = __shfl_sync(0xffffffffu, wy, src);
        float dq = dot[0];
#pragma unroll
        for (int k = 1; k < L2_GROUP; ++k) dq = (q == k) ? dot[k] : dq;
        if (lane < 4 * L2_GROUP && ((grp >> q) & 1u)) {
          const int cx = pxl + (corner & 1), cy = pyl + (corner >> 1);
          const float wgt = ((corner & 1) ? pwx : 1.f - pwx) * ((corner >> 1) ? pwy : 1.f - pwy);
          if (cx >= 0 && cx < M && cy >= 0 && cy < M) atomicAdd(&sdm[cy * M + cx], wgt * dq);
        }
      }
    }
  }

  // cross-warp reduction of dvec, then one atomic per channel per CTA
#pragma unroll
  for (int j = 0; j < LB_MAXJ; ++j)
    if (cv[j]) *reinterpret_cast<float4*>(&sdv[warp * D + j * 128 + lane * 4]) = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float t = 0.f;
    for (int wv = 0; wv < L2_ROWS; ++wv) t += sdv[wv * D + c];
    if (t != 0.f) atomicAdd(dvecs + (int64_t)o * D + c, t);
  }
  if (dmasks) {
    for (int i = threadIdx.x; i < M * M; i += blockDim.x) {
      float t = sdm[i];
      if (t != 0.f) atomicAdd(dmasks + (int64_t)o * M * M + i, t);
    }
  }
}

}  // namespace

// =============================================================================
// Second-generation layout forward (default since round 2, 269 -> 195 us; SG2IM_LAYOUT_V2=0: first generation).  The first
// generation launches one CTA per (image, row, 32-pixel segment) = 16 384 CTAs
// at VG-128, and each of them walks the same dependent chain of global loads
// (image -> object list -> boxes / mask texels / vectors) before it can write
// 20 KB: 0.26 ms for 0.4 GB (profiles/r01_kernel_table_tf32.txt).  Here a CTA
// owns an (image, 4-row band): the image's object data (boxes, vectors, masks
// up to 16x16) is staged in shared memory ONCE, then the band's 32-pixel
// segments are produced from shared memory only; the noise channels of a
// segment are fetched into registers before the segment's arithmetic starts.
// Images with more than 16 objects take further passes that accumulate into the
// output (read-modify-write); TF32 rounding is applied by the last pass only.
// =============================================================================
namespace {

constexpr int F2_ROWS = 4, F2_OBJ = 16, F2_MAXM = 16, F2_THREADS = 256, F2_MAXACC = 8;
constexpr int F2_NOISE_PER_THREAD = (32 * LF_PIX) / F2_THREADS;   // noise slab of 32 channels x 32 px

__global__ void __launch_bounds__(F2_THREADS)
layout_fwd_v2_kernel(const float* __restrict__ vecs, const float* __restrict__ boxes,
                     const float* __restrict__ masks, int M, const int32_t* __restrict__ img_ptr,
                     const int32_t* __restrict__ img_ent, int D, int H, int W, int align,
                     const float* __restrict__ noise, int noise_c, int64_t nsn, int64_t nsc,
                     int64_t nsh, int64_t nsw, float* __restrict__ out, int64_t ocs, int rnd) {
  SG_DYN_SMEM(float, f2sm);
  float* sV = f2sm;                               // [F2_OBJ][D]
  float* sS = sV + F2_OBJ * D;                    // [F2_OBJ][LF_PIX]
  float* sN = sS + F2_OBJ * LF_PIX;               // [LF_PIX][33] noise transpose tile
  float* sM = sN + LF_PIX * 33;                   // [F2_OBJ][M*M] when M <= F2_MAXM
  __shared__ float4 sBox[F2_OBJ];
  __shared__ int sObj[F2_OBJ];
  const int bands = (H + F2_ROWS - 1) / F2_ROWS;
  const int band = blockIdx.x % bands;
  const int64_t n = blockIdx.x / bands;
  const int h0 = band * F2_ROWS;
  const int h1 = h0 + F2_ROWS < H ? h0 + F2_ROWS : H;
  const int segs = (W + LF_PIX - 1) / LF_PIX;
  const int t = threadIdx.x;
  const int G = D >> 2;
  const bool mask_in_smem = masks != nullptr && M <= F2_MAXM;
  const int32_t ob = img_ptr[n], oe = img_ptr[n + 1];
  const int32_t npass = oe > ob ? (oe - ob + F2_OBJ - 1) / F2_OBJ : 1;   // an empty image still writes zeros

  for (int32_t pass = 0; pass < npass; ++pass) {
    const int32_t c0 = ob + pass * F2_OBJ;
    const int nobj = oe - c0 < F2_OBJ ? (oe - c0 > 0 ? oe - c0 : 0) : F2_OBJ;
    const bool first = pass == 0, last = pass == npass - 1;
    __syncthreads();                              // previous pass fully consumed
    if (t < nobj) {
      int o = img_ent[c0 + t] >> 1;
      sObj[t] = o;
      sBox[t] = *reinterpret_cast<const float4*>(boxes + (int64_t)o * 4);
    }
    __syncthreads();
    for (int i = t; i < nobj * G; i += F2_THREADS) {
      int ol = i / G, g = i - ol * G;
      reinterpret_cast<float4*>(sV)[i] =
          *reinterpret_cast<const float4*>(vecs + (int64_t)sObj[ol] * D + g * 4);
    }
    if (mask_in_smem) {
      const int MM = M * M;
      for (int i = t; i < nobj * MM; i += F2_THREADS) {
        int ol = i / MM, r = i - ol * MM;
        sM[i] = masks[(int64_t)sObj[ol] * MM + r];
      }
    }
    __syncthreads();

    for (int h = h0; h < h1; ++h) {
      for (int seg = 0; seg < segs; ++seg) {
        const int w0 = seg * LF_PIX;
        const int npx = W - w0 < LF_PIX ? W - w0 : LF_PIX;
        // noise of this segment (first 32-channel slab) into registers before anything else
        float nz[F2_NOISE_PER_THREAD];
        if (noise && first) {
#pragma unroll
          for (int k = 0; k < F2_NOISE_PER_THREAD; ++k) {
            int i = t + k * F2_THREADS;
            int c = i / LF_PIX, pp = i - c * LF_PIX;
            nz[k] = (c < noise_c && pp < npx)
                        ? noise[n * nsn + (int64_t)c * nsc + (int64_t)h * nsh + (int64_t)(w0 + pp) * nsw]
                        : 0.f;
          }
        }
        // phase 1: S[o][p], one bilinear mask sample per (object, pixel)
        for (int i = t; i < nobj * LF_PIX; i += F2_THREADS) {
          int ol = i / LF_PIX, pp = i - ol * LF_PIX;
          float sv = 0.f;
          if (pp < npx) {
            float4 bx = sBox[ol];
            float gx = grid_coord(w0 + pp, W, bx.x, 1.f / (bx.z - bx.x));
            float gy = grid_coord(h, H, bx.y, 1.f / (bx.w - bx.y));
            int xl, yl; float wx, wy;
            const float* mk = !masks ? nullptr
                              : (mask_in_smem ? sM + ol * M * M : masks + (int64_t)sObj[ol] * M * M);
            sv = sample_mask(mk, M, align, gx, gy, xl, yl, wx, wy);
          }
          sS[i] = sv;
        }
        __syncthreads();
        // phase 2: out[p][c..c+3] (+)= sum_o S[o][p] * vec[o][c..c+3]
        float* orow = out + ((n * H + h) * (int64_t)W + w0) * ocs;
        const int nout = npx * G;
#pragma unroll
        for (int k = 0; k < F2_MAXACC; ++k) {
          int idx = t + k * F2_THREADS;
          if (idx < nout) {
            int pp = idx / G, g = idx - pp * G;
            float4* dst = reinterpret_cast<float4*>(orow + (int64_t)pp * ocs + g * 4);
            float4 a = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *dst;
            for (int ol = 0; ol < nobj; ++ol) {
              float sv = sS[ol * LF_PIX + pp];
              if (sv != 0.f) {
                float4 v = reinterpret_cast<const float4*>(sV)[ol * G + g];
                a.x += v.x * sv; a.y += v.y * sv; a.z += v.z * sv; a.w += v.w * sv;
              }
            }
            if (rnd && last) { a.x = tf32_rn(a.x); a.y = tf32_rn(a.y); a.z = tf32_rn(a.z); a.w = tf32_rn(a.w); }
            *dst = a;
          }
        }
        if (noise && first) {
          // NCHW -> NHWC through a padded tile; slabs beyond the first 32 channels are fetched here
          for (int cb = 0; cb < noise_c; cb += 32) {
#pragma unroll
            for (int k = 0; k < F2_NOISE_PER_THREAD; ++k) {
              int i = t + k * F2_THREADS;
              int c = i / LF_PIX, pp = i - c * LF_PIX;
              float v = nz[k];
              if (cb > 0)
                v = (cb + c < noise_c && pp < npx)
                        ? noise[n * nsn + (int64_t)(cb + c) * nsc + (int64_t)h * nsh + (int64_t)(w0 + pp) * nsw]
                        : 0.f;
              sN[pp * 33 + c] = v;
            }
            __syncthreads();
            for (int i = t; i < 32 * LF_PIX; i += F2_THREADS) {
              int pp = i / 32, c = i - pp * 32;
              if (cb + c < noise_c && pp < npx) {
                float v = sN[pp * 33 + c];
                orow[(int64_t)pp * ocs + D + cb + c] = rnd ? tf32_rn(v) : v;
              }
            }
            __syncthreads();                      // sN (and sS) free for the next slab / segment
          }
        } else {
          __syncthreads();                        // sS free for the next segment
        }
      }
    }
  }
}

}  // namespace

// preconditions checked by sg2im_layout_fwd (float4 path)
int sg2im_layout_fwd_v2(const float* vecs, const float* boxes, const float* masks, int64_t M,
                        const int32_t* img_ptr, const int32_t* img_ent, int64_t N, int64_t D,
                        int64_t H, int64_t W, int align, const float* noise, int64_t noise_c,
                        int64_t nsn, int64_t nsc, int64_t nsh, int64_t nsw, float* out, int64_t ocs,
                        int rnd, cudaStream_t st) {
  const bool stage_masks = masks != nullptr && M <= F2_MAXM;
  size_t smem = (size_t)(F2_OBJ * D + F2_OBJ * LF_PIX + LF_PIX * 33 +
                         (stage_masks ? F2_OBJ * M * M : 0)) * sizeof(float);
#ifndef SG2IM_EMUL
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(layout_fwd_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
#endif
  int64_t bands = ceil_div64(H, F2_ROWS);
  SG_LAUNCH(layout_fwd_v2_kernel, (unsigned)(N * bands), F2_THREADS, smem, st, 
      vecs, boxes, masks, (int)M, img_ptr, img_ent, (int)D, (int)H, (int)W, align, noise,
      (int)noise_c, nsn, nsc, nsh, nsw, out, ocs, rnd);
  return 0;
}

// preconditions checked by sg2im_layout_bwd
int sg2im_layout_bwd_v2(const float* dout, int64_t dcs, const float* vecs, const float* boxes,
                        const float* masks, int64_t M, const int64_t* obj_to_img, int64_t N,
                        int64_t O, int64_t D, int64_t H, int64_t W, int align, float* dvecs,
                        float* dmasks, cudaStream_t st) {
  dim3 grid((unsigned)O, (unsigned)ceil_div64(H, L2_ROWS), (unsigned)ceil_div64(W, L2_COLS));
  size_t smem = (size_t)(M * M + L2_ROWS * D) * sizeof(float);
  SG_LAUNCH(layout_bwd_v2_kernel, grid, L2_ROWS * 32, smem, st, dout, dcs, vecs, boxes, masks, (int)M,
                                                         obj_to_img, N, (int)D, (int)H, (int)W,
                                                         align, dvecs, dmasks);
  return 0;
}
