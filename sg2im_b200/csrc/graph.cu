// Scene-graph convolution data movement: CSR build, triple gather, ordered
// segment sum.  Replaces the gather / scatter_add sequence of
// sg2im/graph.py:71-114 (and _pool_samples' index build, layout.py:146-148).
// HBM/latency-bound integer + fp32 work; no tensor cores.
#include "common.cuh"

namespace {

// one warp per destination row: count hits (pass 0) or write entries in
// (role, t) order (pass 1) using ballot + popc for a stable compaction.
template <int PASS>
__global__ void csr_scan_rows(const int64_t* __restrict__ idx, int64_t T, int64_t stride,
                              int nroles, int64_t R, int32_t* __restrict__ row_ptr,
                              int32_t* __restrict__ entries) {
  int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  int lane = threadIdx.x & 31;
  int32_t pos = PASS == 1 ? row_ptr[r] : 0;
  for (int role = 0; role < nroles; ++role) {
    for (int64_t base = 0; base < T; base += 32) {
      int64_t t = base + lane;
      bool hit = t < T && idx[t * stride + role] == r;
      unsigned m = __ballot_sync(0xffffffffu, hit);
      if (PASS == 1 && hit) {
        int rank = __popc(m & ((1u << lane) - 1u));
        entries[pos + rank] = (int32_t)(t * 2 + role);
      }
      pos += __popc(m);
    }
  }
  if (PASS == 0 && lane == 0) row_ptr[r + 1] = pos;       // count, scanned next
}

// single-block in-place scan: row_ptr[0]=0, row_ptr[r+1] = sum counts[0..r]
__global__ void csr_scan_counts(int32_t* __restrict__ row_ptr, int64_t R) {
  __shared__ int32_t part[1024];
  int tid = threadIdx.x;
  int64_t per = (R + blockDim.x - 1) / blockDim.x;
  int64_t b = tid * per, e = b + per < R ? b + per : R;
  int32_t s = 0;
  for (int64_t i = b; i < e; ++i) s += row_ptr[i + 1];
  part[tid] = s;
  __syncthreads();
  for (int off = 1; off < (int)blockDim.x; off <<= 1) {
    int32_t v = tid >= off ? part[tid - off] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int32_t run = part[tid] - s;                             // exclusive prefix
  if (tid == 0) row_ptr[0] = 0;
  for (int64_t i = b; i < e; ++i) { run += row_ptr[i + 1]; row_ptr[i + 1] = run; }
}

template <int VEC>
__global__ void triple_gather_kernel(const float* __restrict__ rows, const float* __restrict__ mid,
                                     const int64_t* __restrict__ edges, int64_t T, int64_t Wr,
                                     int64_t Wm, const int32_t* __restrict__ row_ptr,
                                     float* __restrict__ out) {
  int64_t roww = 2 * Wr + Wm;
  int64_t per_row = roww / VEC;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * per_row) return;
  int64_t t = i / per_row;
  int64_t c = (i - t * per_row) * VEC;
  float v[VEC];
  if (c >= Wr && c < Wr + Wm) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = mid ? mid[t * Wm + (c - Wr) + j] : 0.f;
  } else {
    int role = c >= Wr ? 1 : 0;
    int64_t cc = role ? c - Wr - Wm : c;
    int64_t r = edges[t * 2 + role];
    float cnt = 1.f;
    if (row_ptr) { int32_t n = row_ptr[r + 1] - row_ptr[r]; cnt = (float)(n > 1 ? n : 1); }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float x = rows[r * Wr + cc + j];
      v[j] = row_ptr ? x / cnt : x;
    }
  }
  if (VEC == 4) {
    *reinterpret_cast<float4*>(out + t * roww + c) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    out[t * roww + c] = v[0];
  }
}

template <int VEC>
__global__ void segment_sum_kernel(const float* __restrict__ src, int64_t stride, int64_t off0,
                                   int64_t off1, int64_t W, const int32_t* __restrict__ row_ptr,
                                   const int32_t* __restrict__ entries, int64_t R, int avg,
                                   float* __restrict__ out) {
  int64_t per_row = W / VEC;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * per_row) return;
  int64_t r = i / per_row;
  int64_t c = (i - r * per_row) * VEC;
  int32_t b = row_ptr[r], e = row_ptr[r + 1];
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  // strictly sequential fp32 adds in CSR order: the bit-exact contract
  for (int32_t k = b; k < e; ++k) {
    int32_t en = entries[k];
    const float* p = src + (int64_t)(en >> 1) * stride + ((en & 1) ? off1 : off0) + c;
    if (VEC == 4) {
      float4 x = *reinterpret_cast<const float4*>(p);
      acc[0] = __fadd_rn(acc[0], x.x); acc[1] = __fadd_rn(acc[1], x.y);
      acc[2] = __fadd_rn(acc[2], x.z); acc[3] = __fadd_rn(acc[3], x.w);
    } else {
      acc[0] = __fadd_rn(acc[0], p[0]);
    }
  }
  if (avg) {
    float cnt = (float)((e - b) > 1 ? (e - b) : 1);
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = __fdiv_rn(acc[j], cnt);
  }
  if (VEC == 4) {
    *reinterpret_cast<float4*>(out + r * W + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  } else {
    out[r * W + c] = acc[0];
  }
}

}  // namespace

extern "C" int sg2im_csr_build(const int64_t* idx, int64_t T, int64_t idx_stride, int nroles,
                               int64_t R, int32_t* row_ptr, int32_t* entries,
                               sg2im_stream_t stream) {
  SG_ARG(idx != nullptr || T == 0);
  SG_ARG(row_ptr != nullptr && (entries != nullptr || T == 0));
  SG_ARG(T >= 0 && R >= 1 && (nroles == 1 || nroles == 2) && idx_stride >= nroles);
  SG_ARG(T < (1ll << 30));
  cudaStream_t st = as_stream(stream);
  const int warps = 8;
  unsigned grid = (unsigned)ceil_div64(R, warps);
  SG_LAUNCH(csr_scan_rows<0>, grid, warps * 32, 0, st, idx, T, idx_stride, nroles, R, row_ptr, entries);
  SG_LAUNCH(csr_scan_counts, 1, 1024, 0, st, row_ptr, R);
  SG_LAUNCH(csr_scan_rows<1>, grid, warps * 32, 0, st, idx, T, idx_stride, nroles, R, row_ptr, entries);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_triple_gather(const float* rows, const float* mid, const int64_t* edges,
                                   int64_t T, int64_t Wr, int64_t Wm, const int32_t* row_ptr,
                                   float* out, sg2im_stream_t stream) {
  SG_ARG(rows && edges && out);
  SG_ARG(T >= 0 && Wr >= 1 && Wm >= 0);
  if (T == 0) return 0;
  cudaStream_t st = as_stream(stream);
  bool vec = (Wr % 4 == 0) && (Wm % 4 == 0) && aligned16(rows) && aligned16(out) &&
             (mid == nullptr || aligned16(mid));
  int64_t total = T * (2 * Wr + Wm) / (vec ? 4 : 1);
  unsigned grid = (unsigned)ceil_div64(total, 256);
  if (vec) SG_LAUNCH(triple_gather_kernel<4>, grid, 256, 0, st, rows, mid, edges, T, Wr, Wm, row_ptr, out);
  else     SG_LAUNCH(triple_gather_kernel<1>, grid, 256, 0, st, rows, mid, edges, T, Wr, Wm, row_ptr, out);
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_segment_sum(const float* src, int64_t src_stride, int64_t off0, int64_t off1,
                                 int64_t W, const int32_t* row_ptr, const int32_t* entries,
                                 int64_t R, int avg, float* out, sg2im_stream_t stream) {
  SG_ARG(src && row_ptr && entries && out);
  SG_ARG(R >= 1 && W >= 1 && off0 >= 0 && off1 >= 0 && src_stride >= W);
  cudaStream_t st = as_stream(stream);
  bool vec = (W % 4 == 0) && (src_stride % 4 == 0) && (off0 % 4 == 0) && (off1 % 4 == 0) &&
             aligned16(src) && aligned16(out);
  int64_t total = R * W / (vec ? 4 : 1);
  unsigned grid = (unsigned)ceil_div64(total, 128);
  if (vec) SG_LAUNCH(segment_sum_kernel<4>, grid, 128, 0, st, src, src_stride, off0, off1, W, row_ptr, entries, R, avg, out);
  else     SG_LAUNCH(segment_sum_kernel<1>, grid, 128, 0, st, src, src_stride, off0, off1, W, row_ptr, entries, R, avg, out);
  SG_LAUNCH_OK();
  return 0;
}
