// Weight layout conversion between the checkpoint-compatible OIHW fp32 master
// parameters and the kernel layouts, tiled through shared memory so both the
// read and the writes are 128-byte coalesced (torch's permute().contiguous()
// copies ran at ~10 % of HBM bandwidth and were 1.4 ms of a 17 ms step):
//   pack:   w[co][ci][t]  ->  fwd[t][co][ci]            (tcgen05 forward B operand)
//                             dgr[T-1-t][ci][co]        (data-gradient B operand:
//                                                        flipped taps, channels swapped)
//   unpack: dw[t][ci][co] ->  grad[co][ci][t]  (+= optional)   (wgrad output -> OIHW)
#include "common.cuh"

namespace {
constexpr int PT = 32;          // 32 x 32 (co x ci) tile, all taps

// T is a template parameter (0 = runtime) so that every index split is a
// division by a constant; ragged edge tiles take the same code with bounds.
template <int TT>
__global__ void __launch_bounds__(256)
pack_oihw_kernel(const float* __restrict__ w, int Co, int Ci, int CiUse, int Trt,
                 float* __restrict__ fwd, float* __restrict__ dgr, int rnd) {
  SG_DYN_SMEM(float, sm);                 // [PT co][PT * T + 1]
  const int T = TT ? TT : Trt;
  const int ld = PT * T + 1;
  const int co0 = blockIdx.y * PT, ci0 = blockIdx.x * PT;
  const int nci = min(PT, CiUse - ci0), nco = min(PT, Co - co0);
  const int run = PT * T;                       // floats per co row of a full tile
  // coalesced read: for each co a run of nci*T contiguous floats
  for (int i = threadIdx.x; i < PT * run; i += 256) {
    int c = i / run, r = i - c * run;
    if (c < nco && r < nci * T) {
      float v = w[((size_t)(co0 + c) * Ci + ci0) * T + r];
      sm[c * ld + r] = rnd ? tf32_rn(v) : v;
    }
  }
  __syncthreads();
  if (fwd) {
    for (int i = threadIdx.x; i < T * PT * PT; i += 256) {
      int ci = i % PT; int r = i / PT; int co = r % PT; int t = r / PT;
      if (ci < nci && co < nco)
        fwd[((size_t)t * Co + co0 + co) * CiUse + ci0 + ci] = sm[co * ld + ci * T + t];
    }
  }
  if (dgr) {
    for (int i = threadIdx.x; i < T * PT * PT; i += 256) {
      int co = i % PT; int r = i / PT; int ci = r % PT; int t = r / PT;
      if (ci < nci && co < nco)
        dgr[((size_t)(T - 1 - t) * CiUse + ci0 + ci) * Co + co0 + co] = sm[co * ld + ci * T + t];
    }
  }
}

template <int TT>
__global__ void __launch_bounds__(256)
unpack_wgrad_kernel(const float* __restrict__ dw, int Co, int Ci, int CiUse, int Trt,
                    float* __restrict__ grad, int accumulate) {
  SG_DYN_SMEM(float, sm);
  const int T = TT ? TT : Trt;
  const int ld = PT * T + 1;
  const int co0 = blockIdx.y * PT, ci0 = blockIdx.x * PT;
  const int nci = min(PT, CiUse - ci0), nco = min(PT, Co - co0);
  const int run = PT * T;
  for (int i = threadIdx.x; i < T * PT * PT; i += 256) {
    int co = i % PT; int r = i / PT; int ci = r % PT; int t = r / PT;
    if (ci < nci && co < nco)
      sm[co * ld + ci * T + t] = dw[((size_t)t * CiUse + ci0 + ci) * Co + co0 + co];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PT * run; i += 256) {
    int c = i / run, r = i - c * run;
    if (c < nco && r < nci * T) {
      size_t o = ((size_t)(co0 + c) * Ci + ci0) * T + r;
      grad[o] = accumulate ? grad[o] + sm[c * ld + r] : sm[c * ld + r];
    }
  }
}

template <int TT>
void launch_pack(dim3 grid, size_t smem, cudaStream_t st, const float* w, int Co, int Ci, int cu,
                 int T, float* f, float* d, int rnd) {
#ifndef SG2IM_EMUL
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pack_oihw_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
#endif
  SG_LAUNCH(pack_oihw_kernel<TT>, grid, 256, smem, st, w, Co, Ci, cu, T, f, d, rnd);
}
template <int TT>
void launch_unpack(dim3 grid, size_t smem, cudaStream_t st, const float* dw, int Co, int Ci, int cu,
                   int T, float* g, int acc) {
#ifndef SG2IM_EMUL
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(unpack_wgrad_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
#endif
  SG_LAUNCH(unpack_wgrad_kernel<TT>, grid, 256, smem, st, dw, Co, Ci, cu, T, g, acc);
}
}  // namespace

extern "C" int sg2im_pack_weights(const float* w, int64_t Cout, int64_t Cin, int64_t cin_use,
                                  int64_t taps, float* w_fwd, float* w_dgrad, int round_tf32,
                                  sg2im_stream_t stream) {
  SG_ARG(w && (w_fwd || w_dgrad));
  SG_ARG(Cout >= 1 && Cin >= 1 && cin_use >= 1 && cin_use <= Cin && taps >= 1 && taps <= 64);
  dim3 grid((unsigned)ceil_div64(cin_use, PT), (unsigned)ceil_div64(Cout, PT));
  size_t smem = (size_t)PT * (PT * taps + 1) * sizeof(float);
  cudaStream_t st = as_stream(stream);
  int Co = (int)Cout, Ci = (int)Cin, cu = (int)cin_use, T = (int)taps;
  switch (T) {
    case 1: launch_pack<1>(grid, smem, st, w, Co, Ci, cu, T, w_fwd, w_dgrad, round_tf32); break;
    case 4: launch_pack<4>(grid, smem, st, w, Co, Ci, cu, T, w_fwd, w_dgrad, round_tf32); break;
    case 9: launch_pack<9>(grid, smem, st, w, Co, Ci, cu, T, w_fwd, w_dgrad, round_tf32); break;
    case 16: launch_pack<16>(grid, smem, st, w, Co, Ci, cu, T, w_fwd, w_dgrad, round_tf32); break;
    default: launch_pack<0>(grid, smem, st, w, Co, Ci, cu, T, w_fwd, w_dgrad, round_tf32); break;
  }
  SG_LAUNCH_OK();
  return 0;
}

extern "C" int sg2im_unpack_wgrad(const float* dw, int64_t Cout, int64_t Cin, int64_t cin_use,
                                  int64_t taps, float* grad_oihw, int accumulate,
                                  sg2im_stream_t stream) {
  SG_ARG(dw && grad_oihw);
  SG_ARG(Cout >= 1 && Cin >= 1 && cin_use >= 1 && cin_use <= Cin && taps >= 1 && taps <= 64);
  dim3 grid((unsigned)ceil_div64(cin_use, PT), (unsigned)ceil_div64(Cout, PT));
  size_t smem = (size_t)PT * (PT * taps + 1) * sizeof(float);
  cudaStream_t st = as_stream(stream);
  int Co = (int)Cout, Ci = (int)Cin, cu = (int)cin_use, T = (int)taps;
  switch (T) {
    case 1: launch_unpack<1>(grid, smem, st, dw, Co, Ci, cu, T, grad_oihw, accumulate); break;
    case 4: launch_unpack<4>(grid, smem, st, dw, Co, Ci, cu, T, grad_oihw, accumulate); break;
    case 9: launch_unpack<9>(grid, smem, st, dw, Co, Ci, cu, T, grad_oihw, accumulate); break;
    case 16: launch_unpack<16>(grid, smem, st, dw, Co, Ci, cu, T, grad_oihw, accumulate); break;
    default: launch_unpack<0>(grid, smem, st, dw, Co, Ci, cu, T, grad_oihw, accumulate); break;
  }
  SG_LAUNCH_OK();
  return 0;
}
