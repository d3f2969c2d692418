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

__global__ void __launch_bounds__(256)
pack_oihw_kernel(const float* __restrict__ w, int Co, int Ci, int CiUse, int T,
                 float* __restrict__ fwd, float* __restrict__ dgr) {
  extern __shared__ float sm[];                 // [PT co][PT ci * T + 1]
  const int ld = PT * T + 1;
  const int co0 = blockIdx.y * PT, ci0 = blockIdx.x * PT;
  const int nci = min(PT, CiUse - ci0), nco = min(PT, Co - co0);
  // coalesced read: for each co a run of nci*T contiguous floats
  for (int i = threadIdx.x; i < nco * nci * T; i += blockDim.x) {
    int c = i / (nci * T), r = i - c * (nci * T);
    sm[c * ld + r] = w[((size_t)(co0 + c) * Ci + ci0) * T + r];
  }
  __syncthreads();
  if (fwd) {
    for (int i = threadIdx.x; i < T * nco * nci; i += blockDim.x) {
      int ci = i % nci; int r = i / nci; int co = r % nco; int t = r / nco;
      fwd[((size_t)t * Co + co0 + co) * CiUse + ci0 + ci] = sm[co * ld + ci * T + t];
    }
  }
  if (dgr) {
    for (int i = threadIdx.x; i < T * nci * nco; i += blockDim.x) {
      int co = i % nco; int r = i / nco; int ci = r % nci; int t = r / nci;
      dgr[((size_t)(T - 1 - t) * CiUse + ci0 + ci) * Co + co0 + co] = sm[co * ld + ci * T + t];
    }
  }
}

__global__ void __launch_bounds__(256)
unpack_wgrad_kernel(const float* __restrict__ dw, int Co, int Ci, int CiUse, int T,
                    float* __restrict__ grad, int accumulate) {
  extern __shared__ float sm[];                 // [PT co][PT ci * T + 1]
  const int ld = PT * T + 1;
  const int co0 = blockIdx.y * PT, ci0 = blockIdx.x * PT;
  const int nci = min(PT, CiUse - ci0), nco = min(PT, Co - co0);
  for (int i = threadIdx.x; i < T * nci * nco; i += blockDim.x) {
    int co = i % nco; int r = i / nco; int ci = r % nci; int t = r / nci;
    sm[co * ld + ci * T + t] = dw[((size_t)t * CiUse + ci0 + ci) * Co + co0 + co];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nco * nci * T; i += blockDim.x) {
    int c = i / (nci * T), r = i - c * (nci * T);
    size_t o = ((size_t)(co0 + c) * Ci + ci0) * T + r;
    grad[o] = accumulate ? grad[o] + sm[c * ld + r] : sm[c * ld + r];
  }
}
}  // namespace

extern "C" int sg2im_pack_weights(const float* w, int64_t Cout, int64_t Cin, int64_t cin_use,
                                  int64_t taps, float* w_fwd, float* w_dgrad,
                                  sg2im_stream_t stream) {
  SG_ARG(w && (w_fwd || w_dgrad));
  SG_ARG(Cout >= 1 && Cin >= 1 && cin_use >= 1 && cin_use <= Cin && taps >= 1 && taps <= 64);
  dim3 grid((unsigned)ceil_div64(cin_use, PT), (unsigned)ceil_div64(Cout, PT));
  size_t smem = (size_t)PT * (PT * taps + 1) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pack_oihw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(unpack_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  pack_oihw_kernel<<<grid, 256, smem, as_stream(stream)>>>(w, (int)Cout, (int)Cin, (int)cin_use,
                                                           (int)taps, w_fwd, w_dgrad);
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
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pack_oihw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(unpack_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  unpack_wgrad_kernel<<<grid, 256, smem, as_stream(stream)>>>(dw, (int)Cout, (int)Cin, (int)cin_use,
                                                              (int)taps, grad_oihw, accumulate);
  SG_LAUNCH_OK();
  return 0;
}
