// Weight layout conversion between the checkpoint-compatible OIHW fp32 master
// parameters and the kernel layouts, tiled through shared memory so both the
// read and the writes are 128-byte coalesced (torch's permute().contiguous()
// copies ran at ~10 % of HBM bandwidth and were 1.4 ms of a 17 ms step):
//   pack:   w[co][ci][t]  ->  fwd[t][co][ci]            (tcgen05 forward B operand)
//                             dgr[T-1-t][ci][co]        (data-gradient B operand:
//                                                        flipped taps, channels swapped)
//   unpack: dw[t][ci][co] ->  grad[co][ci][t]  (+= optional)   (wgrad output -> OIHW)
//   split:  kcc[t][ci][co] -> bf16 hi / mid operand copies of ALL convolution weights of a
//           network in one launch (the 'bf16x3' arithmetic's pre-split B operands, below)
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

// ----------------------------------------------------------------------------------------------
// Pre-split weight operands of the bf16 arithmetic.  The tensor-core kernels can split fp32
// operand tiles themselves (converter warps, tc_common.cuh), but weights are constant within a
// training step and every CTA re-reads them: splitting them once per step removes that work —
// and its shared-memory traffic, which is what bounds those kernels — from every tile of every
// launch.  One launch per network: entry e describes a weight stored tap-major / input channel /
// output channel fastest ("kcc", what the weight-gradient kernel writes and Adam updates) and the
// two operand copies, each a K-major matrix whose rows hold 32-channel blocks
// [32 x bf16 hi | 32 x bf16 mid] (the layout the converter warps produce in shared memory):
//   fwd[t][co][ci blocks]          B operand of the forward convolution
//   dgr[T-1-t][ci][co blocks]      B operand of the data gradient (flipped taps)
// Pad channels inside the last block are written as zeros.
namespace {
struct SplitEntry {           // 8 x int64, built by the host (ops.SplitShadows)
  long long src, fwd, dgr, taps, Cin, Cout, first_tile, s2d_c;
  // s2d_c > 0: src is a 4x4 stride-2 filter [16][s2d_c][Cout]; the copies are those of the
  // equivalent 2x2 stride-1 filter on the space-to-depth input: taps = 4 (ty, tx),
  // Cin = 4 * s2d_c with channel (py * 2 + px) * s2d_c + c  <->  filter tap (2 ty + py, 2 tx + px)
};

__global__ void __launch_bounds__(256)
split_weights_kernel(const SplitEntry* __restrict__ tab, int n_entries) {
  __shared__ float tile[32][33];
  const int bid = (int)blockIdx.x;
  int e = 0, hi = n_entries - 1;                       // last entry whose first_tile <= bid
  while (e < hi) {
    const int mid = (e + hi + 1) >> 1;
    if ((long long)bid >= tab[mid].first_tile) e = mid; else hi = mid - 1;
  }
  const SplitEntry E = tab[e];
  const int Cin = (int)E.Cin, Cout = (int)E.Cout, T = (int)E.taps;
  const int tci = (Cin + 31) >> 5, tco = (Cout + 31) >> 5;
  int r = bid - (int)E.first_tile;
  const int bco = r % tco; r /= tco;
  const int bci = r % tci; const int tap = r / tci;
  const int ci0 = bci * 32, co0 = bco * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* src0 = reinterpret_cast<const float*>(E.src);
  const int sc = (int)E.s2d_c;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ci = ci0 + warp + 8 * k;
    size_t row = (size_t)tap * Cin + ci;                 // row of the [taps][Cin][Cout] master
    if (sc > 0) {
      const int q = ci / sc, c = ci - q * sc;            // (py, px) block of the space-to-depth channels
      row = (size_t)((2 * (tap >> 1) + (q >> 1)) * 4 + 2 * (tap & 1) + (q & 1)) * sc + c;
    }
    tile[warp + 8 * k][lane] = (ci < Cin && co0 + lane < Cout) ? src0[row * Cout + co0 + lane] : 0.f;
  }
  __syncthreads();
  const int cin_pad = tci * 32, cout_pad = tco * 32;
  uint32_t* fwd = reinterpret_cast<uint32_t*>(E.fwd);
  uint32_t* dgr = reinterpret_cast<uint32_t*>(E.dgr);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int row = warp + 8 * k;
    if (dgr && ci0 + row < Cin) {                       // row = input channel, lanes = output channels
      const float v = tile[row][lane], nb = __shfl_xor_sync(0xffffffffu, v, 1);
      uint32_t hi, mid;
      split_bf16x2(v, nb, hi, mid);
      if (!(lane & 1)) {
        uint32_t* d = dgr + ((size_t)(T - 1 - tap) * Cin + ci0 + row) * cout_pad + co0 + (lane >> 1);
        d[0] = hi; d[16] = mid;
      }
    }
    if (fwd && co0 + row < Cout) {                      // row = output channel, lanes = input channels
      const float v = tile[lane][row], nb = __shfl_xor_sync(0xffffffffu, v, 1);
      uint32_t hi, mid;
      split_bf16x2(v, nb, hi, mid);
      if (!(lane & 1)) {
        uint32_t* d = fwd + ((size_t)tap * Cout + co0 + row) * cin_pad + ci0 + (lane >> 1);
        d[0] = hi; d[16] = mid;
      }
    }
  }
}
}  // namespace

extern "C" int sg2im_split_weights(const int64_t* table, int64_t n_entries, int64_t total_tiles,
                                   sg2im_stream_t stream) {
  SG_ARG(table && n_entries >= 1 && total_tiles >= 1 && total_tiles < (1ll << 31));
  SG_LAUNCH(split_weights_kernel, (unsigned)total_tiles, 256, 0, as_stream(stream),
            reinterpret_cast<const SplitEntry*>(table), (int)n_entries);
  SG_LAUNCH_OK();
  return 0;
}
