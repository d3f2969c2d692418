// Cluster / TMA-multicast variant of the tcgen05 weight-gradient kernel
// (conv_wgrad_tc.cu).  OPT-IN (SG2IM_WGRAD_MC=1): written after round 1's GPU
// budget was spent, assembled with ptxas (UTMALDG.4D.MULTICAST, UTCBAR.MULTICAST)
// but NOT yet run on hardware.
//
// Why (profiles/r01_prof_conv_wgrad_tc_v2.txt): TMEM holds 128 ci x
// 512 accumulator columns, so a 3x3 filter needs several (co tile, tap pass) work
// items per ci tile, and each of them re-streams the same X halo tiles (30 KB per
// 32 pixels) from L2: 3.07 GB of L2->SM traffic at 6.8 TB/s for the 288->64 layer,
// tensor pipe 42 %.  Here the CS (2 or 4) work items that share a ci tile and a
// pixel range run as ONE thread-block cluster in lockstep: CTA rank 0 fetches
// every X halo tile once and TMA-multicasts it into the same smem slot of all CS
// CTAs; each CTA fetches only its own dY tile.  L2->SM bytes per CTA and stage
// drop from 30 + b to 30/CS + b KB (b = 8 KB at BN = 64).
//
// Protocol per pipeline stage s (all barriers at identical smem offsets in every CTA):
//   full[s]    count 1, tx = X bytes (multicast, issued by rank 0) + own dY bytes;
//              each CTA's producer thread arms its own barrier, rank 0 also issues X.
//   empty_b[s] count 1, local: the CTA's MMAs have released the dY half -> own producer.
//   empty_a[s] count CS, lives in every CTA but only rank 0's copy is used: every
//              CTA's MMA warp commits to it with a multicast arrive (mask = rank 0);
//              rank 0's producer waits for all CS before overwriting the X slot.
// Cluster barrier after mbarrier init (no remote arrive may hit an uninitialised
// barrier) and before exit (no CTA may retire while a peer can still signal it).
#include <cstdlib>
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int MC_THREADS = 192;
constexpr int MC_A_ATOM = 8192;             // halo tile of one 32-channel atom, padded to 1 KB
constexpr int MC_B_ATOM = 4096;             // 32 pixel rows x 128 B
constexpr int MC_A_STAGE = 4 * MC_A_ATOM;   // M = 128 channels = 4 atoms

struct McParams {
  int Cin, Cout, KH, KW, P, taps;
  int RH, pitch;
  int tiles_w, tiles_h, total_ptiles;
  int ci_tiles, co_tiles, passes, T, splits, per_split;
  int a_bytes;
  int subgroups;                  // (co_tiles * passes) / CS
  float* dw;
};

template <int BN>
struct MCfg {
  static constexpr int B_STAGE = (BN / 32) * MC_B_ATOM;
  static constexpr int STAGE = MC_A_STAGE + B_STAGE;
  static constexpr int STAGES = BN == 256 ? 3 : (BN == 128 ? 4 : 5);
  static constexpr int SMEM_BYTES = STAGES * STAGE + 1024 + 256;
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                    ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};

template <int BN, int CS>
__global__ void __launch_bounds__(MC_THREADS, 1)
conv_wgrad_tc_mc_kernel(const __grid_constant__ CUtensorMap tmX,
                        const __grid_constant__ CUtensorMap tmDY, const McParams p) {
  using C = MCfg<BN>;
  SG_DYN_SMEM(uint8_t, smem_raw);
  // dynamic smem starts at the same offset in every CTA of the launch; the 1 KB round-up is
  // therefore identical too, which the multicast (same-offset) addressing relies on
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE);
  uint64_t* full = bars;                          // [STAGES]
  uint64_t* empty_b = bars + C::STAGES;           // [STAGES] local
  uint64_t* empty_a = bars + 2 * C::STAGES;       // [STAGES] used on rank 0
  uint64_t* tfull = bars + 3 * C::STAGES;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int cluster_id = blockIdx.x / CS, nclusters = gridDim.x / CS;
  const int total_items = p.ci_tiles * p.splits * p.subgroups;      // per cluster

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1); mbar_init(&empty_b[i], 1); mbar_init(&empty_a[i], CS);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    tc_alloc(tmem_slot, 512u);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // every CTA's barriers exist before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // cluster item -> (ci tile, pixel split, subgroup); this CTA's member = subgroup*CS + rank
  auto decode = [&](int item, int& ci0, int& co0, int& pass, int& t0, int& t1) {
    int sub = item % p.subgroups; item /= p.subgroups;
    int ci = item % p.ci_tiles;
    int split = item / p.ci_tiles;
    int member = sub * CS + (int)rank;
    int co = member % p.co_tiles;
    pass = member / p.co_tiles;
    ci0 = ci * 128; co0 = co * BN;
    t0 = split * p.per_split;
    t1 = t0 + p.per_split < p.total_ptiles ? t0 + p.per_split : p.total_ptiles;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      const uint16_t all_mask = (uint16_t)((1u << CS) - 1u);
      for (int item = cluster_id; item < total_items; item += nclusters) {
        int ci0, co0, pass, t0, t1;
        decode(item, ci0, co0, pass, t0, t1);
        int na = (p.Cin - ci0 + 31) / 32; if (na > 4) na = 4;
        int nb = (p.Cout - co0 + 31) / 32; if (nb > BN / 32) nb = BN / 32;
        const uint32_t bytes = (uint32_t)(na * p.a_bytes + nb * MC_B_ATOM);
        for (int pt = t0; pt < t1; ++pt) {
          int tw = pt % p.tiles_w;
          int r = pt / p.tiles_w;
          int th = r % p.tiles_h;
          int n = r / p.tiles_h;
          int x0 = tw * 8, y0 = th * p.RH;
          mbar_wait(&empty_b[s], ph ^ 1);                       // own MMAs released the slot
          if (rank == 0) mbar_wait(&empty_a[s], ph ^ 1);        // ... and every peer's did
          mbar_expect_tx(&full[s], bytes);
          uint8_t* sa = smem + s * C::STAGE;
          uint8_t* sb = sa + MC_A_STAGE;
          if (rank == 0)
            for (int a = 0; a < na; ++a)
              tma_load_4d_mc(sa + a * MC_A_ATOM, &tmX, &full[s], all_mask, ci0 + a * 32, x0 - p.P,
                             y0 - p.P, n);
          for (int b = 0; b < nb; ++b)
            tma_load_4d(sb + b * MC_B_ATOM, &tmDY, &full[s], co0 + b * 32, x0, y0, n);
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, elected lane issues) =====================
    const uint32_t leader = lane == 0 ? 1u : 0u;
    const uint32_t d_hi = 32u | (1u << 14) | (1u << 29);
    const uint32_t a_lo0 = (smem_u32(smem) >> 4) | ((uint32_t)(MC_A_ATOM >> 4) << 16);
    const uint32_t b_lo0 = ((smem_u32(smem) + MC_A_STAGE) >> 4) | ((uint32_t)(MC_B_ATOM >> 4) << 16);
    const uint32_t pitch16 = (uint32_t)p.pitch * 8u;
    const uint32_t row_wrap = (uint32_t)(p.pitch - p.KW) * 8u;
    int s = 0; uint32_t ph = 0;
    uint32_t acc_ph = 0;
    for (int item = cluster_id; item < total_items; item += nclusters) {
      int ci0, co0, pass, t0, t1;
      decode(item, ci0, co0, pass, t0, t1);
      const int tap0 = pass * p.T;
      const int ntap = (p.taps - tap0) < p.T ? (p.taps - tap0) : p.T;
      const int ky0 = tap0 / p.KW, kx0 = tap0 - ky0 * p.KW;
      const uint32_t tap_off0 = (uint32_t)(ky0 * p.pitch + kx0) * 8u;
      mbar_wait(tempty, acc_ph ^ 1);
      tc_fence_after();
      for (int pt = t0; pt < t1; ++pt) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t first = (pt > t0) ? 1u : 0u;
        uint32_t at = a_lo0 + (uint32_t)s * (C::STAGE >> 4) + tap_off0;
        const uint32_t bt = b_lo0 + (uint32_t)s * (C::STAGE >> 4);
        uint32_t d_tmem = tmem_base;
        int kx = kx0;
        for (int tl = 0; tl < ntap; ++tl) {
          tc_mma_tf32_lh(d_tmem, at, d_hi, bt, d_hi, C::IDESC, first, leader);
          tc_mma_tf32_lh(d_tmem, at + pitch16, d_hi, bt + 64, d_hi, C::IDESC, 1u, leader);
          tc_mma_tf32_lh(d_tmem, at + 2 * pitch16, d_hi, bt + 128, d_hi, C::IDESC, 1u, leader);
          tc_mma_tf32_lh(d_tmem, at + 3 * pitch16, d_hi, bt + 192, d_hi, C::IDESC, 1u, leader);
          d_tmem += BN;
          at += 8u;
          if (++kx == p.KW) { kx = 0; at += row_wrap; }
        }
        tc_commit(&empty_b[s], leader);                          // own dY half free
        tc_commit_mc(&empty_a[s], (uint16_t)1);                  // X half free: tell rank 0
        if (++s == C::STAGES) { s = 0; ph ^= 1; }
      }
      tc_commit(tfull, leader);
      acc_ph ^= 1;
    }
  } else {
    // ===================== epilogue: TMEM -> vector atomics into dW =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t acc_ph = 0;
    for (int item = cluster_id; item < total_items; item += nclusters) {
      int ci0, co0, pass, t0, t1;
      decode(item, ci0, co0, pass, t0, t1);
      mbar_wait(tfull, acc_ph);
      tc_fence_after();
      const int ci = ci0 + row;
      const bool valid = ci < p.Cin;
      for (int tl = 0; tl < p.T; ++tl) {
        int tap = pass * p.T + tl;
        if (tap >= p.taps) break;
        float* drow = p.dw + ((long long)tap * p.Cin + ci) * p.Cout + co0;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * BN);
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          if (co0 + ch * 32 >= p.Cout) break;
          float v[32];
          tc_ld32(taddr + ch * 32, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              atomicAdd(reinterpret_cast<float4*>(drow + ch * 32 + j),
                        make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      acc_ph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // peers may still signal this CTA's barriers until here
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, 512u);
  }
}

template <int BN, int CS>
int launch_mc(const CUtensorMap& tmX, const CUtensorMap& tmDY, const McParams& p, cudaStream_t st) {
  using C = MCfg<BN>;
  const int items = p.ci_tiles * p.splits * p.subgroups;
#ifdef SG2IM_EMUL
  int max_clusters = num_sms() / CS;
  int nclusters = items < max_clusters ? items : max_clusters;
  if (nclusters < 1) nclusters = 1;
  emul_launch_cluster(CS, dim3((unsigned)(nclusters * CS)), dim3(MC_THREADS), (size_t)C::SMEM_BYTES,
                      [=]() { conv_wgrad_tc_mc_kernel<BN, CS>(tmX, tmDY, p); });
  (void)st;
  return 0;
#else
  auto kern = conv_wgrad_tc_mc_kernel<BN, CS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      sg2im_set_error("conv_wgrad_tc_mc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(MC_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cfg.attrs = attr; cfg.numAttrs = 1;
  // as many co-resident clusters as the device takes (one CTA per SM by smem), capped by the work
  cfg.gridDim = dim3((unsigned)(num_sms() / CS * CS));
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
    (void)cudaGetLastError();
    max_clusters = num_sms() / CS / 2;                           // conservative guess
    if (max_clusters < 1) max_clusters = 1;
  }
  int nclusters = items < max_clusters ? items : max_clusters;
  cfg.gridDim = dim3((unsigned)(nclusters * CS));
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmX, tmDY, p);
  if (e != cudaSuccess) {
    sg2im_set_error("conv_wgrad_tc_mc: launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
#endif
}

}  // namespace

// Called by sg2im_conv_wgrad_tc when SG2IM_WGRAD_MC=1.  `f` = the geometry the
// single-CTA kernel computed: {Cin, Cout, KH, KW, P, taps, RH, pitch, tiles_w, tiles_h,
// total_ptiles, ci_tiles, co_tiles, passes, T, a_bytes}.  Returns 0 when launched, >0 a
// CUDA error, -1 when the shape does not form clusters (caller runs the plain kernel).
int sg2im_wgrad_mc_launch(const CUtensorMap* tmX, const CUtensorMap* tmDY, const int* f, float* dw,
                          int BN, cudaStream_t st) {
  McParams p;
  p.Cin = f[0]; p.Cout = f[1]; p.KH = f[2]; p.KW = f[3]; p.P = f[4]; p.taps = f[5];
  p.RH = f[6]; p.pitch = f[7]; p.tiles_w = f[8]; p.tiles_h = f[9]; p.total_ptiles = f[10];
  p.ci_tiles = f[11]; p.co_tiles = f[12]; p.passes = f[13]; p.T = f[14]; p.a_bytes = f[15];
  p.dw = dw;
  const int members = p.co_tiles * p.passes;        // work items sharing one X stream
  int CS = members % 4 == 0 ? 4 : (members % 2 == 0 ? 2 : 1);
  if (CS == 1) return -1;
  p.subgroups = members / CS;
  // pixel splits: ~3 waves of clusters, at least 8 pixel tiles each
  const long long clusters_hint = tc::num_sms() / CS;
  long long base = (long long)p.ci_tiles * p.subgroups;
  long long want = (3 * clusters_hint + base - 1) / base;
  long long max_split = p.total_ptiles / 8 > 0 ? p.total_ptiles / 8 : 1;
  if (want > max_split) want = max_split;
  if (want < 1) want = 1;
  p.per_split = (int)((p.total_ptiles + want - 1) / want);
  p.splits = (p.total_ptiles + p.per_split - 1) / p.per_split;
  if (BN == 64) return CS == 4 ? launch_mc<64, 4>(*tmX, *tmDY, p, st) : launch_mc<64, 2>(*tmX, *tmDY, p, st);
  if (BN == 128) return CS == 4 ? launch_mc<128, 4>(*tmX, *tmDY, p, st) : launch_mc<128, 2>(*tmX, *tmDY, p, st);
  return CS == 4 ? launch_mc<256, 4>(*tmX, *tmDY, p, st) : launch_mc<256, 2>(*tmX, *tmDY, p, st);
}
