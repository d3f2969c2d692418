// TEST INFRASTRUCTURE ONLY — a host-side SIMT emulation of the CUDA builtins the
// non-tensor-core kernels of sg2im_b200/csrc use, so that the KERNEL SOURCES THEMSELVES
// (compiled by g++ with -DSG2IM_EMUL) can be executed on the CPU and checked against the
// oracle without a GPU: one OS thread per CUDA thread of a block, blocks run one after
// another, __syncthreads / warp shuffles / ballots as real barriers, atomics under a lock.
// It checks index arithmetic, tiling, masking and reduction logic of the exact source that
// nvcc compiles; it does not model memory ordering, timing or the tensor-core / TMA paths.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

typedef void* cudaStream_t;
typedef int cudaError_t;

struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

namespace emul {

struct Block {
  unsigned nthreads = 0;
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint32_t> xch;            // one exchange slot per thread (shuffles / ballots)
  std::vector<unsigned char> dyn_smem;
};

inline Block*& current() { static Block* b = nullptr; return b; }
inline unsigned long long& blocks_run() { static unsigned long long n = 0; return n; }   // test introspection
inline std::mutex& atomic_lock() { static std::mutex m; return m; }

}  // namespace emul

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

// linear thread id inside the block (x fastest, like the hardware's warp packing)
static inline unsigned emul_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }
static inline unsigned emul_lane() { return emul_tid() & 31u; }
static inline unsigned emul_warp() { return emul_tid() >> 5; }

static inline void __syncthreads() { emul::current()->block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  emul::current()->warp_bar[emul_warp()]->arrive_and_wait();
}
static inline void* emul_dynamic_smem() { return emul::current()->dyn_smem.data(); }

template <class T>
static inline T emul_shfl_from(T v, unsigned src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  emul::Block* b = emul::current();
  uint32_t bits;
  std::memcpy(&bits, &v, 4);
  b->xch[emul_tid()] = bits;
  b->warp_bar[emul_warp()]->arrive_and_wait();
  uint32_t got = b->xch[(emul_tid() & ~31u) + (src_lane & 31u)];
  b->warp_bar[emul_warp()]->arrive_and_wait();      // slots free for the next exchange
  T out;
  std::memcpy(&out, &got, 4);
  return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emul_shfl_from(v, (unsigned)src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emul_shfl_from(v, emul_lane() ^ (unsigned)m); }
static inline unsigned __ballot_sync(unsigned, bool pred) {
  emul::Block* b = emul::current();
  b->xch[emul_tid()] = pred ? 1u : 0u;
  b->warp_bar[emul_warp()]->arrive_and_wait();
  unsigned m = 0;
  const unsigned lanes = std::min(32u, b->nthreads - (emul_tid() & ~31u));
  for (unsigned l = 0; l < lanes; ++l) m |= (b->xch[(emul_tid() & ~31u) + l] & 1u) << l;
  b->warp_bar[emul_warp()]->arrive_and_wait();
  return m;
}

template <class T> static inline T atomicAdd(T* p, T v) {
  std::lock_guard<std::mutex> g(emul::atomic_lock());
  T old = *p; *p = old + v; return old;
}
static inline uint32_t atomicMin(uint32_t* p, uint32_t v) {
  std::lock_guard<std::mutex> g(emul::atomic_lock());
  uint32_t old = *p; *p = std::min(old, v); return old;
}
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) {
  std::lock_guard<std::mutex> g(emul::atomic_lock());
  uint32_t old = *p; *p = std::max(old, v); return old;
}

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline float __fadd_rn(float a, float b) { return a + b; }       // no contraction on the host build
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
using std::min;
using std::max;

// Run `body` (a call of the kernel with its arguments bound) for every thread of every block.
// Threads that return early simply drop out of the block barrier, like exited CUDA threads.
static inline void emul_launch(dim3 grid, dim3 block, size_t dyn_smem_bytes,
                               const std::function<void()>& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  emul::Block blk;
  blk.nthreads = nthreads;
  blk.xch.assign(nthreads, 0);
  blk.dyn_smem.assign(dyn_smem_bytes + 64, 0);
  emul::current() = &blk;
  const unsigned nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        ++emul::blocks_run();
        blk.block_bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)nthreads);
        blk.warp_bar.clear();
        for (unsigned w = 0; w < nwarps; ++w) {
          unsigned lanes = std::min(32u, nthreads - w * 32);
          blk.warp_bar.push_back(std::make_unique<std::barrier<>>((std::ptrdiff_t)lanes));
        }
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; ++t)
          th.emplace_back([&, t]() {
            threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            blockIdx = uint3{bx, by, bz};
            blockDim = block;
            gridDim = grid;
            body();
            blk.block_bar->arrive_and_drop();
          });
        for (auto& x : th) x.join();
      }
  emul::current() = nullptr;
}
