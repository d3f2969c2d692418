// TEST INFRASTRUCTURE ONLY — a host-side SIMT emulation of the CUDA builtins the
// non-tensor-core kernels of sg2im_b200/csrc use, so that the KERNEL SOURCES THEMSELVES
// (compiled by g++ with -DSG2IM_EMUL) can be executed on the CPU and checked against the
// oracle without a GPU.  Two execution models behind the same builtins:
//   default               every CUDA thread of a block is a cooperative fiber (ucontext) of ONE OS
//                         thread; __syncthreads / warp shuffles / ballots yield to the next fiber
//                         until the barrier completes.  Fast and deterministic.
//   -DSG2IM_EMUL_THREADS  one OS thread per CUDA thread with real barriers and locked atomics:
//                         true concurrency, for the sanitizer runs (tools/emul_sanitize.sh; TSan
//                         then sees missing-barrier races, ASan sees out-of-bounds accesses).
// Blocks run one after another in both models.
// It checks index arithmetic, tiling, masking and reduction logic of the exact source that
// nvcc compiles; it does not model memory ordering, timing or the tensor-core / TMA paths.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#ifndef SG2IM_EMUL_THREADS
#include <csetjmp>
#include <ucontext.h>
#endif

typedef void* cudaStream_t;
typedef int cudaError_t;

struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
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
inline unsigned long long& blocks_run() { static unsigned long long n = 0; return n; }   // test introspection
inline unsigned long long& cluster_blocks_run() { static unsigned long long n = 0; return n; }
inline std::mutex& atomic_lock() { static std::mutex m; return m; }
}  // namespace emul

#ifdef SG2IM_EMUL_THREADS
inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;
#else
inline uint3 threadIdx, blockIdx;                       // one OS thread: plain globals, set per fiber
inline dim3 blockDim, gridDim;
#endif

// linear thread id inside the block (x fastest, like the hardware's warp packing)
static inline unsigned emul_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }
static inline unsigned emul_lane() { return emul_tid() & 31u; }
static inline unsigned emul_warp() { return emul_tid() >> 5; }

#ifdef SG2IM_EMUL_THREADS
// ------------------------------------------------------------------ OS-thread model
namespace emul {
struct Block {
  unsigned nthreads = 0;
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint32_t> xch;            // one exchange slot per thread (shuffles / ballots)
  std::vector<unsigned char> dyn_smem;
};
inline Block*& current() { static Block* b = nullptr; return b; }
inline void block_sync() { current()->block_bar->arrive_and_wait(); }
inline void warp_sync() { current()->warp_bar[emul_warp()]->arrive_and_wait(); }
}  // namespace emul
#else
// ------------------------------------------------------------------ fiber model
#include <unordered_map>
namespace emul {
struct MBar {                          // functional mbarrier (tests/emul/tc_emul.h)
  uint32_t init_count = 0, pending = 0, phase = 0;
  long long tx = 0;
  bool inited = false;
};
struct Fiber {
  ucontext_t ctx;                      // first entry only (makecontext); afterwards jmp_buf switches,
  jmp_buf jb;                          // which skip swapcontext's signal-mask system call
  std::vector<char> stack;
  uint3 tid;
  unsigned block = 0;                  // index of its CTA inside the gang
  bool started = false, done = false;
};
struct Block {                         // one CTA
  unsigned nthreads = 0, rank = 0;
  uint3 bid;
  unsigned alive = 0, arrived = 0;
  unsigned long long gen = 0;
  std::vector<unsigned> w_alive, w_arrived;
  std::vector<unsigned long long> w_gen;
  std::vector<uint32_t> xch;
  std::vector<unsigned char> smem_store;
  unsigned char* smem = nullptr;       // 1024-byte aligned
  size_t smem_bytes = 0;
  std::unordered_map<uint32_t, MBar> mbar;   // keyed by shared-memory offset
  std::vector<float> tmem;             // [128 lanes][512 columns], allocated on first use
};
struct Gang {                          // the CTAs that run concurrently: one block, or one cluster
  std::vector<Block> blocks;
  std::vector<Fiber> fibers;
  ucontext_t main_ctx;
  jmp_buf main_jb;
  unsigned cur = 0;
  unsigned c_alive = 0, c_arrived = 0;
  unsigned long long c_gen = 0;
  const std::function<void()>* body = nullptr;
};
inline Gang*& gang() { static Gang* g = nullptr; return g; }
// optional hooks of a device model layered on top (tc_emul.h: asynchronous TMA / tensor pipe):
// called once per scheduler sweep over the gang's fibers, and when the gang has finished
inline void (*&sweep_hook())() { static void (*h)() = nullptr; return h; }
inline void (*&drain_hook())() { static void (*h)() = nullptr; return h; }
inline Block* current() { Gang* g = gang(); return &g->blocks[g->fibers[g->cur].block]; }
inline void yield() {
  Gang* g = gang();
  if (!_setjmp(g->fibers[g->cur].jb)) _longjmp(g->main_jb, 1);
}
inline void block_sync() {
  Block* b = current();
  const unsigned long long g = b->gen;
  if (++b->arrived == b->alive) { b->arrived = 0; ++b->gen; return; }
  while (b->gen == g) yield();
}
inline void warp_sync() {
  Block* b = current();
  const unsigned w = emul_warp();
  const unsigned long long g = b->w_gen[w];
  if (++b->w_arrived[w] == b->w_alive[w]) { b->w_arrived[w] = 0; ++b->w_gen[w]; return; }
  while (b->w_gen[w] == g) yield();
}
inline void cluster_sync() {
  Gang* G = gang();
  const unsigned long long g = G->c_gen;
  if (++G->c_arrived == G->c_alive) { G->c_arrived = 0; ++G->c_gen; return; }
  while (G->c_gen == g) yield();
}
inline void fiber_exit_bookkeeping() {
  // an exited thread no longer takes part in barriers; complete any barrier it was the last missing of
  Gang* G = gang();
  Block* b = current();
  const unsigned w = (G->cur % b->nthreads) >> 5;
  if (--b->alive > 0 && b->arrived == b->alive) { b->arrived = 0; ++b->gen; }
  if (--b->w_alive[w] > 0 && b->w_arrived[w] == b->w_alive[w]) { b->w_arrived[w] = 0; ++b->w_gen[w]; }
  if (--G->c_alive > 0 && G->c_arrived == G->c_alive) { G->c_arrived = 0; ++G->c_gen; }
}
inline void trampoline() {
  // a fiber lives for the whole process: it runs one CUDA thread of block after block, launch
  // after launch (no per-block context creation, hence no per-block system calls)
  for (;;) {
    Gang* G = gang();
    (*G->body)();
    G = gang();
    Fiber& me = G->fibers[G->cur];
    me.done = true;
    fiber_exit_bookkeeping();
    if (!_setjmp(me.jb)) _longjmp(G->main_jb, 1);      // parked until the next block
  }
}
}  // namespace emul
#endif

static inline void __syncthreads() { emul::block_sync(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::warp_sync(); }
#ifdef SG2IM_EMUL_THREADS
static inline void* emul_dynamic_smem() { return emul::current()->dyn_smem.data(); }
#else
static inline void* emul_dynamic_smem() { return emul::current()->smem; }
#endif

template <class T>
static inline T emul_shfl_from(T v, unsigned src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  emul::Block* b = emul::current();
  uint32_t bits;
  std::memcpy(&bits, &v, 4);
  b->xch[emul_tid()] = bits;
  emul::warp_sync();
  uint32_t got = b->xch[(emul_tid() & ~31u) + (src_lane & 31u)];
  emul::warp_sync();                                   // slots free for the next exchange
  T out;
  std::memcpy(&out, &got, 4);
  return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emul_shfl_from(v, (unsigned)src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emul_shfl_from(v, emul_lane() ^ (unsigned)m); }
static inline unsigned __ballot_sync(unsigned, bool pred) {
  emul::Block* b = emul::current();
  b->xch[emul_tid()] = pred ? 1u : 0u;
  emul::warp_sync();
  unsigned m = 0;
  const unsigned lanes = std::min(32u, b->nthreads - (emul_tid() & ~31u));
  for (unsigned l = 0; l < lanes; ++l) m |= (b->xch[(emul_tid() & ~31u) + l] & 1u) << l;
  emul::warp_sync();
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
// Threads that return early stop taking part in that block's barriers, like exited CUDA threads.
#ifdef SG2IM_EMUL_THREADS
static inline void emul_launch(dim3 grid, dim3 block, size_t dyn_smem_bytes,
                               const std::function<void()>& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  if (nthreads == 0 || grid.x == 0 || grid.y == 0 || grid.z == 0) return;
  emul::Block blk;
  blk.nthreads = nthreads;
  blk.xch.assign(nthreads, 0);
  blk.dyn_smem.assign(dyn_smem_bytes + 64, 0);
  emul::current() = &blk;
  const unsigned nwarps = (nthreads + 31) / 32;
  const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
  auto arm = [&]() {                                   // fresh barriers for the next block
    blk.block_bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)nthreads);
    blk.warp_bar.clear();
    for (unsigned w = 0; w < nwarps; ++w) {
      unsigned lanes = std::min(32u, nthreads - w * 32);
      blk.warp_bar.push_back(std::make_unique<std::barrier<>>((std::ptrdiff_t)lanes));
    }
  };
  arm();
  std::barrier<> turn((std::ptrdiff_t)nthreads);       // all threads, between blocks
  std::vector<std::thread> th;
  th.reserve(nthreads);
  for (unsigned t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() {
      threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
      blockDim = block;
      gridDim = grid;
      for (unsigned long long b = 0; b < nblocks; ++b) {
        blockIdx = uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y),
                         (unsigned)(b / ((unsigned long long)grid.x * grid.y))};
        body();
        blk.block_bar->arrive_and_drop();              // done with this block's barrier
        turn.arrive_and_wait();                        // everyone finished block b
        if (t == 0) { ++emul::blocks_run(); if (b + 1 < nblocks) arm(); }
        turn.arrive_and_wait();                        // barriers re-armed
      }
    });
  for (auto& x : th) x.join();
  emul::current() = nullptr;
}
#else
// `cluster` consecutive blocks (along x) run concurrently as one gang; cluster == 1 is a plain launch.
static inline void emul_launch_cluster(unsigned cluster, dim3 grid, dim3 block, size_t dyn_smem_bytes,
                                       const std::function<void()>& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  if (nthreads == 0 || grid.x == 0 || grid.y == 0 || grid.z == 0) return;
  static emul::Gang G;                                 // fiber stacks are reused across launches
  constexpr size_t STACK = 256 * 1024;
  const unsigned nfib = cluster * nthreads;
  if (G.fibers.empty()) G.fibers.resize(4096);         // never reallocated: live fibers hold pointers into it
  if (nfib > G.fibers.size()) { std::fprintf(stderr, "emul: more than %zu threads per gang\n", G.fibers.size()); std::abort(); }
  for (unsigned f = 0; f < nfib; ++f)
    if (G.fibers[f].stack.size() != STACK) G.fibers[f].stack.resize(STACK);
  G.blocks.resize(cluster);
  G.body = &body;
  const unsigned nwarps = (nthreads + 31) / 32;
  emul::gang() = &G;
  blockDim = block;
  gridDim = grid;
  const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
  for (unsigned long long b0 = 0; b0 < nblocks; b0 += cluster) {
    G.c_alive = nfib; G.c_arrived = 0;
    for (unsigned r = 0; r < cluster; ++r) {
      emul::Block& blk = G.blocks[r];
      const unsigned long long b = b0 + r;
      blk.nthreads = nthreads; blk.rank = r;
      blk.bid = uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y),
                      (unsigned)(b / ((unsigned long long)grid.x * grid.y))};
      blk.alive = nthreads; blk.arrived = 0;
      blk.w_alive.assign(nwarps, 0); blk.w_arrived.assign(nwarps, 0); blk.w_gen.assign(nwarps, 0);
      blk.xch.assign(nthreads, 0);
      blk.smem_store.assign(dyn_smem_bytes + 2048, 0);
      blk.smem = reinterpret_cast<unsigned char*>(
          (reinterpret_cast<uintptr_t>(blk.smem_store.data()) + 1023) & ~(uintptr_t)1023);
      blk.smem_bytes = dyn_smem_bytes;
      blk.mbar.clear();
      for (unsigned t = 0; t < nthreads; ++t) {
        emul::Fiber& f = G.fibers[r * nthreads + t];
        f.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
        f.block = r;
        f.done = false;
        ++blk.w_alive[t >> 5];
        if (!f.started) {                                // created once per process
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack.data();
          f.ctx.uc_stack.ss_size = f.stack.size();
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, (void (*)())emul::trampoline, 0);
        }
      }
    }
    unsigned remaining = nfib;
    while (remaining) {
      if (emul::sweep_hook()) emul::sweep_hook()();
      for (unsigned f = 0; f < nfib; ++f) {
        emul::Fiber& fb = G.fibers[f];
        if (fb.done) continue;
        G.cur = f;
        threadIdx = fb.tid;
        blockIdx = G.blocks[fb.block].bid;
        if (!_setjmp(G.main_jb)) {
          if (fb.started) _longjmp(fb.jb, 1);
          fb.started = true;
          setcontext(&fb.ctx);                           // first entry: onto the fiber's own stack
        }
        if (fb.done) --remaining;
      }
    }
    if (emul::drain_hook()) emul::drain_hook()();
    emul::blocks_run() += cluster;
    if (cluster > 1) emul::cluster_blocks_run() += cluster;
  }
  emul::gang() = nullptr;
}
static inline void emul_launch(dim3 grid, dim3 block, size_t dyn_smem_bytes,
                               const std::function<void()>& body) {
  emul_launch_cluster(1, grid, block, dyn_smem_bytes, body);
}
#endif
