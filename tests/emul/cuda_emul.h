// cuda_emul.h -- a small CUDA execution-model emulator for CPU tests (TEST INFRASTRUCTURE).
//
// Compiles kernels written for nvcc as plain C++ and runs them with CUDA semantics: every thread of a
// block is a cooperative fiber (ucontext) on ONE OS thread, blocks run one after another in blockIdx
// order.  __syncthreads, the *_sync warp collectives (ballot, shfl, match.any) and atomics behave as on
// the device as long as the code is convergent where CUDA requires it.  A kernel that waits on a flag
// published by an EARLIER block (decoupled look-back, tickets taken in launch order) works; waiting on a
// later block would spin forever, as it may on hardware.
//
// Usage:   #include "cuda_emul.h"   (before the kernel source)
//          cuemu::launch(grid, block, [&] { my_kernel(args...); });
#pragma once

#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static   // one block runs at a time, so function-local statics are the block's shared memory

struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace cuemu {

struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
    bool at_block_barrier = false;
};

struct WarpCtx {
    unsigned arrived = 0;       // lanes that reached the current collective
    unsigned gen = 0;
    unsigned long long val[32];
};
// one context per member mask: lanes outside a partial-mask collective may meanwhile wait in another one
struct WarpSet {
    std::map<unsigned, WarpCtx> by_mask;
};

struct State {
    std::vector<Fiber> fibers;
    std::vector<WarpSet> warps;
    ucontext_t main_ctx;
    int cur = 0, n_threads = 0, n_done = 0;
    unsigned bar_arrived = 0, bar_gen = 0;
    std::function<void()> body;
};

inline State& st()
{
    static State s;
    return s;
}

}  // namespace cuemu

inline dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace cuemu {

inline void switch_to(int next)
{
    State& s = st();
    const int prev = s.cur;
    if (next == prev) return;
    s.cur = next;
    threadIdx.x = (unsigned)next;
    swapcontext(&s.fibers[prev].ctx, &s.fibers[next].ctx);
}

// give the other fibers of the block a turn
inline void yield()
{
    State& s = st();
    int n = s.cur;
    for (int i = 0; i < s.n_threads; ++i) {
        n = (n + 1) % s.n_threads;
        if (!s.fibers[n].done) break;
    }
    switch_to(n);
}

inline void trampoline()
{
    State& s = st();
    s.body();
    s.fibers[s.cur].done = true;
    ++s.n_done;
    // a block barrier counts exited threads as arrived
    if (s.n_done == s.n_threads) {
        setcontext(&s.main_ctx);
    }
    for (;;) yield();   // never scheduled again once done (yield skips finished fibers)
}

template <typename F>
void launch(dim3 grid, dim3 block, F&& kernel)
{
    State& s = st();
    gridDim = grid;
    blockDim = block;
    const int nt = (int)block.x;
    if (nt % 32) {
        fprintf(stderr, "cuemu: block size must be a multiple of 32\n");
        abort();
    }
    s.body = kernel;
    for (unsigned b = 0; b < grid.x; ++b) {
        blockIdx = dim3(b);
        s.n_threads = nt;
        s.n_done = 0;
        s.bar_arrived = 0;
        s.fibers.resize(nt);
        s.warps.assign(nt / 32, WarpSet());
        for (int t = 0; t < nt; ++t) {
            Fiber& f = s.fibers[t];
            f.done = false;
            if (f.stack.empty()) f.stack.resize(256 << 10);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.data();
            f.ctx.uc_stack.ss_size = f.stack.size();
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        s.cur = 0;
        threadIdx = dim3(0);
        swapcontext(&s.main_ctx, &s.fibers[0].ctx);
    }
}

inline void block_barrier()
{
    State& s = st();
    const unsigned gen = s.bar_gen;
    ++s.bar_arrived;
    for (;;) {
        if (s.bar_gen != gen) return;
        if (s.bar_arrived + (unsigned)s.n_done >= (unsigned)s.n_threads) {
            s.bar_arrived = 0;
            ++s.bar_gen;
            return;
        }
        yield();
    }
}

// The lanes of `mask` of the calling warp exchange a 64-bit value; returns the context with val[] filled.
// Two phases so that a lane racing ahead into the next collective cannot overwrite values still being read.
inline WarpCtx& warp_exchange(unsigned mask, unsigned long long v)
{
    State& s = st();
    WarpCtx& w = s.warps[threadIdx.x >> 5].by_mask[mask];
    const unsigned lane = threadIdx.x & 31;
    const unsigned expect = (unsigned)__builtin_popcount(mask);
    if (!((mask >> lane) & 1u)) {
        fprintf(stderr, "cuemu: lane %u calls a collective whose mask %08x excludes it\n", lane, mask);
        abort();
    }
    // phase 1: publish
    unsigned gen = w.gen;
    w.val[lane] = v;
    if (++w.arrived == expect) {
        w.arrived = 0;
        ++w.gen;
    } else {
        while (w.gen == gen) yield();
    }
    return w;
}
inline void warp_release(unsigned mask)
{
    State& s = st();
    WarpCtx& w = s.warps[threadIdx.x >> 5].by_mask[mask];
    const unsigned expect = (unsigned)__builtin_popcount(mask);
    const unsigned gen = w.gen;
    if (++w.arrived == expect) {
        w.arrived = 0;
        ++w.gen;
    } else {
        while (w.gen == gen) yield();
    }
}

}  // namespace cuemu

inline void __syncthreads() { cuemu::block_barrier(); }
inline void __syncwarp(unsigned mask = 0xffffffffu)
{
    cuemu::warp_exchange(mask, 0);
    cuemu::warp_release(mask);
}
inline int __syncthreads_or(int pred)
{
    static int acc;
    if (threadIdx.x == 0) acc = 0;
    cuemu::block_barrier();
    if (pred) acc = 1;
    cuemu::block_barrier();
    const int r = acc;
    cuemu::block_barrier();
    return r;
}

inline unsigned __ballot_sync(unsigned mask, int pred)
{
    cuemu::WarpCtx& w = cuemu::warp_exchange(mask, pred ? 1 : 0);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) m |= (w.val[i] ? 1u : 0u) << i;
    cuemu::warp_release(mask);
    return m;
}
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src)
{
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    cuemu::WarpCtx& w = cuemu::warp_exchange(mask, raw);
    const unsigned long long r = w.val[src & 31];
    cuemu::warp_release(mask);
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
template <typename T>
inline T __shfl_up_sync(unsigned m, T v, unsigned d)
{
    const int lane = threadIdx.x & 31;
    const int src = lane >= (int)d ? lane - (int)d : lane;
    return __shfl_sync(m, v, src);
}
template <typename T>
inline T __shfl_down_sync(unsigned m, T v, unsigned d)
{
    const int lane = threadIdx.x & 31;
    const int src = lane + (int)d < 32 ? lane + (int)d : lane;
    return __shfl_sync(m, v, src);
}
template <typename T>
inline T __shfl_xor_sync(unsigned m, T v, int x)
{
    const int lane = threadIdx.x & 31;
    return __shfl_sync(m, v, lane ^ x);
}
template <typename T>
inline unsigned __match_any_sync(unsigned mask, T v)
{
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    cuemu::WarpCtx& w = cuemu::warp_exchange(mask, raw);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) m |= (w.val[i] == raw ? 1u : 0u) << i;
    cuemu::warp_release(mask);
    return m;
}

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
inline int __clzll(unsigned long long x) { return x ? __builtin_clzll(x) : 64; }

// single OS thread: plain read-modify-write is atomic
template <typename T, typename U>
inline T atomicAdd(T* p, U v) { const T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename U>
inline T atomicOr(T* p, U v) { const T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename U>
inline T atomicMin(T* p, U v) { const T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename U>
inline T atomicMax(T* p, U v) { const T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename U>
inline T atomicExch(T* p, U v) { const T o = *p; *p = (T)v; return o; }
template <typename T, typename U, typename V>
inline T atomicCAS(T* p, U cmp, V v) { const T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T, typename U> inline void __stcs(T* p, U v) { *p = (T)v; }

inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
inline double __dsqrt_rn(double a) { return __builtin_sqrt(a); }

using std::max;
using std::min;
