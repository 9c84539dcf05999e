// TEST INFRASTRUCTURE — NOT PRODUCT CODE, never loaded by the pandelos_b200 package.
//
// A small SIMT *logic* emulator: compiles the repository's .cu kernel sources with g++ (-DPD_EMU) and runs
// every thread of a block as a cooperative fiber (ucontext) on one OS thread, one block after another.
// Barriers and full-mask warp intrinsics are rendezvous points between fibers, so shared-memory hazards,
// indexing and warp-cooperative code behave as on hardware for race-free kernels.  It exists because the build
// container has no GPU: it lets `-m "not gpu"` tests exercise kernel logic against the oracle on tiny inputs
// before a GPU slot is spent.  It says nothing about performance and is not a fallback: the product library
// (libpandelos_b200.so) is only ever built by nvcc for sm_100a and refuses to run without a CUDA device.
#pragma once

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

namespace cuemu {

struct Barrier {
    unsigned expected = 0, arrived = 0, gen = 0;
};

struct Warp {
    Barrier bar;
    unsigned long long xbuf[32];
};

struct Fiber {
    ucontext_t ctx;
    bool done = false;
    unsigned tid = 0;
};

struct State {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    std::vector<char> stacks;
    Barrier block_bar;
    unsigned cur = 0;
    unsigned nthreads = 0;
    std::function<void()> body;
    std::vector<unsigned char> dyn_smem;
};

State& st();
void yield();
void launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body);
void barrier_wait(Barrier& b);
inline unsigned char* dyn_smem() { return st().dyn_smem.data(); }
inline Warp& cur_warp() { return st().warps[st().cur / 32]; }
inline unsigned lane() { return st().cur & 31; }

template <class T>
inline unsigned long long to_bits(T v) {
    unsigned long long b = 0;
    static_assert(sizeof(T) <= 8, "shuffle payload too wide");
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T from_bits(unsigned long long b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

// all-lanes exchange: publish, rendezvous, read a full snapshot, rendezvous
inline void exchange(unsigned long long mine, unsigned long long out[32]) {
    Warp& w = cur_warp();
    w.xbuf[lane()] = mine;
    barrier_wait(w.bar);
    memcpy(out, w.xbuf, sizeof(w.xbuf));
    barrier_wait(w.bar);
}

}  // namespace cuemu

extern dim3 threadIdx, blockIdx, blockDim, gridDim;
static const int warpSize = 32;

inline void __syncthreads() { cuemu::barrier_wait(cuemu::st().block_bar); }
inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::barrier_wait(cuemu::cur_warp().bar); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}

// Only full-mask, convergent use is supported (what the kernels are written to).
template <class T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    unsigned long long all[32];
    cuemu::exchange(cuemu::to_bits(v), all);
    unsigned l = cuemu::lane();
    unsigned base = l & ~static_cast<unsigned>(width - 1);
    return cuemu::from_bits<T>(all[base + (static_cast<unsigned>(src) & static_cast<unsigned>(width - 1))]);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
    unsigned long long all[32];
    cuemu::exchange(cuemu::to_bits(v), all);
    unsigned l = cuemu::lane();
    unsigned in_seg = l & static_cast<unsigned>(width - 1);
    return in_seg >= delta ? cuemu::from_bits<T>(all[l - delta]) : v;
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
    unsigned long long all[32];
    cuemu::exchange(cuemu::to_bits(v), all);
    unsigned l = cuemu::lane();
    unsigned in_seg = l & static_cast<unsigned>(width - 1);
    return in_seg + delta < static_cast<unsigned>(width) ? cuemu::from_bits<T>(all[l + delta]) : v;
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    unsigned long long all[32];
    cuemu::exchange(cuemu::to_bits(v), all);
    (void)width;
    return cuemu::from_bits<T>(all[cuemu::lane() ^ static_cast<unsigned>(m)]);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned long long all[32];
    cuemu::exchange(pred ? 1ull : 0ull, all);
    unsigned n = std::min(32u, cuemu::st().nthreads - (cuemu::st().cur & ~31u));
    unsigned r = 0;
    for (unsigned i = 0; i < n; i++)
        if (all[i]) r |= 1u << i;  // exited lanes have their slot zeroed by the scheduler
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) {
    // inactive lanes (beyond blockDim) publish nothing; treat as true
    unsigned long long all[32];
    cuemu::exchange(pred ? 1ull : 0ull, all);
    unsigned n = std::min(32u, cuemu::st().nthreads - (cuemu::st().cur & ~31u));
    for (unsigned i = 0; i < n; i++)
        if (!all[i]) return 0;
    (void)m;
    return 1;
}
template <class T>
inline unsigned __match_any_sync(unsigned, T v) {
    unsigned long long all[32];
    cuemu::exchange(cuemu::to_bits(v), all);
    unsigned n = std::min(32u, cuemu::st().nthreads - (cuemu::st().cur & ~31u));
    unsigned r = 0;
    unsigned long long mine = cuemu::to_bits(v);
    for (unsigned i = 0; i < n; i++)
        if (all[i] == mine) r |= 1u << i;
    return r;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz(static_cast<unsigned>(v)) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll(static_cast<unsigned long long>(v)) : 64; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __umulhi(unsigned a, unsigned b) { return static_cast<unsigned>((static_cast<unsigned long long>(a) * b) >> 32); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __int2float_rn(int v) { return static_cast<float>(v); }
inline float __uint2float_rn(unsigned v) { return static_cast<float>(v); }
inline int __float_as_int(float f) { return cuemu::from_bits<int>(cuemu::to_bits(f)); }
inline float __int_as_float(int i) { return cuemu::from_bits<float>(cuemu::to_bits(i)); }
inline unsigned __float_as_uint(float f) { return cuemu::from_bits<unsigned>(cuemu::to_bits(f)); }
inline float __uint_as_float(unsigned u) { return cuemu::from_bits<float>(cuemu::to_bits(u)); }
template <class T>
inline T __ldg(const T* p) { return *p; }

// single OS thread => plain read-modify-write is atomic
template <class T>
inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned* p, int v) { unsigned o = *p; *p = o + static_cast<unsigned>(v); return o; }
template <class T>
inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T>
inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T>
inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T>
inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T>
inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

using std::max;
using std::min;
inline unsigned min(unsigned a, int b) { return a < static_cast<unsigned>(b) ? a : static_cast<unsigned>(b); }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
