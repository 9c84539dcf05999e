// Thin device-runtime vocabulary used by the engine: memory, copies, streams, events, launches.
//
// Product build (nvcc, sm_100a): plain CUDA runtime calls, every status checked; any failure throws pd::Error,
// which the C ABI turns into an error code + message (there is no CPU path to fall back to).
// PD_EMU build (g++, tests/emu only): the same vocabulary mapped onto host memory and the fiber-based SIMT logic
// emulator in tests/emu/cuemu.h — test infrastructure for a GPU-less container, never shipped.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#ifdef PD_EMU
#include "cuemu.h"
#include <chrono>
#else
#include <cuda_runtime.h>
#include <chrono>

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>
#endif

namespace pd {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

namespace rt {

#ifndef PD_EMU

#define PD_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            throw pd::Error(-3, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" +  \
                                    std::to_string(__LINE__) + ")");                                       \
    } while (0)

typedef cudaStream_t stream_t;
typedef cudaEvent_t event_t;

inline int device_count() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
inline void set_device(int d) { PD_CUDA(cudaSetDevice(d)); }
inline int current_device() { int d = 0; PD_CUDA(cudaGetDevice(&d)); return d; }
inline int sm_count() {
    int d = current_device(), n = 0;
    PD_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d));
    return n;
}
inline size_t max_optin_smem() {
    int d = current_device(), n = 0;
    PD_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, d));
    return static_cast<size_t>(n);
}
// Device memory comes from a small caching allocator: an index build allocates ~20 multi-gigabyte buffers, and
// cudaMalloc / cudaFree of those cost milliseconds each (cudaFree also synchronises the device).  dfree() only
// queues a block; collect() — called where the engine has synchronised anyway — waits for the device and makes the
// queued blocks reusable by later dmalloc() calls of a similar size.  Blocks are cached per device.
struct BlockCache {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;                      // size -> block
    std::unordered_map<void*, std::pair<size_t, int>> live;        // block -> (size, device), handed out or queued
    std::vector<void*> pending;                                    // freed, possibly still in use by queued kernels
    size_t cached = 0;
    size_t keep_limit = 0;                                         // idle bytes kept before trimming (0 = not asked yet)
    uint64_t misses = 0, miss_bytes = 0, miss_ns = 0, hits = 0;    // dmalloc() calls that went to cudaMalloc / were served from the cache
};
inline BlockCache& block_cache(int dev) {
    static BlockCache caches[16];
    return caches[dev & 15];
}
inline size_t round_block(size_t n) {
    if (n == 0) n = 1;
    const size_t g = n >= (size_t(8) << 20) ? (size_t(2) << 20) : 512;
    return (n + g - 1) / g * g;
}
inline void trim_cache(BlockCache& c, size_t keep) {  // c.mu held
    while (c.cached > keep && !c.free_blocks.empty()) {
        auto it = std::prev(c.free_blocks.end());
        c.cached -= it->first;
        cudaFree(it->second);
        c.live.erase(it->second);
        c.free_blocks.erase(it);
    }
}
inline void collect() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    BlockCache& c = block_cache(dev);
    // only the blocks queued BEFORE the synchronisation are known to be idle after it: a block another thread frees
    // while this one waits may still be read by kernels queued after the wait began
    std::vector<void*> idle;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        if (c.pending.empty()) return;
        idle.swap(c.pending);
    }
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(c.mu);
    for (void* p : idle) {
        const size_t sz = c.live[p].first;
        c.free_blocks.emplace(sz, p);
        c.cached += sz;
    }
    // Give memory back only when the idle blocks exceed three quarters of the device (asked once: cudaMemGetInfo and
    // cudaFree cost milliseconds to tens of milliseconds, and a trimmed block is the next build's cudaMalloc);
    // an allocation that fails empties the cache anyway (dmalloc).
    if (c.keep_limit == 0) {
        size_t free_b = 0, total_b = 0;
        c.keep_limit = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess ? total_b / 4 * 3 : ~size_t(0);
        if (const char* e = getenv("PD_CACHE_KEEP_MB")) c.keep_limit = std::max<size_t>(1, (size_t)atoll(e)) << 20;  // hosts that share the GPU
    }
    if (c.cached > c.keep_limit) trim_cache(c, c.keep_limit);
}
inline void* dmalloc(size_t n) {
    n = round_block(n);
    int dev = current_device();
    BlockCache& c = block_cache(dev);
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.free_blocks.lower_bound(n);
        if (it != c.free_blocks.end() && it->first <= n + n / 8 + (size_t(1) << 20)) {
            void* p = it->second;
            c.cached -= it->first;
            c.free_blocks.erase(it);
            c.hits++;
            return p;
        }
    }
    void* p = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {  // give everything cached back to the driver and try once more
        cudaGetLastError();
        collect();
        {
            std::lock_guard<std::mutex> lk(c.mu);
            trim_cache(c, 0);
        }
        e = cudaMalloc(&p, n);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw pd::Error(-5, std::string("cudaMalloc of ") + std::to_string(n) + " bytes: " + cudaGetErrorString(e));
    }
    std::lock_guard<std::mutex> lk(c.mu);
    c.live[p] = std::make_pair(n, dev);
    c.misses++;
    c.miss_bytes += n;
    c.miss_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    return p;
}
inline void dfree(void* p) {
    if (!p) return;
    for (int d = 0; d < 16; d++) {
        BlockCache& c = block_cache(d);
        std::lock_guard<std::mutex> lk(c.mu);
        if (c.live.count(p)) {
            c.pending.push_back(p);
            return;
        }
    }
    cudaFree(p);
}
// gives every idle cached block of every device (and the idle pinned host blocks) back to the driver
inline void trim_all();
// allocator statistics since the last call (PD_TRACE): cache hits, cudaMalloc calls, their bytes and host time
inline void cache_stats(int dev, uint64_t* hits, uint64_t* misses, uint64_t* miss_bytes, uint64_t* miss_ns) {
    BlockCache& c = block_cache(dev);
    std::lock_guard<std::mutex> lk(c.mu);
    *hits = c.hits; *misses = c.misses; *miss_bytes = c.miss_bytes; *miss_ns = c.miss_ns;
    c.hits = c.misses = c.miss_bytes = c.miss_ns = 0;
}
// pinned host blocks are cached the same way (cudaMallocHost of the per-call result buffers costs milliseconds)
struct HostCache {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;
    std::unordered_map<void*, size_t> live;
    size_t cached = 0;
};
inline HostCache& host_cache() {
    static HostCache c;
    return c;
}
inline void* hmalloc(size_t n) {
    n = round_block(n);
    HostCache& c = host_cache();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.free_blocks.lower_bound(n);
        if (it != c.free_blocks.end() && it->first <= n + n / 4 + (size_t(1) << 20)) {
            void* p = it->second;
            c.cached -= it->first;
            c.free_blocks.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    PD_CUDA(cudaMallocHost(&p, n));
    std::lock_guard<std::mutex> lk(c.mu);
    c.live[p] = n;
    return p;
}
inline void hfree(void* p) {
    if (!p) return;
    HostCache& c = host_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) {
        cudaFreeHost(p);
        return;
    }
    if (c.cached + it->second > (size_t(16) << 30)) {  // keep at most 16 GiB of pinned memory around
        cudaFreeHost(p);
        c.live.erase(it);
        return;
    }
    c.free_blocks.emplace(it->second, p);
    c.cached += it->second;
}
inline void trim_all() {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return;
    for (int d = 0; d < 16; d++) {
        BlockCache& c = block_cache(d);
        bool any;
        {
            std::lock_guard<std::mutex> lk(c.mu);
            any = !c.pending.empty() || !c.free_blocks.empty();
        }
        if (!any) continue;
        if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); continue; }
        collect();
        std::lock_guard<std::mutex> lk(c.mu);
        trim_cache(c, 0);
    }
    cudaSetDevice(cur);
    HostCache& h = host_cache();
    std::lock_guard<std::mutex> lk(h.mu);
    for (auto& kv : h.free_blocks) {
        cudaFreeHost(kv.second);
        h.live.erase(kv.second);
    }
    h.free_blocks.clear();
    h.cached = 0;
}
// Streams and events are pooled per device and never given back to the driver: creating / destroying them goes through the
// driver's resource manager, whose lock is shared by every process on the box — a monitoring poll (nvidia-smi, NVML) that
// holds it for 25 ms stalls every such call of every rank for as long (measured at 8 ranks: whole build steps +25 .. +120 ms).
// Kernel launches, async copies and stream waits do not take that lock.  A stream is idle when it is handed back (every
// caller synchronises first).
struct HandlePool {
    std::mutex mu;
    std::vector<cudaStream_t> streams[16];
    std::vector<cudaEvent_t> events[16];
    std::unordered_map<void*, int> home;   // handle -> the device it was made on (it goes back to that device's list)
};
inline HandlePool& handle_pool() {
    static HandlePool* p = new HandlePool;   // (never destroyed: the CUDA runtime may be gone at exit)
    return *p;
}
inline stream_t stream_create() {
    const int d = current_device() & 15;
    HandlePool& hp = handle_pool();
    {
        std::lock_guard<std::mutex> lk(hp.mu);
        if (!hp.streams[d].empty()) {
            stream_t s = hp.streams[d].back();
            hp.streams[d].pop_back();
            return s;
        }
    }
    stream_t s;
    PD_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    std::lock_guard<std::mutex> lk(hp.mu);
    hp.home[reinterpret_cast<void*>(s)] = d;
    return s;
}
inline void stream_destroy(stream_t s) {
    HandlePool& hp = handle_pool();
    std::lock_guard<std::mutex> lk(hp.mu);
    auto it = hp.home.find(reinterpret_cast<void*>(s));
    if (it != hp.home.end()) hp.streams[it->second].push_back(s);
}
inline void sync(stream_t s) { PD_CUDA(cudaStreamSynchronize(s)); }
// non-blocking: true when everything queued on the stream has finished; throws on a failed stream
inline bool stream_idle(stream_t s) {
    const cudaError_t e = cudaStreamQuery(s);
    if (e == cudaErrorNotReady) return false;
    PD_CUDA(e);
    return true;
}
inline void h2d(void* d, const void* h, size_t n, stream_t s) { if (n) PD_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
inline void d2h(void* h, const void* d, size_t n, stream_t s) { if (n) PD_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); }
inline void d2d(void* dst, const void* src, size_t n, stream_t s) { if (n) PD_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, s)); }
inline void zero(void* d, size_t n, stream_t s) { if (n) PD_CUDA(cudaMemsetAsync(d, 0, n, s)); }
inline void fill_byte(void* d, int v, size_t n, stream_t s) { if (n) PD_CUDA(cudaMemsetAsync(d, v, n, s)); }
inline event_t event_create() {
    const int d = current_device() & 15;
    HandlePool& hp = handle_pool();
    {
        std::lock_guard<std::mutex> lk(hp.mu);
        if (!hp.events[d].empty()) {
            event_t e = hp.events[d].back();
            hp.events[d].pop_back();
            return e;
        }
    }
    event_t e;
    PD_CUDA(cudaEventCreate(&e));
    std::lock_guard<std::mutex> lk(hp.mu);
    hp.home[reinterpret_cast<void*>(e)] = d;
    return e;
}
inline void event_destroy(event_t e) {
    HandlePool& hp = handle_pool();
    std::lock_guard<std::mutex> lk(hp.mu);
    auto it = hp.home.find(reinterpret_cast<void*>(e));
    if (it != hp.home.end()) hp.events[it->second].push_back(e);
}
inline void event_record(event_t e, stream_t s) { PD_CUDA(cudaEventRecord(e, s)); }
inline float event_ms(event_t a, event_t b) { float ms = 0; PD_CUDA(cudaEventSynchronize(b)); PD_CUDA(cudaEventElapsedTime(&ms, a, b)); return ms; }
inline void check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw pd::Error(-3, std::string("launch ") + what + ": " + cudaGetErrorString(e));
}
template <class K>
inline void allow_smem(K kern, size_t bytes) {
    if (bytes <= 48 * 1024) return;
    // once per (kernel, device): cudaFuncSetAttribute stalls for milliseconds when other host threads have work in flight
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;
    const std::pair<const void*, int> key(reinterpret_cast<const void*>(kern), current_device());
    std::lock_guard<std::mutex> lk(mu);
    size_t& have = done[key];
    if (have >= bytes) return;
    PD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    have = bytes;
}
template <class K>
inline int occupancy(K kern, int threads, size_t smem) {
    int n = 0;
    PD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem));
    return n;
}

#define PD_LAUNCH(kern, grid, block, smem, stream, ...)              \
    do {                                                             \
        kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);    \
        pd::rt::check_launch(#kern);                                 \
    } while (0)

#define PD_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]

#else  // ---------------------------------------------------------------- PD_EMU (tests only)

typedef int stream_t;
struct event_rec { std::chrono::steady_clock::time_point t; };
typedef event_rec* event_t;

inline int device_count() { return 1; }
inline void set_device(int) {}
inline int current_device() { return 0; }
inline int sm_count() { return 2; }  // tiny grid: the emulator runs blocks one after another
inline size_t max_optin_smem() { return 227 * 1024; }
inline void* dmalloc(size_t n) {
    void* p = malloc(n ? n : 1);
    if (p) memset(p, 0xA5, n);  // device memory is not zero-initialised
    return p;
}
inline void dfree(void* p) { free(p); }
inline void collect() {}
inline void trim_all() {}
inline void* hmalloc(size_t n) { return malloc(n ? n : 1); }
inline void hfree(void* p) { free(p); }
inline stream_t stream_create() { return 0; }
inline void stream_destroy(stream_t) {}
inline void sync(stream_t) {}
inline bool stream_idle(stream_t) { return true; }
inline void h2d(void* d, const void* h, size_t n, stream_t) { if (n) memcpy(d, h, n); }
inline void d2h(void* h, const void* d, size_t n, stream_t) { if (n) memcpy(h, d, n); }
inline void d2d(void* dst, const void* src, size_t n, stream_t) { if (n) memmove(dst, src, n); }
inline void zero(void* d, size_t n, stream_t) { if (n) memset(d, 0, n); }
inline void fill_byte(void* d, int v, size_t n, stream_t) { if (n) memset(d, v, n); }
inline event_t event_create() { return new event_rec; }
inline void event_destroy(event_t e) { delete e; }
inline void event_record(event_t e, stream_t) { e->t = std::chrono::steady_clock::now(); }
inline float event_ms(event_t a, event_t b) { return std::chrono::duration<float, std::milli>(b->t - a->t).count(); }
template <class K>
inline void allow_smem(K, size_t) {}
template <class K>
inline int occupancy(K, int, size_t) { return 1; }

template <class K, class... A>
inline void emu_launch(K kern, dim3 grid, dim3 block, size_t smem, A... args) {
    cuemu::launch(grid, block, smem, [=]() { kern(args...); });
}

#define PD_LAUNCH(kern, grid, block, smem, stream, ...)                              \
    do {                                                                             \
        (void)(stream);                                                              \
        pd::rt::emu_launch(kern, dim3(grid), dim3(block), (smem), __VA_ARGS__);      \
    } while (0)

#define PD_DYNAMIC_SMEM(name) unsigned char* name = cuemu::dyn_smem()

#endif

// RAII device buffer
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) { release(); p = static_cast<T*>(dmalloc(count * sizeof(T))); n = count; }
    void ensure(size_t count) { if (count > n) alloc(count); }
    // per-call buffers whose size follows the input: grow with headroom so that a run of slightly larger calls does
    // not reallocate (a reallocation costs a cudaMalloc and, through collect(), a device-wide synchronisation)
    void grow(size_t count) { if (count > n) alloc(count + count / 4); }
    void release() { if (p) dfree(p); p = nullptr; n = 0; }
    size_t bytes() const { return n * sizeof(T); }
};

template <class T>
struct PinBuf {
    T* p = nullptr;
    size_t n = 0;
    PinBuf() {}
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    ~PinBuf() { release(); }
    void ensure(size_t count) {
        if (count > n) { release(); p = static_cast<T*>(hmalloc(count * sizeof(T))); n = count; }
    }
    void grow(size_t count) { if (count > n) ensure(count + count / 4); }
    void release() { if (p) hfree(p); p = nullptr; n = 0; }
};

}  // namespace rt
}  // namespace pd
