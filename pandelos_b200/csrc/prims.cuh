// Device-wide primitives written for this engine: exclusive scan and a stable LSD radix sort of 64-bit keys.
//
// Both are plain multi-kernel algorithms (no inter-block spinning, no cooperative launch): the sort is
// "per-tile histogram -> one flat scan of the digit-major histogram -> stable scatter", 24 B of HBM traffic per
// key per 8-bit pass.  It replaces the reference's serial counting_sort_ext passes (ig/native/library.cpp:172-187,
// 270-278).  Because k-mer keys are produced in gene order, only the RANK bits need sorting: stability supplies the
// (rank, seq) total order the reference obtains with extra seq passes.
#pragma once

#include "pd_rt.h"

namespace pd {
namespace prims {

static const int kScanThreads = 512;
static const int kScanItems = 8;
static const int kScanTile = kScanThreads * kScanItems;  // 4096

static const int kSortThreads = 256;
static const int kSortItems = 16;
static const int kSortTile = kSortThreads * kSortItems;  // 4096 keys per block
static const int kSortWarps = kSortThreads / 32;
static const int kRadix = 256;
static const int kSortGroup = 8;  // keys ranked together

// ---------------------------------------------------------------------------------------------- block helpers

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, unsigned lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += o;
    }
    return v;
}

// Exclusive scan of one value per thread over a block of THREADS threads (multiple of 32, <= 1024).
// `scratch` holds 33 words of shared memory.  Returns the exclusive prefix; *total gets the block sum.
template <int THREADS>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* scratch, uint32_t* total) {
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < (unsigned)(THREADS / 32)) ? scratch[lane] : 0u;
        uint32_t wi = warp_incl_scan(w, lane);
        scratch[lane] = wi - w;
        if (lane == 31) scratch[32] = wi;
    }
    __syncthreads();
    uint32_t res = scratch[warp] + incl - v;
    *total = scratch[32];
    __syncthreads();
    return res;
}

// ---------------------------------------------------------------------------------------------- exclusive scan

// phase 1: per-tile sums
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                                                    uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t scratch[33];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        uint64_t idx = base + (uint64_t)i * kScanThreads + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    uint32_t tot;
    block_excl_scan<kScanThreads>(s, scratch, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// phase 3 (and the single-block case): scan a tile, add the tile's offset.  Thread-blocked layout (each thread owns
// kScanItems consecutive items) so the local scan is one serial pass + one block scan.
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* in, uint32_t* out, uint64_t n,
                                                                   const uint32_t* tile_offsets, uint32_t* total_out) {
    __shared__ uint32_t scratch[33];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        s += v[i];
    }
    uint32_t tot;
    uint32_t pre = block_excl_scan<kScanThreads>(s, scratch, &tot);
    pre += tile_offsets ? tile_offsets[blockIdx.x] : 0u;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) out[base + i] = pre;
        pre += v[i];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *total_out = pre;
}

// Exclusive scan of n uint32 (sum must fit uint32). in == out allowed.  *d_total (device, may be null) gets the sum.
// `tmp` must hold scan_tmp_words(n) uint32.
inline size_t scan_tmp_words(uint64_t n) {
    size_t w = 0;
    uint64_t t = (n + kScanTile - 1) / kScanTile;
    while (t > 1) {
        w += t;
        t = (t + kScanTile - 1) / kScanTile;
    }
    return w + 1;
}

inline void exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* tmp, uint32_t* d_total,
                               rt::stream_t st, uint64_t* launches = nullptr) {
    if (n == 0) {
        if (d_total) rt::zero(d_total, sizeof(uint32_t), st);
        return;
    }
    uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles == 1) {
        PD_LAUNCH(scan_apply_kernel, 1, kScanThreads, 0, st, in, out, n, (const uint32_t*)nullptr, d_total);
        if (launches) *launches += 1;
        return;
    }
    PD_LAUNCH(scan_reduce_kernel, (unsigned)tiles, kScanThreads, 0, st, in, n, tmp);
    if (launches) *launches += 1;
    exclusive_scan_u32(tmp, tmp, tiles, tmp + tiles, nullptr, st, launches);
    PD_LAUNCH(scan_apply_kernel, (unsigned)tiles, kScanThreads, 0, st, in, out, n, (const uint32_t*)tmp, d_total);
    if (launches) *launches += 1;
}

// ---------------------------------------------------------------------------------------------- radix sort

// hist[d * tiles + tile] = number of keys of this tile whose digit (bits [shift, shift+8)) is d
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift,
                                                                   uint32_t tiles, uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[4][kRadix];  // 4 copies to thin out same-address conflicts
    const unsigned tid = threadIdx.x;
    for (unsigned i = tid; i < 4 * kRadix; i += kSortThreads) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
    uint32_t* mine = h[(tid >> 5) & 3];
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        uint64_t idx = base + (uint64_t)i * kSortThreads + tid;
        if (idx < n) atomicAdd(&mine[(unsigned)(keys[idx] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    if (tid < kRadix) hist[(uint64_t)tid * tiles + blockIdx.x] = h[0][tid] + h[1][tid] + h[2][tid] + h[3][tid];
}

// Stable scatter of one tile.  `offs` is the exclusive scan of the digit-major histogram, i.e.
// offs[d * tiles + tile] = first output slot of this tile's digit-d keys.
// Order inside the tile: warp w owns keys [w*512, (w+1)*512); round j covers 32 consecutive keys, one per lane.
__global__ void __launch_bounds__(kSortThreads, 4) radix_scatter_kernel(const uint64_t* __restrict__ keys, uint64_t* __restrict__ out,
                                                                      uint64_t n, int shift, uint32_t tiles,
                                                                      const uint32_t* __restrict__ offs) {
    PD_DYNAMIC_SMEM(smem_raw);
    uint64_t* stage = reinterpret_cast<uint64_t*>(smem_raw);                           // kSortTile keys
    uint32_t* wc = reinterpret_cast<uint32_t*>(smem_raw + sizeof(uint64_t) * kSortTile);  // [kSortWarps][kRadix]
    uint32_t* tile_excl = wc + kSortWarps * kRadix;                                       // [kRadix]
    uint32_t* gdelta = tile_excl + kRadix;                                                // [kRadix]
    uint32_t* scratch = gdelta + kRadix;                                                  // 33

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t tile_base = (uint64_t)blockIdx.x * kSortTile;
    const uint32_t tile_n = (uint32_t)((n - tile_base < (uint64_t)kSortTile) ? (n - tile_base) : (uint64_t)kSortTile);

    for (unsigned i = tid; i < kSortWarps * kRadix; i += kSortThreads) wc[i] = 0;
    __syncthreads();

    uint64_t key[kSortItems];
    uint16_t rnk[kSortItems];
    uint32_t* my_wc = wc + warp * kRadix;
    const unsigned lt_mask = (1u << lane) - 1u;
    // Ranking in groups of kSortGroup keys: first the group's peer masks (independent match instructions, all in
    // flight together), then the per-digit counter updates, which must stay in key order (stability)
#pragma unroll
    for (int j0 = 0; j0 < kSortItems; j0 += kSortGroup) {
        unsigned peers[kSortGroup];
        unsigned dig[kSortGroup];
#pragma unroll
        for (int jj = 0; jj < kSortGroup; jj++) {
            const int j = j0 + jj;
            const uint32_t local = warp * (32 * kSortItems) + j * 32 + lane;
            const bool valid = local < tile_n;
            key[j] = valid ? keys[tile_base + local] : 0ull;
            // invalid lanes take a private pseudo-digit so they never join a valid lane's peer group
            dig[jj] = valid ? ((unsigned)(key[j] >> shift) & 0xFFu) : (256u + lane);
            peers[jj] = __match_any_sync(0xffffffffu, dig[jj]);
        }
#pragma unroll
        for (int jj = 0; jj < kSortGroup; jj++) {
            const int j = j0 + jj;
            const unsigned leader = __ffs(peers[jj]) - 1;
            uint32_t base_cnt = 0;
            if (dig[jj] < 256u && lane == leader) {
                base_cnt = my_wc[dig[jj]];
                my_wc[dig[jj]] = base_cnt + __popc(peers[jj]);
            }
            base_cnt = __shfl_sync(0xffffffffu, base_cnt, leader);
            rnk[j] = (uint16_t)(base_cnt + __popc(peers[jj] & lt_mask));
            __syncwarp();
        }
    }
    __syncthreads();

    // per digit: exclusive scan over the warps (tile order), tile totals, then exclusive scan over digits
    uint32_t cnt = 0;
    if (tid < kRadix) {
        uint32_t run = 0;
        for (int w = 0; w < kSortWarps; w++) {
            uint32_t t = wc[w * kRadix + tid];
            wc[w * kRadix + tid] = run;
            run += t;
        }
        cnt = run;
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan<kSortThreads>(cnt, scratch, &tot);
    if (tid < kRadix) {
        tile_excl[tid] = ex;
        gdelta[tid] = offs[(uint64_t)tid * tiles + blockIdx.x] - ex;
    }
    __syncthreads();

#pragma unroll
    for (int j = 0; j < kSortItems; j++) {
        const uint32_t local = warp * (32 * kSortItems) + j * 32 + lane;
        if (local < tile_n) {
            const unsigned d = (unsigned)(key[j] >> shift) & 0xFFu;
            stage[tile_excl[d] + my_wc[d] + rnk[j]] = key[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        const uint32_t idx = i * kSortThreads + tid;
        if (idx < tile_n) {
            const uint64_t kv = stage[idx];
            const unsigned d = (unsigned)(kv >> shift) & 0xFFu;
            out[(uint64_t)gdelta[d] + idx] = kv;
        }
    }
}

inline size_t radix_scatter_smem() {
    return sizeof(uint64_t) * kSortTile + sizeof(uint32_t) * (kSortWarps * kRadix + 2 * kRadix + 33);
}
inline uint32_t radix_tiles(uint64_t n) { return (uint32_t)((n + kSortTile - 1) / kSortTile); }
// words of uint32 scratch needed by radix_sort_u64 for n keys
inline size_t radix_tmp_words(uint64_t n) {
    size_t h = (size_t)radix_tiles(n) * kRadix;
    return h + scan_tmp_words(h);
}

// Stable LSD sort of keys[0..n) by bits [bit_lo, bit_hi) in 8-bit digits.  Ping-pongs between `keys` and `alt`;
// returns the buffer that holds the result.  n < 2^32.
inline uint64_t* radix_sort_u64(uint64_t* keys, uint64_t* alt, uint64_t n, int bit_lo, int bit_hi, uint32_t* tmp,
                                rt::stream_t st, uint64_t* launches = nullptr) {
    if (n == 0) return keys;
    const uint32_t tiles = radix_tiles(n);
    const size_t hwords = (size_t)tiles * kRadix;
    rt::allow_smem(radix_scatter_kernel, radix_scatter_smem());
    uint64_t* src = keys;
    uint64_t* dst = alt;
    for (int shift = bit_lo; shift < bit_hi; shift += 8) {
        PD_LAUNCH(radix_hist_kernel, tiles, kSortThreads, 0, st, (const uint64_t*)src, n, shift, tiles, tmp);
        if (launches) *launches += 1;
        exclusive_scan_u32(tmp, tmp, hwords, tmp + hwords, nullptr, st, launches);
        PD_LAUNCH(radix_scatter_kernel, tiles, kSortThreads, radix_scatter_smem(), st, (const uint64_t*)src, dst, n, shift, tiles,
                  (const uint32_t*)tmp);
        if (launches) *launches += 1;
        uint64_t* t = src;
        src = dst;
        dst = t;
    }
    return src;
}

}  // namespace prims
}  // namespace pd
