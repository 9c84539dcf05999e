// K-mer key production + stable LSD radix sort in ONE SWEEP per digit (decoupled look-back), for the index build.
//
// Replaces, for the k-mer keys, the encode kernel + "histogram -> scan -> scatter" passes of prims.cuh: the reference's
// do_ranking / kmers.emplace_back (ig/native/library.cpp:134-150, 234-265) and its counting_sort_ext passes
// (library.cpp:172-187, 270-278).
//
//   tile_gene_kernel    gene that holds the first key of every 4096-key tile (so a tile finds its genes without a search
//                       over all S genes)
//   kmer_hist_kernel    reads the RESIDUES (1 B per k-mer, not the 8-B keys), makes every k-mer rank with the rolling
//                       update of library.cpp:75-79 and counts all digits of all passes at once
//   digit_bins_kernel   exclusive scan of each pass's 256 counts = first output slot of every digit
//   onesweep_kernel<1>  FIRST pass: makes the keys again from the residues (they are never written in gene order), ranks
//                       them by digit 0 and scatters them: 1 B read + 8 B written per key
//   onesweep_kernel<0>  later passes: 8 B read + 8 B written per key.  A tile publishes its digit counts, looks back over
//                       the tiles before it (aggregate / inclusive-prefix words, one per digit) and scatters: no separate
//                       histogram read of the keys, no scan kernels.
//
// Key = rank << seq_bits | gene.  Only the rank bits are sorted: keys are produced in gene order and every pass is stable,
// which yields the (rank, gene) order the reference gets from extra passes over the gene bytes.
// A rank slice [lo, hi) can be given: keys outside it are dropped by the first pass (rank-range sharding of the build
// over several GPUs); with lo = 0, hi = 2^63 every key is kept.
#pragma once

#include "index_kernels.cuh"
#include "pd_rt.h"
#include "prims.cuh"

namespace pd {
namespace sortk {

static const int kThreads = 256;
static const int kItems = 16;
static const int kTile = kThreads * kItems;  // 4096 keys per tile
static const int kWarps = kThreads / 32;
static const int kRadix = 256;
static const int kMaxPasses = 8;   // 63 rank bits at most
static const int kMaxSlices = 32;  // ranks of a sharded build
static const uint64_t kNoKey = ~0ull;  // never a key: rank_bits + seq_bits <= 63
static const uint32_t kInclusive = 0x80000000u;  // status word: 0 = not ready; bit 31 = inclusive prefix; else aggregate + 1

struct EncodeSrc {
    const uint8_t* res;
    const uint64_t* gene_off;   // S + 1
    const uint32_t* key_off;    // S + 1: exclusive scan of the k-mer counts
    const uint32_t* tile_gene;  // tiles + 1
    uint32_t S;
    int k;
    uint32_t base;
    uint64_t mult;   // base^(k-1)
    int seq_bits;
    uint64_t lo, hi;  // ranks kept: [lo, hi)
    uint64_t N;       // key positions
    uint64_t pos_lo, pos_hi;  // key positions kept: [pos_lo, pos_hi) (a rank's share of the genes in a sharded build)
    uint32_t n_slices;        // sharded build: the rank space is cut into n_slices slices at cuts[1 .. n_slices - 1]
    uint64_t cuts[kMaxSlices + 1];
    ik::ValTable vt;
};

// the slice (= destination rank of a sharded build) of a k-mer rank
__device__ __forceinline__ unsigned slice_of(const EncodeSrc& s, uint64_t r) {
    unsigned d = 0;
    for (uint32_t i = 1; i < s.n_slices; i++) d += r >= s.cuts[i] ? 1u : 0u;
    return d;
}

// tile_gene[t] = the gene whose k-mers include key t * kTile, for t < tiles; tile_gene[tiles] = S - 1
__global__ void __launch_bounds__(256) tile_gene_kernel(const uint32_t* __restrict__ key_off, uint32_t S, uint32_t tiles,
                                                         uint32_t* __restrict__ tile_gene) {
    const uint32_t g = blockIdx.x * 256u + threadIdx.x;
    if (g == 0) tile_gene[tiles] = S ? S - 1 : 0;
    if (g >= S) return;
    const uint32_t a = key_off[g], b = key_off[g + 1];
    if (a == b) return;
    for (uint32_t t = (a + kTile - 1) / kTile; (uint64_t)t * kTile < b; t++) tile_gene[t] = g;
}

// The keys [i0, i0 + n) (n <= kItems consecutive key positions) made by one thread: out(j, key) for j < kItems, kNoKey
// for j >= n and for ranks outside [lo, hi).  RankT = uint32_t when base^k < 2^32.  `val` is the alphabet table in
// shared memory; genes [g_lo, g_hi] are known to hold key i0.
template <typename RankT, typename Out>
__device__ __forceinline__ void encode_keys(const EncodeSrc& s, const uint8_t* val, uint64_t i0, uint32_t n, uint32_t g_lo, uint32_t g_hi, Out out) {
    if (n == 0) {
#pragma unroll
        for (int j = 0; j < kItems; j++) out(j, kNoKey);
        return;
    }
    uint32_t a = g_lo, b = g_hi;  // the last gene whose first key is <= i0 (a gene without k-mers never is: the next one starts at the same key)
    while (a < b) {
        const uint32_t m = a + (b - a + 1) / 2;
        if ((uint64_t)s.key_off[m] <= i0) a = m;
        else b = m - 1;
    }
    uint32_t g = a;
    uint64_t kend = s.key_off[g + 1];
    const uint8_t* p = s.res + s.gene_off[g] + (i0 - s.key_off[g]);
    const RankT base = (RankT)s.base, mult = (RankT)s.mult;
    RankT r = 0;
    bool fresh = true;
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        uint64_t key = kNoKey;
        if ((uint32_t)j < n) {
            const uint64_t i = i0 + j;
            if (i >= kend) {  // the next gene that has k-mers
                g++;
                while ((uint64_t)s.key_off[g + 1] <= i) g++;
                kend = s.key_off[g + 1];
                p = s.res + s.gene_off[g];
                fresh = true;
            }
            if (fresh) {
                r = 0;
                for (int x = 0; x < s.k; x++) r = r * base + (RankT)val[p[x]];
                fresh = false;
            } else {
                r = (r - (RankT)val[p[-1]] * mult) * base + (RankT)val[p[s.k - 1]];  // library.cpp:75-79
            }
            p++;
            if ((uint64_t)r >= s.lo && (uint64_t)r < s.hi && i >= s.pos_lo && i < s.pos_hi) key = ((uint64_t)r << s.seq_bits) | g;
        }
        out(j, key);
    }
}

// hist[p * 256 + d] += kept keys whose digit p is d, for all passes at once (SLICES: hist[d] += kept keys of slice d).
// Tiles [tile0, tile0 + tiles).
template <typename RankT, bool SLICES>
__global__ void __launch_bounds__(kThreads) kmer_hist_kernel(EncodeSrc s, uint32_t tile0, uint32_t tiles, int passes, uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[2][kMaxPasses * kRadix];
    __shared__ uint8_t val[256];
    const unsigned tid = threadIdx.x;
    for (unsigned i = tid; i < 2 * kMaxPasses * kRadix; i += kThreads) (&h[0][0])[i] = 0;
    val[tid] = s.vt.v[tid];
    __syncthreads();
    uint32_t* mine = h[(tid >> 5) & 1];
    for (uint32_t tile = tile0 + blockIdx.x; tile < tile0 + tiles; tile += gridDim.x) {
        const uint64_t i0 = (uint64_t)tile * kTile + (uint64_t)tid * kItems;
        const uint32_t n = i0 >= s.N ? 0u : (uint32_t)(s.N - i0 < (uint64_t)kItems ? s.N - i0 : (uint64_t)kItems);
        const int seq_bits = s.seq_bits;
        encode_keys<RankT>(s, val, i0, n, s.tile_gene[tile], s.tile_gene[tile + 1], [&](int, uint64_t key) {
            if (key != kNoKey) {
                const uint64_t r = key >> seq_bits;
                if (SLICES) atomicAdd(&mine[slice_of(s, r)], 1u);
                else
                    for (int p = 0; p < passes; p++) atomicAdd(&mine[p * kRadix + ((unsigned)(r >> (8 * p)) & 0xFFu)], 1u);
            }
        });
    }
    __syncthreads();
    for (unsigned i = tid; i < (unsigned)passes * kRadix; i += kThreads) {
        const uint32_t v = h[0][i] + h[1][i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// the same from keys that exist already (the keys a rank of a sharded build received): 16-byte loads
__global__ void __launch_bounds__(kThreads) keys_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift0, int passes,
                                                              uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[2][kMaxPasses * kRadix];
    const unsigned tid = threadIdx.x;
    for (unsigned i = tid; i < 2 * kMaxPasses * kRadix; i += kThreads) (&h[0][0])[i] = 0;
    __syncthreads();
    uint32_t* mine = h[(tid >> 5) & 1];
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + tid; i < n; i += stride) {
        const uint64_t r = keys[i] >> shift0;
        for (int p = 0; p < passes; p++) atomicAdd(&mine[p * kRadix + ((unsigned)(r >> (8 * p)) & 0xFFu)], 1u);
    }
    __syncthreads();
    for (unsigned i = tid; i < (unsigned)passes * kRadix; i += kThreads) {
        const uint32_t v = h[0][i] + h[1][i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// bins[p][d] = keys with a smaller digit p; totals[0] = number of keys
__global__ void __launch_bounds__(kRadix) digit_bins_kernel(const uint32_t* __restrict__ hist, int passes, uint32_t* __restrict__ bins,
                                                             uint32_t* __restrict__ totals) {
    __shared__ uint32_t scratch[33];
    for (int p = 0; p < passes; p++) {
        const uint32_t c = hist[p * kRadix + threadIdx.x];
        uint32_t tot;
        const uint32_t ex = prims::block_excl_scan<kRadix>(c, scratch, &tot);
        bins[p * kRadix + threadIdx.x] = ex;
        if (p == 0 && threadIdx.x == 0) totals[0] = tot;
    }
}

struct SweepArgs {
    const uint64_t* keys;  // MODE 0 input (n keys)
    uint64_t* out;
    uint64_t n;            // key positions of the input (MODE 1: all positions N, whatever the slice keeps)
    int shift;             // bit position of this pass's digit inside the key
    const uint32_t* bins;  // this pass: [256]
    uint32_t* status;      // this pass: tiles x 256 words, zero at launch
    uint32_t* counter;     // this pass: tile ticket, zero at launch
    uint32_t tile0;        // MODE 1 / 2: first tile of key positions this launch covers (tile = tile0 + ticket)
};

inline size_t sweep_smem_bytes() {
    // keys of the tile (padded for the blocked -> striped transpose of the first pass), per-warp digit counters,
    // tile-exclusive digit offsets, output deltas, scan scratch, ticket, alphabet table
    return sizeof(uint64_t) * (kTile + kTile / 16) + sizeof(uint32_t) * (kWarps * kRadix + 2 * kRadix + 33 + 3) + 256;
}

// Lanes of the warp whose 8-bit digit equals mine, from eight ballots (one per digit bit) instead of match.any: the match
// instruction was the top stall of the first version of this kernel (36 % of the samples waited on its result).
// `vm` = ballot of the lanes that hold a key; the result is only meaningful for those.
__device__ __forceinline__ unsigned match_digit(unsigned dig, unsigned vm) {
    unsigned r = vm;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const bool p = (dig >> b) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, p);
        r &= p ? bal : ~bal;
    }
    return r;
}

// MODE 0: keys from `a.keys`.  MODE 1: keys made from the residues (EncodeSrc).  MODE 2: as MODE 1, but the "digit" is the
// key's slice of the rank space: one stable pass that groups a rank's share of the keys by destination rank (sharded build).
// Order of work in a tile: load -> per-warp digit counts (shared-memory atomics) -> offsets, tile counts PUBLISHED ->
// rank (ballots) -> keys to their place in the tile buffer -> look-back (by now the tiles before have published) -> out.
template <int MODE, typename RankT>
__global__ void __launch_bounds__(kThreads, 4) onesweep_kernel(SweepArgs a, EncodeSrc s) {
    PD_DYNAMIC_SMEM(smem_raw);
    uint64_t* stage = reinterpret_cast<uint64_t*>(smem_raw);                                       // kTile + kTile / 16
    uint32_t* wc = reinterpret_cast<uint32_t*>(smem_raw + sizeof(uint64_t) * (kTile + kTile / 16));  // [kWarps][kRadix]
    uint32_t* tile_excl = wc + kWarps * kRadix;                                                     // [kRadix]
    uint32_t* gdelta = tile_excl + kRadix;                                                          // [kRadix]
    uint32_t* scratch = gdelta + kRadix;                                                            // 33
    uint32_t* ticket = scratch + 33;                                                                // 1 (+2 pad)
    uint8_t* val = reinterpret_cast<uint8_t*>(ticket + 3);                                          // 256

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) *ticket = atomicAdd(a.counter, 1u);  // tiles in the order the CTAs start: every tile before mine is running or done
    for (unsigned i = tid; i < kWarps * kRadix; i += kThreads) wc[i] = 0;
    if (MODE != 0) val[tid] = s.vt.v[tid];
    __syncthreads();
    const uint32_t tile = *ticket;                      // status / look-back order
    const uint32_t ptile = tile + (MODE != 0 ? a.tile0 : 0u);  // tile of key positions
    const uint64_t tile_base = (uint64_t)ptile * kTile;
    const uint32_t tile_n = (uint32_t)((a.n - tile_base < (uint64_t)kTile) ? (a.n - tile_base) : (uint64_t)kTile);

    uint64_t key[kItems];
    if (MODE != 0) {
        // blocked encode (a thread rolls over kItems consecutive k-mers), transposed through shared memory to the striped
        // order the ranking wants; one pad word per 16 keeps both sides at the two wavefronts an 8-byte access needs
        const uint32_t l0 = tid * kItems;
        const uint32_t n = l0 >= tile_n ? 0u : (tile_n - l0 < (uint32_t)kItems ? tile_n - l0 : (uint32_t)kItems);
        encode_keys<RankT>(s, val, tile_base + l0, n, s.tile_gene[ptile], s.tile_gene[ptile + 1],
                           [&](int j, uint64_t kv) { stage[l0 + j + ((l0 + j) >> 4)] = kv; });
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kItems; j++) {
            const uint32_t local = warp * (32 * kItems) + j * 32 + lane;
            key[j] = stage[local + (local >> 4)];
        }
    } else {
#pragma unroll
        for (int j = 0; j < kItems; j++) {
            const uint32_t local = warp * (32 * kItems) + j * 32 + lane;
            key[j] = local < tile_n ? a.keys[tile_base + local] : kNoKey;
        }
    }
    const bool full = MODE == 0 && tile_n == (uint32_t)kTile;  // uniform: every lane holds a key in every round

    // ---- per-warp digit counts
    uint32_t* my_wc = wc + warp * kRadix;
#pragma unroll
    for (int j = 0; j < kItems; j++)
        if (key[j] != kNoKey) atomicAdd(&my_wc[MODE == 2 ? slice_of(s, key[j] >> s.seq_bits) : ((unsigned)(key[j] >> a.shift) & 0xFFu)], 1u);
    __syncthreads();  // (also: the buffer of the first pass's transpose is free again)

    // ---- per digit (thread = digit): exclusive scan over the warps, tile total, exclusive scan over the digits;
    // the tile's counts go out at once so that later tiles can look back over them while this one still ranks
    uint32_t cnt = 0;
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            const uint32_t t = wc[w * kRadix + tid];
            wc[w * kRadix + tid] = run;
            run += t;
        }
        cnt = run;
    }
    volatile uint32_t* st = a.status + (size_t)tile * kRadix + tid;
    *st = tile == 0 ? (kInclusive | cnt) : cnt + 1u;
    uint32_t tile_kept;
    const uint32_t ex = prims::block_excl_scan<kThreads>(cnt, scratch, &tile_kept);
    tile_excl[tid] = ex;
    __syncthreads();

    // ---- rank and place: key order inside the warp's 512 keys = (round j, lane).  The warp's counter of a digit runs
    // from its exclusive offset; the first lane of every peer group advances it
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        const bool has = key[j] != kNoKey;
        const unsigned d = MODE == 2 ? (has ? slice_of(s, key[j] >> s.seq_bits) : 0u) : ((unsigned)(key[j] >> a.shift) & 0xFFu);
        const unsigned vm = full ? 0xffffffffu : __ballot_sync(0xffffffffu, has);
        const unsigned peers = match_digit(d, vm);
        const unsigned leader = __ffs(peers) - 1;
        uint32_t base_cnt = 0;
        if (has && lane == leader) {
            base_cnt = my_wc[d];
            my_wc[d] = base_cnt + __popc(peers);
        }
        base_cnt = __shfl_sync(0xffffffffu, base_cnt, has ? leader : lane);
        if (has) stage[tile_excl[d] + base_cnt + __popc(peers & lt_mask)] = key[j];
        __syncwarp();
    }

    // ---- decoupled look-back over the tiles before mine, one digit per thread
    {
        uint32_t before = 0;
        if (tile != 0) {
            for (uint32_t t = tile; t-- > 0;) {
                volatile uint32_t* sp = a.status + (size_t)t * kRadix + tid;
                uint32_t v;
                while ((v = *sp) == 0u) {
                }
                if (v & kInclusive) {
                    before += v & ~kInclusive;
                    break;
                }
                before += v - 1u;
            }
            *st = kInclusive | (before + cnt);
        }
        gdelta[tid] = a.bins[tid] + before - ex;
    }
    __syncthreads();

    // ---- out, in runs of equal digit
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const uint32_t idx = i * kThreads + tid;
        if (idx < tile_kept) {
            const uint64_t kv = stage[idx];
            const unsigned d = MODE == 2 ? slice_of(s, kv >> s.seq_bits) : ((unsigned)(kv >> a.shift) & 0xFFu);
            a.out[(uint64_t)gdelta[d] + idx] = kv;
        }
    }
}

inline uint32_t tiles_of(uint64_t n) { return (uint32_t)((n + kTile - 1) / kTile); }
inline int passes_of(int rank_bits) { return std::max(1, (rank_bits + 7) / 8); }

}  // namespace sortk
}  // namespace pd
