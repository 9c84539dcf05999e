// Index-construction kernels: the device counterpart of preprocessSequences (reference ig/native/library.cpp:189-335).
//
//   gene_meta_kernel      per gene: k-mer count, (count, genome), input checks   library.cpp:250-262
//   byte_hist_kernel      alphabet histogram                                     library.cpp:216-228
//   encode_kernel         positional k-mer rank -> 64-bit key                    library.cpp:75-79,134-150,234-265
//   (prims.cuh)           stable LSD radix sort on the rank bits                  library.cpp:270-278
//   entry_count_kernel /  one read of the sorted keys each: count dedup          library.cpp:280-287
//   entry_apply_kernel    (equal (rank, gene) -> one entry with its multiplicity) and rank groups incl. the tail
//                         merge                                                  library.cpp:297-306
//   fwd_count_kernel /    per-gene forward lists (the transpose of the posting   library.cpp:314-328
//   fwd_partition_kernel/ lists) by counting sort on the gene: count, scan, partition by gene bucket, place
//   fwd_place_kernel
//   gene_visited_kernel   cost model: total_visited per gene                      library.cpp:327
//
// HBM layout produced (all SoA, 32-bit indices; the reference's 16-B kmer_rank / 24-B kmers_range records are gone):
//   post[U]     uint32 seq | bit 31 (count > 1)   entries sorted by (rank, seq): the posting lists, group after group
//   post_cnt[U] uint32 count                      read by the scoring kernels only where bit 31 is set (U/N > 0.999: rare)
//   fwd[R]      uint2 (group start, group length | own-count>1 flag) one per (gene, shared k-mer), genes ascending;
//               inside a gene: short posting lists first, then long, then huge ones (order inside a class is free)
//   fwd_cnt[R]  uint32 the gene's own multiplicity of that k-mer (written only where the flag is set)
//   fwd_ptr[S+1], cls[S] (three 21-bit class counts), meta[S] = (kseq_len, genome), visited[S] (uint64)
#pragma once

#include "pd_rt.h"
#include "prims.cuh"

namespace pd {
namespace ik {

struct ValTable {
    uint8_t v[256];
};

static const int kClsBits = 21;  // a gene holds < 2^20 k-mers
static const unsigned long long kClsMask = (1ull << kClsBits) - 1ull;

// ---- per gene: kseq_lengths, meta, checks.  flags[0]: bit 0 = offsets descend, bit 1 = a gene of 2^20 or more
// residues; flags[1] = largest genome id; flags[2] = longest gene in k-mers.  kseq has S + 1 entries (the last one 0) so
// its exclusive scan ends in N.
__global__ void __launch_bounds__(256) gene_meta_kernel(const uint64_t* __restrict__ off, const uint32_t* __restrict__ genome_ids,
                                                         uint32_t S, int k, uint32_t* __restrict__ kseq, uint2* __restrict__ meta,
                                                         uint32_t* __restrict__ flags, unsigned long long* __restrict__ n_kmers) {
    const uint32_t s = blockIdx.x * 256u + threadIdx.x;
    uint32_t err = 0, gid = 0, kl = 0;
    if (s < S) {
        const uint64_t o0 = off[s], o1 = off[s + 1];
        if (o1 < o0) {
            err = 1;
        } else {
            const uint64_t len = o1 - o0;
            if (len >= (1ull << 20)) err = 2;
            else if (len >= (uint64_t)k) kl = (uint32_t)(len - (uint64_t)k + 1);  // library.cpp:250
        }
        gid = genome_ids[s];
        kseq[s] = kl;
        meta[s] = make_uint2(kl, gid);
    } else if (s == S) {
        kseq[S] = 0;
    }
    uint32_t ksum = kl;   // per warp < 32 * 2^20: no overflow
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        err |= __shfl_xor_sync(0xffffffffu, err, d);
        const uint32_t o = __shfl_xor_sync(0xffffffffu, gid, d);
        gid = o > gid ? o : gid;
        const uint32_t ok = __shfl_xor_sync(0xffffffffu, kl, d);
        kl = ok > kl ? ok : kl;
        ksum += __shfl_xor_sync(0xffffffffu, ksum, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (err) atomicOr(&flags[0], err);
        atomicMax(&flags[1], gid);
        atomicMax(&flags[2], kl);
        if (ksum) atomicAdd(n_kmers, (unsigned long long)ksum);   // the exact k-mer total in 64 bits (the 32-bit scan may wrap)
    }
}

// ---- sharded build: key positions [pos[0], pos[1]) of the genes whose residues start in [t0, t1) (t1 = all ones: to the end)
__global__ void share_bounds_kernel(const uint64_t* __restrict__ off, const uint32_t* __restrict__ key_off, uint32_t S, uint64_t t0, uint64_t t1,
                                    uint32_t* __restrict__ pos) {
    if (blockIdx.x != 0 || threadIdx.x > 1) return;
    const uint64_t t = threadIdx.x == 0 ? t0 : t1;
    uint32_t lo = 0, hi = S;   // first gene g in [0, S] with off[g] >= t (S when none)
    if (t == ~0ull) lo = S;
    else
        while (lo < hi) {
            const uint32_t m = lo + (hi - lo) / 2;
            if (off[m] < t) lo = m + 1;
            else hi = m;
        }
    pos[threadIdx.x] = key_off[lo];
}

// ---- alphabet histogram: 16-B loads, per-warp privatised shared histograms
__global__ void __launch_bounds__(256) byte_hist_kernel(const uint8_t* __restrict__ res, uint64_t n,
                                                         unsigned long long* __restrict__ hist) {
    __shared__ uint32_t h[8][256];
    const unsigned tid = threadIdx.x;
    for (unsigned i = tid; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
    __syncthreads();
    uint32_t* mine = h[tid >> 5];
    // head up to 16-B alignment, vector body, scalar tail
    uint64_t head = (16 - (reinterpret_cast<uintptr_t>(res) & 15)) & 15;
    if (head > n) head = n;
    const uint64_t nvec = (n - head) / 16;
    const uint4* vp = reinterpret_cast<const uint4*>(res + head);
    const uint64_t gtid = (uint64_t)blockIdx.x * 256 + tid, gsz = (uint64_t)gridDim.x * 256;
    for (uint64_t i = gtid; i < nvec; i += gsz) {
        uint4 q = vp[i];
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            atomicAdd(&mine[w[j] & 0xFF], 1u);
            atomicAdd(&mine[(w[j] >> 8) & 0xFF], 1u);
            atomicAdd(&mine[(w[j] >> 16) & 0xFF], 1u);
            atomicAdd(&mine[w[j] >> 24], 1u);
        }
    }
    const uint64_t tail0 = head + nvec * 16;
    for (uint64_t i = gtid; i < head; i += gsz) atomicAdd(&mine[res[i]], 1u);
    for (uint64_t i = tail0 + gtid; i < n; i += gsz) atomicAdd(&mine[res[i]], 1u);
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) s += h[w][tid];
    if (s) atomicAdd(&hist[tid], (unsigned long long)s);
}

// ---- k-mer encode: one warp per gene, lanes stride over positions; key = rank << seq_bits | gene
__global__ void __launch_bounds__(256) encode_kernel(const uint8_t* __restrict__ res, const uint64_t* __restrict__ gene_off,
                                                      const uint32_t* __restrict__ key_off, uint32_t S, int k, uint32_t base,
                                                      int seq_bits, ValTable vt, uint64_t* __restrict__ keys) {
    __shared__ uint8_t val[256];
    val[threadIdx.x] = vt.v[threadIdx.x];
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * 256u) >> 5;
    for (uint32_t g = (blockIdx.x * 256u + threadIdx.x) >> 5; g < S; g += warps) {
        const uint32_t ko = key_off[g];
        const uint32_t nk = key_off[g + 1] - ko;
        const uint8_t* p = res + gene_off[g];
        for (uint32_t i = lane; i < nk; i += 32) {
            uint64_t r = 0;
            for (int j = 0; j < k; j++) r = r * base + val[p[i + j]];
            keys[(uint64_t)ko + i] = (r << seq_bits) | g;
        }
    }
}

// ---- count dedup + rank groups, two reads of the sorted keys
//
// A key position is an entry head if its (rank, gene) differs from the position before, and a group head if its rank
// differs.  entry_count_kernel counts both per tile; after a scan of the tile counts entry_apply_kernel writes, per
// entry e: post[e] (gene, bit 31 if the key repeats), post_cnt[e], ent_gid[e] = its rank group, and grp_head[g] for
// the entry that opens group g.  The tail merge of the reference (library.cpp:300-306: at the last entry the open run
// [start, i+1) is closed whatever its rank) means the LAST entry never opens a group of its own unless it is the only
// entry: it joins the group before it, and *spurious tells the host that one counted group head was dropped.
static const int kEntThreads = 512;
static const int kEntItems = 8;
static const int kEntTile = kEntThreads * kEntItems;  // 4096 keys per block
static const int kEntWarps = kEntThreads / 32;

// Layout: warp w of a block owns keys [tile + w * 256, + 256); in round j its lanes hold 32 consecutive keys, so
// every load and store is coalesced and prefix counts inside a round are popcounts of ballots.
// Flags of this lane's key in round j: hm / gm = ballots of "entry head" / "group head" over the warp.
__device__ __forceinline__ void entry_round(const uint64_t* __restrict__ keys, uint64_t N, int seq_bits, uint64_t idx, uint64_t& kv,
                                            unsigned& hm, unsigned& gm) {
    const unsigned lane = threadIdx.x & 31;
    const bool in = idx < N;
    kv = in ? keys[idx] : 0ull;
    uint64_t prev = __shfl_up_sync(0xffffffffu, kv, 1);
    if (lane == 0) prev = (idx > 0 && in) ? keys[idx - 1] : 0ull;
    const bool first = idx == 0;
    hm = __ballot_sync(0xffffffffu, in && (first || kv != prev));
    gm = __ballot_sync(0xffffffffu, in && (first || (kv >> seq_bits) != (prev >> seq_bits)));
}

__global__ void __launch_bounds__(kEntThreads) entry_count_kernel(const uint64_t* __restrict__ keys, uint64_t N, int seq_bits,
                                                                   uint32_t* __restrict__ tile_heads, uint32_t* __restrict__ tile_gheads) {
    __shared__ uint32_t s_h, s_g;
    if (threadIdx.x == 0) s_h = s_g = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * kEntTile + (uint64_t)warp * (32 * kEntItems);
    uint32_t h = 0, g = 0;
#pragma unroll
    for (int j = 0; j < kEntItems; j++) {
        uint64_t kv;
        unsigned hm, gm;
        entry_round(keys, N, seq_bits, base + j * 32 + lane, kv, hm, gm);
        h += __popc(hm);
        g += __popc(gm);
    }
    if (lane == 0) {
        atomicAdd(&s_h, h);
        atomicAdd(&s_g, g);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        tile_heads[blockIdx.x] = s_h;
        tile_gheads[blockIdx.x] = s_g;
    }
}

// What a rank of a sharded build (rank-range slices, sort_kernels.cuh) sends to the others besides its postings: one bit per
// entry, set where the entry opens a rank group, and the list of its entries held more than once (position, count).
// All null / zero in a single-GPU build.
struct ShardOut {
    uint32_t* head_bits;   // zero-initialised, (U + 31) / 32 words
    uint32_t* multi;       // pairs (entry, count)
    uint32_t* n_multi;     // running length of `multi`
    uint32_t multi_cap;    // pairs that fit (the counter keeps running)
    uint32_t tail_merge;   // 0: this slice does not end the entry list, the reference's tail merge (library.cpp:300-306) is not its business
};

__global__ void __launch_bounds__(kEntThreads, 2) entry_apply_kernel(const uint64_t* __restrict__ keys, uint64_t N, int seq_bits,
                                                                   const uint32_t* __restrict__ tile_head_off,
                                                                   const uint32_t* __restrict__ tile_ghead_off, uint32_t U,
                                                                   uint32_t* __restrict__ post, uint32_t* __restrict__ post_cnt,
                                                                   uint32_t* __restrict__ ent_gid, uint32_t* __restrict__ grp_head,
                                                                   uint64_t* __restrict__ ent_rank, uint32_t* __restrict__ spurious,
                                                                   ShardOut so) {
    __shared__ uint32_t w_h[kEntWarps], w_g[kEntWarps];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const uint64_t base = (uint64_t)blockIdx.x * kEntTile + (uint64_t)warp * (32 * kEntItems);
    uint64_t kv[kEntItems];
    unsigned hm[kEntItems], gm[kEntItems];
    uint32_t h = 0, g = 0;
#pragma unroll
    for (int j = 0; j < kEntItems; j++) {
        entry_round(keys, N, seq_bits, base + j * 32 + lane, kv[j], hm[j], gm[j]);
        h += __popc(hm[j]);
        g += __popc(gm[j]);
    }
    if (lane == 0) {
        w_h[warp] = h;
        w_g[warp] = g;
    }
    __syncthreads();
    uint32_t e0 = tile_head_off[blockIdx.x], g0 = tile_ghead_off[blockIdx.x];  // entries / group heads before this warp's keys
    for (unsigned w = 0; w < warp; w++) {
        e0 += w_h[w];
        g0 += w_g[w];
    }
    const uint64_t seq_mask = (1ull << seq_bits) - 1ull;
#pragma unroll
    for (int j = 0; j < kEntItems; j++) {
        const uint64_t idx = base + j * 32 + lane;
        // the key after mine (equal keys follow each other): lane + 1, or the next round's / warp's first key
        uint64_t nxt = __shfl_down_sync(0xffffffffu, kv[j], 1);
        if (lane == 31) nxt = idx + 1 < N ? keys[idx + 1] : ~kv[j];
        if ((hm[j] >> lane) & 1u) {
            const uint32_t e = e0 + __popc(hm[j] & lt);
            bool gh = (gm[j] >> lane) & 1u;
            uint32_t gid = g0 + __popc(gm[j] & lt) + (gh ? 1u : 0u) - 1u;
            if (gh && e == U - 1 && e != 0 && so.tail_merge) {  // the tail merge: the last entry joins the group before it
                gh = false;
                gid -= 1;
                *spurious = 1;
            }
            uint32_t cnt = 1;
            if (idx + 1 < N && nxt == kv[j]) {
                cnt = 2;
                while (idx + cnt < N && keys[idx + cnt] == kv[j]) cnt++;
            }
            post[e] = (uint32_t)(kv[j] & seq_mask) | (cnt > 1 ? 0x80000000u : 0u);
            post_cnt[e] = cnt;
            ent_gid[e] = gid;
            if (gh) grp_head[gid] = e;
            if (so.head_bits) {
                if (gh) atomicOr(&so.head_bits[e >> 5], 1u << (e & 31u));
                if (cnt > 1) {
                    const uint32_t m = atomicAdd(so.n_multi, 1u);
                    if (m < so.multi_cap) {
                        so.multi[2 * (size_t)m] = e;
                        so.multi[2 * (size_t)m + 1] = cnt;
                    }
                }
            }
            if (ent_rank) ent_rank[e] = kv[j] >> seq_bits;
        }
        e0 += __popc(hm[j]);
        g0 += __popc(gm[j]);
    }
}

// grp_head[groups] = U, with groups = counted group heads minus the one the tail merge dropped
__global__ void group_tail_kernel(uint32_t* __restrict__ grp_head, uint32_t counted, const uint32_t* __restrict__ spurious, uint32_t U) {
    if (blockIdx.x == 0 && threadIdx.x == 0) grp_head[counted - *spurious] = U;
}

// ---- forward lists by counting sort on the gene.  List classes: 0 short (<= short_max), 1 long, 2 huge (> huge_min).
__device__ __forceinline__ uint32_t list_class(uint32_t gl, uint32_t short_max, uint32_t huge_min) {
    return gl <= short_max ? 0u : (gl > huge_min ? 2u : 1u);
}

// cls[gene] += 1 in the field of the entry's list class, for every entry of a shared group (length >= 2).
// kFwdItems entries per thread, all their gathers issued before the first atomic: the kernel is bound by the latency of
// the dependent gathers (ent_gid -> grp_head), not by DRAM (profiles/r01_fwd_kernels_ncu_full.txt).
static const int kFwdItems = 4;
__global__ void __launch_bounds__(256) fwd_count_kernel(const uint32_t* __restrict__ post, const uint32_t* __restrict__ ent_gid,
                                                         const uint32_t* __restrict__ grp_head, uint32_t U, uint32_t short_max,
                                                         uint32_t huge_min, unsigned long long* __restrict__ cls,
                                                         unsigned long long* __restrict__ visited) {
    const uint64_t e0 = (uint64_t)blockIdx.x * (256u * kFwdItems) + threadIdx.x;
    uint32_t g[kFwdItems], gl[kFwdItems], p[kFwdItems];
#pragma unroll
    for (int j = 0; j < kFwdItems; j++) {
        const uint64_t e = e0 + (uint64_t)j * 256u;
        g[j] = e < U ? ent_gid[e] : 0xFFFFFFFFu;
        p[j] = e < U ? post[e] : 0u;
    }
#pragma unroll
    for (int j = 0; j < kFwdItems; j++) gl[j] = g[j] != 0xFFFFFFFFu ? grp_head[g[j] + 1] - grp_head[g[j]] : 0u;
#pragma unroll
    for (int j = 0; j < kFwdItems; j++)
        if (gl[j] >= 2) {
            atomicAdd(&cls[p[j] & 0x7FFFFFFFu], 1ull << (kClsBits * list_class(gl[j], short_max, huge_min)));
            if (visited) atomicAdd(&visited[p[j] & 0x7FFFFFFFu], (unsigned long long)gl[j]);  // sharded build: this slice's part of total_visited
        }
}

// tot[s] = forward entries of gene s (tot has S + 1 entries, the last one 0: its exclusive scan is fwd_ptr)
// (genes outside [row0, row1) get none: a rank of a sharded build keeps forward lists for its own query rows only)
__global__ void __launch_bounds__(256) fwd_totals_kernel(const unsigned long long* __restrict__ cls, uint32_t S, uint32_t row0, uint32_t row1,
                                                          uint32_t* __restrict__ tot) {
    const uint32_t s = blockIdx.x * 256u + threadIdx.x;
    if (s > S) return;
    uint32_t t = 0;
    if (s < S && s >= row0 && s < row1) {
        const unsigned long long c = cls[s];
        t = (uint32_t)(c & kClsMask) + (uint32_t)((c >> kClsBits) & kClsMask) + (uint32_t)((c >> (2 * kClsBits)) & kClsMask);
    }
    tot[s] = t;
}

// Placing ~R records at random 8-byte slots of a multi-gigabyte array is DRAM-page-miss bound, so the transpose runs
// in two steps: (1) fwd_partition_kernel appends one 16-byte record per shared entry to the region of its gene
// BUCKET (2^bshift consecutive genes; the region of bucket b is exactly the forward-list range of its genes, so
// fwd_ptr gives the region bounds for free); (2) fwd_place_kernel reads the records region after region and places
// them: now all writes of the CTAs in flight fall into one bucket's window of `fwd`, which the L2 holds.
static const int kPartThreads = 512;
static const int kPartItems = 4;
static const int kPartTile = kPartThreads * kPartItems;
static const int kMaxBuckets = 512;

struct FwdRecord {  // 16 B
    uint32_t gene, gs, gl_multi, cls_cnt;  // gl | own-count>1 flag; list class | own count << 2
};

__global__ void __launch_bounds__(kPartThreads) fwd_partition_kernel(const uint32_t* __restrict__ post, const uint32_t* __restrict__ post_cnt,
                                                                      const uint32_t* __restrict__ ent_gid,
                                                                      const uint32_t* __restrict__ grp_head, uint32_t U, uint32_t S,
                                                                      uint32_t short_max, uint32_t huge_min, uint32_t bshift,
                                                                      const uint32_t* __restrict__ fwd_ptr, uint32_t* __restrict__ bucket_cur,
                                                                      uint4* __restrict__ records, uint32_t row0, uint32_t row1,
                                                                      const uint32_t* __restrict__ head_bits, const uint32_t* __restrict__ tile_head_off) {
    __shared__ uint32_t cnt[kMaxBuckets];
    __shared__ uint32_t base[kMaxBuckets];
    __shared__ uint32_t wpre[kPartTile / 32 + 1];  // head bits given instead of ent_gid: group heads before every word of this tile
    __shared__ uint32_t wbits[kPartTile / 32];
    const unsigned tid = threadIdx.x;
    for (unsigned i = tid; i < kMaxBuckets; i += kPartThreads) cnt[i] = 0;
    const uint32_t e0 = blockIdx.x * (uint32_t)kPartTile;
    if (head_bits) {
        // sharded build: an entry's group = number of group heads at or before it, minus one (the tile = kPartTile entries =
        // kPartTile / 32 words of head bits; tile_head_off = heads before the tile)
        if (tid < kPartTile / 32) {
            const uint64_t w = (uint64_t)blockIdx.x * (kPartTile / 32) + tid;
            const uint32_t b = (w * 32 < U) ? head_bits[w] : 0u;
            wbits[tid] = b;
            uint32_t incl = (uint32_t)__popc(b);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if ((tid & 31u) >= (unsigned)d) incl += o;
            }
            wpre[tid + 1] = incl;  // inclusive inside the warp; the second warp adds the first one's total below
        }
        __syncthreads();
        if (tid >= 32 && tid < kPartTile / 32) wpre[tid + 1] += wpre[32];
        if (tid == 0) wpre[0] = 0;
    }
    __syncthreads();
    FwdRecord rec[kPartItems];
    uint32_t pos[kPartItems];
    unsigned have = 0;
    // the gathers level by level over all items (entry -> group -> group bounds), then the ranking atomics: the chains
    // of the items overlap instead of running one after the other.  Entries of genes outside [row0, row1) drop out first.
#pragma unroll
    for (int j = 0; j < kPartItems; j++) {
        const uint32_t l = j * kPartThreads + tid;
        const uint32_t e = e0 + l;
        const bool in = e < U;
        rec[j].gene = in ? post[e] : 0u;   // gene | bit 31
        const uint32_t gene = rec[j].gene & 0x7FFFFFFFu;
        const bool want = in && gene >= row0 && gene < row1;
        uint32_t g = 0;
        if (want) {
            if (head_bits) {
                const uint32_t wi = l >> 5, b = l & 31u;
                g = tile_head_off[blockIdx.x] + wpre[wi] + (uint32_t)__popc(wbits[wi] & (0xFFFFFFFFu >> (31u - b))) - 1u;
            } else {
                g = ent_gid[e];
            }
            have |= 1u << j;
        }
        rec[j].gs = g;  // the group id for now
    }
#pragma unroll
    for (int j = 0; j < kPartItems; j++) {
        const uint32_t g = rec[j].gs;
        const uint32_t gs = (have >> j & 1u) ? grp_head[g] : 0u;
        const uint32_t gl = (have >> j & 1u) ? grp_head[g + 1] - gs : 0u;
        rec[j].gs = gs;
        rec[j].gl_multi = gl;
        if (gl < 2) have &= ~(1u << j);
    }
#pragma unroll
    for (int j = 0; j < kPartItems; j++) {
        if (!(have >> j & 1u)) continue;
        const uint32_t p = rec[j].gene, gl = rec[j].gl_multi;
        const uint32_t e = e0 + j * kPartThreads + tid;
        rec[j].gene = p & 0x7FFFFFFFu;
        rec[j].gl_multi = gl | (p & 0x80000000u);
        rec[j].cls_cnt = list_class(gl, short_max, huge_min) | ((p & 0x80000000u) ? (post_cnt[e] << 2) : (1u << 2));
        pos[j] = atomicAdd(&cnt[rec[j].gene >> bshift], 1u);
    }
    __syncthreads();
    const uint32_t nb = ((S - 1) >> bshift) + 1;
    for (unsigned b = tid; b < nb; b += kPartThreads) {
        const uint32_t c = cnt[b];
        base[b] = c ? fwd_ptr[b << bshift] + atomicAdd(&bucket_cur[b], c) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPartItems; j++) {
        if (have & (1u << j)) {
            const FwdRecord& r = rec[j];
            records[base[r.gene >> bshift] + pos[j]] = make_uint4(r.gene, r.gs, r.gl_multi, r.cls_cnt);
        }
    }
}

// cur3[3 s + c] = first slot of gene s's class-c segment (short, long, huge): fwd_place_kernel only has to bump it
__global__ void __launch_bounds__(256) fwd_cursor_init_kernel(const unsigned long long* __restrict__ cls, const uint32_t* __restrict__ fwd_ptr,
                                                               uint32_t S, uint32_t* __restrict__ cur3) {
    const uint32_t s = blockIdx.x * 256u + threadIdx.x;
    if (s >= S) return;
    const unsigned long long cl = cls[s];
    const uint32_t b = fwd_ptr[s], n0 = (uint32_t)(cl & kClsMask), n1 = (uint32_t)((cl >> kClsBits) & kClsMask);
    cur3[3 * (size_t)s] = b;
    cur3[3 * (size_t)s + 1] = b + n0;
    cur3[3 * (size_t)s + 2] = b + n0 + n1;
}

// record -> its gene's forward list: one atomic on the segment's running slot, one 8-byte store (the per-gene bases are
// folded into the cursors beforehand: two fewer L2 gathers per record than reading cls and fwd_ptr here)
// kFwdItems records per thread: all loads, then all atomics, then all stores, so that a thread keeps several of the
// dependent chains (record -> cursor atomic -> store) in flight.
__global__ void __launch_bounds__(256) fwd_place_kernel(const uint4* __restrict__ records, uint32_t R, uint32_t* __restrict__ cur3,
                                                         uint2* __restrict__ fwd, uint32_t* __restrict__ fwd_cnt) {
    const uint64_t i0 = (uint64_t)blockIdx.x * (256u * kFwdItems) + threadIdx.x;
    uint4 r[kFwdItems];
    uint32_t slot[kFwdItems];
#pragma unroll
    for (int j = 0; j < kFwdItems; j++) {
        const uint64_t i = i0 + (uint64_t)j * 256u;
        r[j] = i < R ? records[i] : make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
    }
#pragma unroll
    for (int j = 0; j < kFwdItems; j++)
        slot[j] = r[j].x != 0xFFFFFFFFu ? atomicAdd(&cur3[3 * (size_t)r[j].x + (r[j].w & 3u)], 1u) : 0u;
#pragma unroll
    for (int j = 0; j < kFwdItems; j++) {
        if (r[j].x == 0xFFFFFFFFu) continue;
        fwd[slot[j]] = make_uint2(r[j].y, r[j].z);  // bit 31 of the length: the gene's own multiplicity > 1 (then fwd_cnt is read)
        if (r[j].z & 0x80000000u) fwd_cnt[slot[j]] = r[j].w >> 2;
    }
}

// visited[s] = sum of the posting-list lengths of gene s's forward entries (computation_costs[].total_visited,
// library.cpp:327); *total += the sum over all genes ("Total cost: N lookups", library.cpp:349).
// fam_key[s] = the smallest posting-list start among the gene's long lists (all lists if it has no long one): genes
// of one family share their conserved k-mers, hence mostly this key; scoring calls process their rows in fam_key
// order so that rows running at the same time read the same posting lists (L2 hits instead of HBM reads).
// One warp per gene.
__global__ void __launch_bounds__(256) gene_visited_kernel(const uint2* __restrict__ fwd, const uint32_t* __restrict__ fwd_ptr, uint32_t s0, uint32_t S,
                                                            uint32_t short_max, unsigned long long* __restrict__ visited,
                                                            uint32_t* __restrict__ fam_key, unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const uint32_t s = s0 + ((blockIdx.x * 256u + threadIdx.x) >> 5);   // genes [s0, S)
    if (s < S) {
        const uint32_t f0 = fwd_ptr[s], f1 = fwd_ptr[s + 1];
        unsigned long long v = 0;
        uint32_t kl = 0x7FFFFFFFu, ka = 0x7FFFFFFFu;
        for (uint32_t f = f0 + lane; f < f1; f += 32) {
            const uint2 fw = fwd[f];
            const uint32_t gl = fw.y & 0x7FFFFFFFu;
            v += gl;
            ka = fw.x < ka ? fw.x : ka;
            if (gl > short_max) kl = fw.x < kl ? fw.x : kl;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, d);
            const uint32_t ol = __shfl_xor_sync(0xffffffffu, kl, d), oa = __shfl_xor_sync(0xffffffffu, ka, d);
            kl = ol < kl ? ol : kl;
            ka = oa < ka ? oa : ka;
        }
        if (lane == 0) {
            if (visited) visited[s] = v;   // (null in a sharded build: the all-reduced totals are already there)
            fam_key[s] = kl != 0x7FFFFFFFu ? kl : ka;
            if (v) atomicAdd(&s_sum, v);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_sum && total) atomicAdd(total, s_sum);
}

// ---- sharded build, after the all-gather: the entry list is `segs` segments of `seg` entries, segment r = rank r's slice
// padded with filler entries (no gene, each its own group).
// fill: the padding of this rank's own segment, before it is sent
__global__ void __launch_bounds__(256) shard_pad_kernel(uint32_t* __restrict__ post, uint32_t* __restrict__ head_bits, uint32_t U, uint32_t seg) {
    const uint32_t e = U + blockIdx.x * 256u + threadIdx.x;
    if (e >= seg) return;
    post[e] = 0x7FFFFFFFu;
    atomicOr(&head_bits[e >> 5], 1u << (e & 31u));
}
// group heads per tile of kPartTile entries (the tiles of fwd_partition_kernel), then (after a scan) every group's first entry
static const int kHeadTileWords = kPartTile / 32;
__global__ void __launch_bounds__(kHeadTileWords) head_count_kernel(const uint32_t* __restrict__ head_bits, uint64_t words, uint32_t* __restrict__ tile_heads) {
    __shared__ uint32_t s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const uint64_t w = (uint64_t)blockIdx.x * kHeadTileWords + threadIdx.x;
    uint32_t n = w < words ? (uint32_t)__popc(head_bits[w]) : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&s_n, n);
    __syncthreads();
    if (threadIdx.x == 0) tile_heads[blockIdx.x] = s_n;
}
__global__ void __launch_bounds__(kHeadTileWords) head_apply_kernel(const uint32_t* __restrict__ head_bits, uint64_t words,
                                                                     const uint32_t* __restrict__ tile_head_off, uint32_t* __restrict__ grp_head) {
    __shared__ uint32_t scratch[33];
    const uint64_t w = (uint64_t)blockIdx.x * kHeadTileWords + threadIdx.x;
    uint32_t bits = w < words ? head_bits[w] : 0u;
    uint32_t tot;
    uint32_t g = tile_head_off[blockIdx.x] + prims::block_excl_scan<kHeadTileWords>((uint32_t)__popc(bits), scratch, &tot);  // heads before this word
    while (bits) {
        const uint32_t b = (uint32_t)__ffs((int)bits) - 1u;
        bits &= bits - 1u;
        grp_head[g++] = (uint32_t)(w * 32 + b);
    }
}
// post_cnt at the positions of the gathered multi lists (segment r: pairs (entry inside rank r's slice, count))
__global__ void __launch_bounds__(256) shard_multi_kernel(const uint32_t* __restrict__ multi, const uint32_t* __restrict__ n_of_seg, uint32_t mseg,
                                                           uint32_t seg, uint32_t segs, uint32_t* __restrict__ post_cnt) {
    const uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    const uint32_t r = (uint32_t)(i / mseg), j = (uint32_t)(i % mseg);
    if (r >= segs || j >= n_of_seg[r]) return;
    const uint32_t* m = multi + 2 * ((size_t)r * mseg + j);
    post_cnt[(size_t)r * seg + m[0]] = m[1];
}
// cost[genome] += total_visited + 1 of its genes (query partitioning by posting-list volume, library.cpp:327)
// first[genome] = its smallest gene, count[genome] = its genes (the caller checks that genomes are contiguous gene ranges)
__global__ void __launch_bounds__(256) genome_cost_kernel(const unsigned long long* __restrict__ visited, const uint2* __restrict__ meta, uint32_t S,
                                                           unsigned long long* __restrict__ cost, unsigned long long* __restrict__ total,
                                                           uint32_t* __restrict__ first, uint32_t* __restrict__ count) {
    const uint32_t s = blockIdx.x * 256u + threadIdx.x;
    const bool in = s < S;
    const uint32_t g = in ? meta[s].y : 0xFFFFFFFFu;
    if (in) {
        const uint32_t gp = s ? meta[s - 1].y : 0xFFFFFFFFu;
        if (gp != g) atomicMin(&first[g], s);   // one atomic per run of equal genomes
    }
    unsigned long long v = in ? visited[s] : 0ull;
    // genes of a genome sit together: a warp whose lanes all belong to one genome adds once
    const uint32_t g0 = __shfl_sync(0xffffffffu, g, 0);
    const bool uniform = __all_sync(0xffffffffu, g == g0);
    if (!uniform && in) {
        atomicAdd(&cost[g], v + 1ull);
        atomicAdd(&count[g], 1u);
    }
    unsigned long long c = in ? v + 1ull : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, d);
        c += __shfl_xor_sync(0xffffffffu, c, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (uniform && g0 != 0xFFFFFFFFu) {
            atomicAdd(&cost[g0], c);
            atomicAdd(&count[g0], 32u);
        }
        if (v) atomicAdd(total, v);
    }
}

// per-entry group (start, length) for pd_entries (tests / diagnostics)
__global__ void __launch_bounds__(256) entry_groups_kernel(const uint32_t* __restrict__ ent_gid, const uint32_t* __restrict__ grp_head,
                                                            uint32_t U, uint32_t* __restrict__ gs_out, uint32_t* __restrict__ gl_out) {
    const uint32_t e = blockIdx.x * 256u + threadIdx.x;
    if (e >= U) return;
    const uint32_t g = ent_gid[e];
    const uint32_t gs = grp_head[g];
    gs_out[e] = gs;
    gl_out[e] = grp_head[g + 1] - gs;
}

}  // namespace ik
}  // namespace pd
