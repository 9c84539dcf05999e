// Index-construction kernels: the device counterpart of preprocessSequences (reference ig/native/library.cpp:189-335).
//
//   byte_hist_kernel      alphabet histogram                      library.cpp:216-228
//   encode_kernel         positional k-mer rank -> 64-bit key     library.cpp:75-79,134-150,234-265
//   (prims.cuh)           stable LSD radix sort on the rank bits   library.cpp:270-278
//   head_scatter_kernel   run heads of equal (rank, seq)          library.cpp:280-287   (count dedup, part 1)
//   entries_kernel        count per entry + rank-run heads         library.cpp:280-287, 297-306 (incl. the tail merge)
//   group_*_kernel        rank groups -> inverted index CSR        library.cpp:297-335
//   fwd_*_kernel          per-gene forward lists + cost model      library.cpp:314-328
//
// HBM layout produced (all SoA, 32-bit indices; the reference's 16-B kmer_rank / 24-B kmers_range records are gone):
//   post[U]     uint32 seq | bit 31 (count > 1)   entries sorted by (rank, seq): the posting lists, group after group
//   post_cnt[U] uint32 count                      read by the scoring kernels only where bit 31 is set (U/N > 0.999: rare)
//   fwd[R]      uint2 (group start, group length | own-count>1 flag) one per (gene, shared k-mer), genes ascending;
//               inside a gene: short posting lists first, then long, then huge ones, ranks ascending in each class
//   fwd_cnt[R]  uint32 the gene's own multiplicity of that k-mer
//   fwd_ptr[S+1], gene_short[S], gene_huge[S], meta[S] = (kseq_len, genome), visited[S] (uint64)
#pragma once

#include "pd_rt.h"
#include "prims.cuh"

namespace pd {
namespace ik {

struct ValTable {
    uint8_t v[256];
};

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
                                                      const uint64_t* __restrict__ key_off, uint32_t S, int k, uint32_t base,
                                                      int seq_bits, ValTable vt, uint64_t* __restrict__ keys) {
    __shared__ uint8_t val[256];
    val[threadIdx.x] = vt.v[threadIdx.x];
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * 256u) >> 5;
    for (uint32_t g = (blockIdx.x * 256u + threadIdx.x) >> 5; g < S; g += warps) {
        const uint64_t ko = key_off[g];
        const uint32_t nk = (uint32_t)(key_off[g + 1] - ko);
        const uint8_t* p = res + gene_off[g];
        for (uint32_t i = lane; i < nk; i += 32) {
            uint64_t r = 0;
            for (int j = 0; j < k; j++) r = r * base + val[p[i + j]];
            keys[ko + i] = (r << seq_bits) | g;
        }
    }
}

// ---- dedup part 1: flags[i] = 1 iff key i starts a run of equal keys
__global__ void __launch_bounds__(256) head_flag_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ flags) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// ent_pos[e] = position in the sorted key list of the e-th distinct key
__global__ void __launch_bounds__(256) head_scatter_kernel(const uint64_t* __restrict__ keys, uint64_t n,
                                                            const uint32_t* __restrict__ excl, uint32_t* __restrict__ ent_pos) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n && (i == 0 || keys[i] != keys[i - 1])) ent_pos[excl[i]] = (uint32_t)i;
}

// per entry: posting (seq, count); rflag = 1 iff the entry opens a rank group.  The tail merge of the reference
// (library.cpp:300-306: at the last entry the open run [start, i+1) is closed whatever its rank) means the last
// entry never opens a group of its own unless it is the only entry.
__global__ void __launch_bounds__(256) entries_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ent_pos,
                                                       uint32_t U, uint64_t N, int seq_bits, uint32_t* __restrict__ post,
                                                       uint32_t* __restrict__ post_cnt, uint32_t* __restrict__ rflag,
                                                       uint64_t* __restrict__ ent_rank) {
    const uint32_t e = blockIdx.x * 256u + threadIdx.x;
    if (e >= U) return;
    const uint32_t p = ent_pos[e];
    const uint32_t pn = (e + 1 < U) ? ent_pos[e + 1] : (uint32_t)N;
    const uint64_t key = keys[p];
    const uint64_t rank = key >> seq_bits;
    const uint32_t cnt = pn - p;
    post[e] = (uint32_t)(key & ((1ull << seq_bits) - 1ull)) | (cnt > 1 ? 0x80000000u : 0u);
    post_cnt[e] = cnt;
    uint32_t f;
    if (e == 0)
        f = 1;
    else if (e == U - 1)
        f = 0;
    else
        f = ((keys[ent_pos[e - 1]] >> seq_bits) != rank) ? 1u : 0u;
    rflag[e] = f;
    if (ent_rank) ent_rank[e] = rank;
}

// grp_head[g] = first entry of group g; ent_gid[e] = group of entry e
__global__ void __launch_bounds__(256) group_heads_kernel(const uint32_t* __restrict__ rflag, const uint32_t* __restrict__ excl,
                                                           uint32_t U, uint32_t* __restrict__ grp_head, uint32_t* __restrict__ ent_gid) {
    const uint32_t e = blockIdx.x * 256u + threadIdx.x;
    if (e >= U) return;
    const uint32_t f = rflag[e];
    const uint32_t g = excl[e] + f - 1u;
    ent_gid[e] = g;
    if (f) grp_head[g] = e;
}

// per entry in a shared group (length >= 2): mark it, count it for its gene, add the group length to the gene's
// cost (computation_costs[].total_visited, library.cpp:327).
__global__ void __launch_bounds__(256) shared_mark_kernel(const uint32_t* __restrict__ post, const uint32_t* __restrict__ ent_gid,
                                                           const uint32_t* __restrict__ grp_head, uint32_t U,
                                                           uint32_t* __restrict__ sflag, uint32_t* __restrict__ gene_cnt,
                                                           uint32_t* __restrict__ gene_short, uint32_t short_max,
                                                           uint32_t* __restrict__ gene_huge, uint32_t huge_min,
                                                           unsigned long long* __restrict__ visited) {
    const uint32_t e = blockIdx.x * 256u + threadIdx.x;
    if (e >= U) return;
    const uint32_t g = ent_gid[e];
    const uint32_t gl = grp_head[g + 1] - grp_head[g];
    const uint32_t gene = post[e] & 0x7FFFFFFFu;
    const uint32_t s = gl >= 2 ? 1u : 0u;
    sflag[e] = s;
    if (s) {
        atomicAdd(&gene_cnt[gene], 1u);
        if (gl <= short_max) atomicAdd(&gene_short[gene], 1u);
        if (gl > huge_min) atomicAdd(&gene_huge[gene], 1u);
        atomicAdd(&visited[gene], (unsigned long long)gl);
    }
}

// compact the shared entries into sort keys (gene << 33 | list class << 31 | entry): sorted on the gene AND the
// class, a gene's forward list holds its short posting lists first, then the long ones, then the huge ones, ranks
// ascending in each part
__global__ void __launch_bounds__(256) fwd_keys_kernel(const uint32_t* __restrict__ post, const uint32_t* __restrict__ sflag,
                                                        const uint32_t* __restrict__ excl, const uint32_t* __restrict__ ent_gid,
                                                        const uint32_t* __restrict__ grp_head, uint32_t short_max, uint32_t huge_min,
                                                        uint32_t U, uint64_t* __restrict__ fkeys) {
    const uint32_t e = blockIdx.x * 256u + threadIdx.x;
    if (e >= U) return;
    if (sflag[e]) {
        const uint32_t g = ent_gid[e];
        const uint32_t gl = grp_head[g + 1] - grp_head[g];
        const uint64_t cls = gl <= short_max ? 0ull : (gl > huge_min ? 2ull : 1ull);
        fkeys[excl[e]] = ((uint64_t)(post[e] & 0x7FFFFFFFu) << 33) | (cls << 31) | e;
    }
}

// forward lists from the gene-sorted keys; bit 31 of the length flags an own multiplicity > 1 (then fwd_cnt is read)
__global__ void __launch_bounds__(256) fwd_fill_kernel(const uint64_t* __restrict__ fkeys, uint32_t R,
                                                        const uint32_t* __restrict__ post_cnt, const uint32_t* __restrict__ ent_gid,
                                                        const uint32_t* __restrict__ grp_head, uint2* __restrict__ fwd,
                                                        uint32_t* __restrict__ fwd_cnt) {
    const uint32_t j = blockIdx.x * 256u + threadIdx.x;
    if (j >= R) return;
    const uint32_t e = (uint32_t)fkeys[j] & 0x7FFFFFFFu;
    const uint32_t g = ent_gid[e];
    const uint32_t gs = grp_head[g];
    const uint32_t cnt = post_cnt[e];
    fwd[j] = make_uint2(gs, (grp_head[g + 1] - gs) | (cnt > 1 ? 0x80000000u : 0u));
    fwd_cnt[j] = cnt;
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
