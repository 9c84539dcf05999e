// The per-genome task of the reference's Java host on the device (reference ig/infoasys/cli/pangenes/Pangenes.java:98-176),
// run over the cells one computeScores(genome) call left in HBM, so that only the network edges travel to the host
// instead of every non-zero cell.  All comparisons are the reference's float comparisons;
// scores are non-negative floats, whose bit patterns order like the values, so maxima / minima are integer atomics.
//
//   inter_mark_kernel      cells (r, c) with genome(c) != g:  bidirectional best hit iff
//                            score == BH[r][genome(c)]  and  score == colmax_g[c]              Pangenes.java:99-104
//                          and per other genome h the largest such score below 1.0             Pangenes.java:116-118
//   row_threshold_kernel   rowthr[r] = min over r's BBH partners c of imax[genome(c)], +inf if none
//                                                                                             Pangenes.java:146-155
//   edge_emit_kernel       the BBH cells, plus cells (r < c, same genome g) with
//                            score == BH[r][g]  and  score == BH[c][g]  and  score >= rowthr[r] Pangenes.java:164-175
#pragma once

#include "pd_rt.h"

namespace pd {
namespace fk {

struct FilterArgs {
    unsigned long long cells;
    const float* score;
    const int32_t* row;     // gene id
    const int32_t* col;     // gene id
    const int32_t* bhrow;   // the row's index inside the genome (= scoresMaxMappings[row])
    const int32_t* g2;      // genome of the column
    uint32_t genome;
    uint32_t G;
    const uint32_t* bh;        // float bits, [rows x G]
    const uint32_t* colmax;    // float bits, [S]
    const uint32_t* local_of;  // gene -> index inside its genome
    uint32_t* imax;            // float bits, [G], zero at entry
    uint32_t* rowthr;          // float bits, [rows], +inf at entry
    uint8_t* flag;             // [cells]
    uint32_t* e_src;
    uint32_t* e_dst;
    float* e_score;
    unsigned long long edge_cap;
    unsigned long long* n_edges;  // keeps counting past edge_cap
};

__global__ void __launch_bounds__(256) inter_mark_kernel(FilterArgs a) {
    const unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.cells) return;
    const uint32_t h = (uint32_t)a.g2[i];
    uint8_t f = 0;
    if (h != a.genome) {
        const float s = a.score[i];
        const uint32_t sb = __float_as_uint(s);
        if (sb == a.bh[(size_t)a.bhrow[i] * a.G + h] && sb == a.colmax[a.col[i]]) {
            f = 1;
            if (s < 1.0f) atomicMax(&a.imax[h], sb);
        }
    }
    a.flag[i] = f;
}

__global__ void __launch_bounds__(256) row_threshold_kernel(FilterArgs a) {
    const unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.cells || !a.flag[i]) return;
    atomicMin(&a.rowthr[a.bhrow[i]], a.imax[a.g2[i]]);
}

__global__ void __launch_bounds__(256) edge_emit_kernel(FilterArgs a) {
    const unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    bool want = false;
    uint32_t r = 0, c = 0;
    float s = 0.f;
    if (i < a.cells) {
        r = (uint32_t)a.row[i];
        c = (uint32_t)a.col[i];
        s = a.score[i];
        if (a.flag[i]) {
            want = true;
        } else if ((uint32_t)a.g2[i] == a.genome && r < c) {
            const uint32_t sb = __float_as_uint(s);
            const uint32_t br = (uint32_t)a.bhrow[i];
            want = sb == a.bh[(size_t)br * a.G + a.genome] && sb == a.bh[(size_t)a.local_of[c] * a.G + a.genome] &&
                   s >= __uint_as_float(a.rowthr[br]);
        }
    }
    const unsigned lane = threadIdx.x & 31;
    const unsigned mb = __ballot_sync(0xffffffffu, want);
    if (mb) {
        const unsigned leader = __ffs((int)mb) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(a.n_edges, (unsigned long long)__popc(mb));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (want) {
            const unsigned long long idx = base + __popc(mb & ((1u << lane) - 1u));
            if (idx < a.edge_cap) {
                a.e_src[idx] = r;
                a.e_dst[idx] = c;
                a.e_score[idx] = s;
            }
        }
    }
}

}  // namespace fk
}  // namespace pd
