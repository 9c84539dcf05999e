// Scoring kernels: the device counterpart of computeScores (reference ig/native/library.cpp:409-527).
//
// One CTA per row gene (persistent CTAs, rows handed out through a global cursor).  The reference's S-sized dense
// accumulators with colour stamps (library.cpp:417-428,467-473) become an open-addressing hash table in shared
// memory keyed by column gene; every posting visited is one coalesced 8-B HBM read and shared-memory atomics only.
//
//   accumulate   inter[c] += min(n, m); pc[c] += m; tc[c] += n          library.cpp:461-479
//                UNIT rows (no multiplicity > 1 anywhere in the row's lists, known from the index): all three sums
//                equal the number of shared k-mers, so ONE counter and ONE atomic per posting.
//   finalize     union, perc, tr_perc, validity, float32 Jaccard          library.cpp:493-505
//   emit         cells with score > 0, SoA in the layout of Scores.java   library.cpp:506-512, 554-575
//   best hits    BH[row][genome(col)] and colmax[col] by atomic max        library.cpp:513-515
//
// Rows whose distinct-column count overflows the table are appended to an overflow list and re-run with a larger
// table; the last resort is score_rows_dense_kernel (global S-sized accumulators, as the reference).
#pragma once

#include "pd_rt.h"

namespace pd {
namespace sk {

static const int kScoreThreads = 256;
static const uint32_t kEmpty = 0xFFFFFFFFu;
static const uint32_t kMaxProbe = 192;

struct ScoreArgs {
    // index
    const uint2* post;
    const uint2* fwd;
    const uint32_t* fwd_cnt;
    const uint32_t* fwd_ptr;
    const uint2* meta;  // (kseq_len, genome)
    // work: (gene, best-hit row) pairs handed out through *cursor; n_rows read from *n_rows_ptr when non-null
    const uint2* rows;
    uint32_t n_rows;
    const unsigned long long* n_rows_ptr;
    uint32_t* cursor;
    // parameters
    uint32_t G;
    float thr;       // 1.0f / (2.0f * (float)k), library.cpp:499
    uint32_t slots;  // hash table slots
    uint32_t gshift; // log2(lanes cooperating on one posting list)
    // outputs
    float* o_score;
    float* o_perc;
    float* o_trperc;
    int32_t* o_row;
    int32_t* o_col;
    int32_t* o_g1;
    int32_t* o_g2;
    unsigned long long cell_cap;
    unsigned long long* n_cells;     // running cell count (keeps counting past cell_cap)
    unsigned long long* n_pairs;     // candidate cells evaluated (col != row)
    uint32_t* bh;                    // float bits, [bh_row * G + genome]
    uint32_t* colmax;                // float bits, [S] (may be null)
    uint2* overflow_rows;            // rows that did not fit `slots`
    unsigned long long* n_overflow;
};

struct DenseArgs {
    uint32_t S;
    uint32_t* acc;      // per CTA: inter[S], pc[S], tc[S], touched[S]; inter/pc/tc all zero between rows
};

// shared finalize: one candidate cell (row r, column c) with its three integer sums
struct RowCtx {
    uint32_t r, bh_row, kr, gr;
};

__device__ __forceinline__ bool finalize_cell(const ScoreArgs& a, const RowCtx& rc, uint32_t c, uint32_t inter, uint32_t pc,
                                              uint32_t tc, float* score, float* perc, float* tr_perc, uint32_t* gc) {
    const uint2 mc = a.meta[c];
    const int uni = (int)rc.kr + (int)mc.x - (int)inter;                       // library.cpp:494-496
    const float p = __fdiv_rn(__int2float_rn((int)pc), __int2float_rn((int)rc.kr));   // :497
    const float t = __fdiv_rn(__int2float_rn((int)tc), __int2float_rn((int)mc.x));    // :498
    const bool valid = (p >= a.thr) || (t >= a.thr);                           // :500
    const float q = __fdiv_rn(__int2float_rn((int)inter), __int2float_rn(uni));
    const float s = valid ? q : 0.0f;                                          // q * 1.0f == q, q * 0.0f == +0 (:501-502)
    *score = s;
    *perc = p;
    *tr_perc = t;
    *gc = mc.y;
    return s > 0.0f;                                                           // :505
}

// warp-aggregated append of the cells wanted by the lanes of a fully converged warp
__device__ __forceinline__ void emit_cells(const ScoreArgs& a, const RowCtx& rc, bool want, uint32_t c, float score, float perc,
                                           float tr_perc, uint32_t gc) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return;
    const unsigned leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(a.n_cells, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) {
        const unsigned long long idx = base + __popc(m & ((1u << lane) - 1u));
        if (idx < a.cell_cap) {
            a.o_score[idx] = score;
            a.o_perc[idx] = perc;
            a.o_trperc[idx] = tr_perc;
            a.o_row[idx] = (int32_t)rc.r;
            a.o_col[idx] = (int32_t)c;
            a.o_g1[idx] = (int32_t)rc.gr;
            a.o_g2[idx] = (int32_t)gc;
        }
        // scores are positive floats: their bit patterns order like the values
        atomicMax(&a.bh[(size_t)rc.bh_row * a.G + gc], __float_as_uint(score));
        if (a.colmax) atomicMax(&a.colmax[c], __float_as_uint(score));
    }
}

__device__ __forceinline__ uint32_t find_slot(uint32_t* keys, uint32_t c, uint32_t slots, volatile int* s_over) {
    uint32_t h = __umulhi(c * 0x9E3779B1u, slots);
#pragma unroll 1
    for (uint32_t probe = 0; probe < kMaxProbe; probe++) {
        const uint32_t k = *(volatile uint32_t*)(keys + h);
        if (k == c) return h;
        if (k == kEmpty) {
            const uint32_t old = atomicCAS(keys + h, kEmpty, c);
            if (old == kEmpty || old == c) return h;
        }
        h++;
        if (h == slots) h = 0;
    }
    *s_over = 1;
    return kEmpty;
}

// MULTI = false: UNIT rows, table = keys[slots] + cnt[slots]            ( 8 B per slot)
// MULTI = true : general rows, table = keys + inter + pc + tc planes     (16 B per slot)
template <bool MULTI>
__global__ void __launch_bounds__(kScoreThreads) score_rows_kernel(ScoreArgs a) {
    PD_DYNAMIC_SMEM(smem_raw);
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* v0 = keys + a.slots;   // UNIT: shared k-mer count; MULTI: inter
    uint32_t* v1 = v0 + a.slots;     // MULTI: pc
    uint32_t* v2 = v1 + a.slots;     // MULTI: tc
    __shared__ uint32_t s_row;
    __shared__ int s_over;

    const unsigned tid = threadIdx.x;
    const uint32_t slots = a.slots;
    const uint32_t n_rows = a.n_rows_ptr ? (uint32_t)*a.n_rows_ptr : a.n_rows;
    const unsigned gl_lanes = 1u << a.gshift;
    const unsigned grp = tid >> a.gshift, lane_in_grp = tid & (gl_lanes - 1u), n_grp = kScoreThreads >> a.gshift;

    for (uint32_t i = tid; i < slots; i += kScoreThreads) {
        keys[i] = kEmpty;
        v0[i] = 0;
        if (MULTI) {
            v1[i] = 0;
            v2[i] = 0;
        }
    }
    unsigned long long pairs = 0;

    for (;;) {
        if (tid == 0) {
            s_row = atomicAdd(a.cursor, 1u);
            s_over = 0;
        }
        __syncthreads();
        const uint32_t ri = s_row;
        if (ri >= n_rows) break;
        const uint2 rw = a.rows[ri];
        RowCtx rc;
        rc.r = rw.x;
        rc.bh_row = rw.y;
        const uint2 mr = a.meta[rc.r];
        rc.kr = mr.x;
        rc.gr = mr.y;
        const uint32_t fb = a.fwd_ptr[rc.r], fe = a.fwd_ptr[rc.r + 1];

        // ---- accumulate along the posting lists of the row's shared k-mers
        for (uint32_t f = fb + grp; f < fe; f += n_grp) {
            if (*(volatile int*)&s_over) break;
            const uint2 fw = a.fwd[f];
            const uint32_t m = MULTI ? a.fwd_cnt[f] : 1u;
            const uint2* pl = a.post + fw.x;
            for (uint32_t p0 = lane_in_grp; p0 < fw.y; p0 += 4 * gl_lanes) {
                // up to four independent 8-B loads in flight per lane
                uint2 e[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t p = p0 + u * gl_lanes;
                    e[u] = (p < fw.y) ? pl[p] : make_uint2(kEmpty, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (e[u].x == kEmpty) continue;
                    const uint32_t h = find_slot(keys, e[u].x, slots, &s_over);
                    if (h == kEmpty) continue;
                    if (MULTI) {
                        atomicAdd(&v0[h], e[u].y < m ? e[u].y : m);
                        atomicAdd(&v1[h], m);
                        atomicAdd(&v2[h], e[u].y);
                    } else {
                        atomicAdd(&v0[h], 1u);
                    }
                }
            }
        }
        __syncthreads();

        if (s_over) {
            // does not fit: hand the row to the next level, wipe the table
            if (tid == 0) {
                const unsigned long long o = atomicAdd(a.n_overflow, 1ull);
                a.overflow_rows[o] = rw;
            }
            for (uint32_t i = tid; i < slots; i += kScoreThreads) {
                keys[i] = kEmpty;
                v0[i] = 0;
                if (MULTI) {
                    v1[i] = 0;
                    v2[i] = 0;
                }
            }
            __syncthreads();
            continue;
        }

        // ---- finalize + emit: every thread walks the same number of slots so warps stay converged
        for (uint32_t b = 0; b < slots; b += kScoreThreads) {
            const uint32_t h = b + tid;
            bool want = false;
            uint32_t c = kEmpty, gc = 0;
            float score = 0.f, perc = 0.f, trp = 0.f;
            if (h < slots) {
                c = keys[h];
                if (c != kEmpty) {
                    const uint32_t inter = v0[h];
                    const uint32_t pc = MULTI ? v1[h] : inter;
                    const uint32_t tc = MULTI ? v2[h] : inter;
                    keys[h] = kEmpty;
                    v0[h] = 0;
                    if (MULTI) {
                        v1[h] = 0;
                        v2[h] = 0;
                    }
                    if (c != rc.r) {  // identity cell dropped (library.cpp:485-487)
                        pairs++;
                        want = finalize_cell(a, rc, c, inter, pc, tc, &score, &perc, &trp, &gc);
                    }
                }
            }
            emit_cells(a, rc, want, c, score, perc, trp, gc);
        }
        __syncthreads();
    }

    // one atomic per warp for the pair statistic
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, d);
    if ((tid & 31) == 0 && pairs) atomicAdd(a.n_pairs, pairs);
}

// Last resort for rows with more distinct columns than any shared-memory table holds: the reference's own scheme,
// S-sized accumulators in global memory (zero between rows) plus a touched list.
__global__ void __launch_bounds__(kScoreThreads) score_rows_dense_kernel(ScoreArgs a, DenseArgs d) {
    __shared__ uint32_t s_row;
    __shared__ uint32_t s_touched;
    const unsigned tid = threadIdx.x;
    const uint32_t n_rows = a.n_rows_ptr ? (uint32_t)*a.n_rows_ptr : a.n_rows;
    uint32_t* inter = d.acc + (size_t)blockIdx.x * 4 * d.S;
    uint32_t* pcv = inter + d.S;
    uint32_t* tcv = pcv + d.S;
    uint32_t* touched = tcv + d.S;
    const unsigned gl_lanes = 1u << a.gshift;
    const unsigned grp = tid >> a.gshift, lane_in_grp = tid & (gl_lanes - 1u), n_grp = kScoreThreads >> a.gshift;
    unsigned long long pairs = 0;

    for (;;) {
        if (tid == 0) {
            s_row = atomicAdd(a.cursor, 1u);
            s_touched = 0;
        }
        __syncthreads();
        const uint32_t ri = s_row;
        if (ri >= n_rows) break;
        const uint2 rw = a.rows[ri];
        RowCtx rc;
        rc.r = rw.x;
        rc.bh_row = rw.y;
        const uint2 mr = a.meta[rc.r];
        rc.kr = mr.x;
        rc.gr = mr.y;
        const uint32_t fb = a.fwd_ptr[rc.r], fe = a.fwd_ptr[rc.r + 1];
        for (uint32_t f = fb + grp; f < fe; f += n_grp) {
            const uint2 fw = a.fwd[f];
            const uint32_t m = a.fwd_cnt[f];
            for (uint32_t p = lane_in_grp; p < fw.y; p += gl_lanes) {
                const uint2 e = a.post[fw.x + p];
                const uint32_t old = atomicAdd(&tcv[e.x], e.y);  // counts are >= 1: old == 0 <=> first touch
                if (old == 0) touched[atomicAdd(&s_touched, 1u)] = e.x;
                atomicAdd(&inter[e.x], e.y < m ? e.y : m);
                atomicAdd(&pcv[e.x], m);
            }
        }
        __threadfence();
        __syncthreads();
        const uint32_t nt = s_touched;
        for (uint32_t b = 0; b < nt; b += kScoreThreads) {
            const uint32_t t = b + tid;
            bool want = false;
            uint32_t c = kEmpty, gc = 0;
            float score = 0.f, perc = 0.f, trp = 0.f;
            if (t < nt) {
                c = *(volatile uint32_t*)&touched[t];
                const uint32_t in = *(volatile uint32_t*)&inter[c], pc = *(volatile uint32_t*)&pcv[c], tc = *(volatile uint32_t*)&tcv[c];
                inter[c] = 0;
                pcv[c] = 0;
                tcv[c] = 0;
                if (c != rc.r) {
                    pairs++;
                    want = finalize_cell(a, rc, c, in, pc, tc, &score, &perc, &trp, &gc);
                }
            }
            emit_cells(a, rc, want, c, score, perc, trp, gc);
        }
        __threadfence();
        __syncthreads();
    }
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, dd);
    if ((tid & 31) == 0 && pairs) atomicAdd(a.n_pairs, pairs);
}

}  // namespace sk
}  // namespace pd
