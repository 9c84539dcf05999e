// Scoring kernels: the device counterpart of computeScores (reference ig/native/library.cpp:409-527).
//
// One CTA per row gene (persistent CTAs, rows handed out through a global cursor, next row's descriptor prefetched
// behind the current row's work).  The reference's S-sized dense accumulators with colour stamps
// (library.cpp:417-428,467-473) become an open-addressing hash table in shared memory keyed by column gene.
//
//   accumulate   inter[c] += min(n, m); pc[c] += m; tc[c] += n                          library.cpp:461-479
//                Written as (1,1,1) + (min(n,m)-1, m-1, n-1): the first part is ONE shared-memory counter per
//                column (one atomic per posting visited); the second part is non-zero only where a k-mer repeats
//                inside a gene (U/N > 0.999: rare) and goes to a small per-CTA side table in global memory (L2).
//   postings     read through a warp-level load-balanced walk: short posting lists of a batch of 32 forward
//                entries are flattened so that every lane has one posting per step (coalesced inside each list);
//                long lists are walked by the whole warp, 256 B per step.
//   finalize     only the columns touched (a list of table slots), integer validity gate, float32 Jaccard for the
//                cells that pass                                                          library.cpp:493-505
//   emit         cells with score > 0, SoA in the layout of Scores.java                   library.cpp:506-512, 554-575
//   best hits    BH[row][genome(col)] and colmax[col] by atomic max                        library.cpp:513-515
//
// Rows whose distinct-column count overflows the table are appended to an overflow list and re-run with a larger
// table; the last resort is score_rows_dense_kernel (global S-sized accumulators, as the reference).
#pragma once

#include "pd_rt.h"

namespace pd {
namespace sk {

static const int kScoreThreads = 256;
static const int kScoreWarps = kScoreThreads / 32;
static const uint32_t kEmpty = 0xFFFFFFFFu;
static const uint32_t kFlag = 0x80000000u;     // "this column has side-table corrections"
static const uint32_t kShortList = 64;          // lists up to this length are flattened
static const uint32_t kXSlots = 2048;           // side-table slots per CTA (global memory)
static const uint32_t kXCap = kXSlots * 3 / 4;
static const uint32_t kFwdMulti = 0x80000000u;  // bit 31 of a forward entry's length: own multiplicity > 1

struct __align__(16) RowDesc {  // 32 B, built on the host per call
    uint32_t gene, bh_row, fb, fe, kr, gr, fm, pad1;  // forward entries [fb, fm) short lists, [fm, fe) long lists
};

struct ScoreArgs {
    // index
    const uint2* post;
    const uint2* fwd;          // (group start, group length | kFwdMulti)
    const uint32_t* fwd_cnt;
    const uint2* meta;         // (kseq_len, genome)
    // work
    const RowDesc* rows;
    uint32_t n_rows;
    uint32_t* cursor;
    // parameters
    uint32_t G;
    float thr;        // 1.0f / (2.0f * (float)k), library.cpp:499
    uint32_t k2;      // 2k: perc >= thr  <=>  2k * pc >= K  (exact for K < 2^20, see finalize_cell)
    uint32_t slots;   // hash table slots
    uint32_t cap;     // touched-list capacity = max distinct columns accepted (<= 3/4 slots)
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
    RowDesc* overflow_rows;          // rows that did not fit
    unsigned long long* n_overflow;
    // side tables: per CTA kXSlots x (key, d_inter, d_pc, d_tc) + kXSlots touched indices, clean between rows
    uint32_t* xtab;
};

struct DenseArgs {
    uint32_t S;
    uint32_t* acc;  // per CTA: inter[S], pc[S], tc[S], touched[S]; inter/pc/tc all zero between rows
};

struct RowCtx {
    uint32_t r, bh_row, kr, gr;
};

// One candidate cell.  Validity gate (library.cpp:497-500) in integers:
//   (float)pc / (float)K >= 1.0f / (2.0f * (float)k)   <=>   2k * pc >= K        for K < 2^20
// (=>) division and reciprocal are correctly rounded and rounding is monotonic; (<=) if 2k*pc <= K - 1 the two reals
// differ by more than 2^-20 relative, far more than the two roundings (2^-24 each) can close.
// Only cells that pass pay for the three float divisions.
__device__ __forceinline__ bool finalize_cell(const ScoreArgs& a, const RowCtx& rc, uint32_t c, uint32_t inter, uint32_t pc,
                                              uint32_t tc, float* score, float* perc, float* tr_perc, uint32_t* gc) {
    const uint2 mc = a.meta[c];
    *gc = mc.y;
    const bool valid = (a.k2 * pc >= rc.kr) || (a.k2 * tc >= mc.x);
    if (!valid || inter == 0) return false;
    const int uni = (int)rc.kr + (int)mc.x - (int)inter;                              // library.cpp:494-496
    *perc = __fdiv_rn(__int2float_rn((int)pc), __int2float_rn((int)rc.kr));            // :497
    *tr_perc = __fdiv_rn(__int2float_rn((int)tc), __int2float_rn((int)mc.x));          // :498
    const float s = __fdiv_rn(__int2float_rn((int)inter), __int2float_rn(uni));        // :501 (valid -> * 1.0f)
    *score = s;
    return s > 0.0f;                                                                   // :505
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

// Per-warp scratch of the load-balanced walk
struct WarpScratch {
    uint32_t bits[kShortList + 1];  // 32 lists x kShortList postings = 2048 marks (+1: the walk reads one word ahead)
    uint32_t gs[32];
    uint32_t pre[32];
    uint32_t m[32];
};

// Visits every posting (column gene c, its count n, the row's own count m) of the row's shared k-mers from the lane
// that read it: fast(c, n, m) first — the common, branch-light case — and slow(c, n, m) for the postings fast()
// declined, gathered per round so that the divergent code runs once per round instead of once per posting.
//   phase A  forward entries [fb, fm): short posting lists, flattened in batches of 32 entries handed out through
//            ctr[0]: every lane has one posting per step, coalesced inside each list
//   phase B  forward entries [fm, fe): long posting lists, handed out one list at a time through ctr[1]; the warp
//            walks the list 4 x 32 consecutive postings (1 KB) per round
// ctr[0..1] are shared counters, zero at entry.  `stop` is polled warp-uniformly.
template <class Fast, class Slow>
__device__ __forceinline__ void for_each_posting(const ScoreArgs& a, uint32_t fb, uint32_t fm, uint32_t fe, WarpScratch* ws_all,
                                                 uint32_t* ctr, volatile int* stop, Fast fast, Slow slow) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = ws_all[warp];
    const unsigned lt = (1u << lane) - 1u;
    // ---- phase A
    for (;;) {
        uint32_t bi = 0;
        if (lane == 0) bi = atomicAdd(ctr, 1u);
        bi = __shfl_sync(0xffffffffu, bi, 0);
        const uint32_t b0 = fb + bi * 32;
        if (b0 >= fm) break;
        if (__any_sync(0xffffffffu, *stop != 0)) break;
        const uint32_t f = b0 + lane;
        const bool has = f < fm;
        uint2 fw = has ? a.fwd[f] : make_uint2(0u, 0u);
        uint32_t m = 1;
        if (fw.y & kFwdMulti) {
            fw.y &= ~kFwdMulti;
            m = a.fwd_cnt[f];
        }
        uint32_t incl = fw.y;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (unsigned)d) incl += o;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t pre = incl - fw.y;
        ws.bits[lane] = 0;
        ws.bits[lane + 32] = 0;
        ws.gs[lane] = fw.x;
        ws.pre[lane] = pre;
        ws.m[lane] = m;
        __syncwarp();
        if (has) atomicOr(&ws.bits[pre >> 5], 1u << (pre & 31));  // lists are >= 2 long: one mark per list
        __syncwarp();
        uint32_t seen = 0;  // marks before this round
        for (uint32_t t0 = 0; t0 < total; t0 += 64) {
            // two steps per round: both postings are in flight before either is consumed
            const uint32_t w0 = ws.bits[t0 >> 5], w1 = ws.bits[(t0 >> 5) + 1];
            const uint32_t ta = t0 + lane, tb = ta + 32;
            const uint32_t oa = (seen + __popc(w0 & (lt | (1u << lane))) - 1u) & 31u;
            seen += __popc(w0);
            const uint32_t ob = (seen + __popc(w1 & (lt | (1u << lane))) - 1u) & 31u;
            seen += __popc(w1);
            uint2 ea = make_uint2(kEmpty, 0u), eb = make_uint2(kEmpty, 0u);
            uint32_t ma = 1, mb = 1;
            if (ta < total) {
                ea = a.post[ws.gs[oa] + (ta - ws.pre[oa])];
                ma = ws.m[oa];
            }
            if (tb < total) {
                eb = a.post[ws.gs[ob] + (tb - ws.pre[ob])];
                mb = ws.m[ob];
            }
            const bool sa = ea.x != kEmpty && !fast(ea.x, ea.y, ma);
            const bool sb = eb.x != kEmpty && !fast(eb.x, eb.y, mb);
            if (sa | sb) {
                if (sa) slow(ea.x, ea.y, ma);
                if (sb) slow(eb.x, eb.y, mb);
            }
        }
        __syncwarp();
    }
    // ---- phase B
    for (;;) {
        uint32_t li = 0;
        if (lane == 0) li = atomicAdd(ctr + 1, 1u);
        li = __shfl_sync(0xffffffffu, li, 0);
        const uint32_t f = fm + li;
        if (f >= fe) break;
        if (__any_sync(0xffffffffu, *stop != 0)) break;
        uint2 fw = a.fwd[f];
        uint32_t mj = 1;
        if (fw.y & kFwdMulti) {
            fw.y &= ~kFwdMulti;
            mj = a.fwd_cnt[f];
        }
        const uint2* pl = a.post + fw.x;
        const uint32_t gl = fw.y;
        for (uint32_t p0 = lane; p0 < gl; p0 += 128) {
            uint2 e[4];
#pragma unroll
            for (int u = 0; u < 4; u++) e[u] = (p0 + 32 * u < gl) ? pl[p0 + 32 * u] : make_uint2(kEmpty, 0u);
            unsigned miss = 0;
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (e[u].x != kEmpty && !fast(e[u].x, e[u].y, mj)) miss |= 1u << u;
            if (miss) {
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (miss & (1u << u)) slow(e[u].x, e[u].y, mj);
            }
            if ((p0 & 0xF80u) == 0xF80u && __any_sync(0xffffffffu, *stop != 0)) break;  // every 32 rounds
        }
    }
}

__device__ __forceinline__ uint32_t x_find_or_insert(uint32_t* xkeys, uint32_t* xtouched, uint32_t c, uint32_t* s_nx,
                                                     volatile int* s_over) {
    uint32_t h = __umulhi(c * 0x85EBCA6Bu, kXSlots);
    for (uint32_t probe = 0; probe < kXSlots; probe++) {
        const uint32_t k = *(volatile uint32_t*)(xkeys + h);
        if (k == c) return h;
        if (k == kEmpty) {
            const uint32_t old = atomicCAS(xkeys + h, kEmpty, c);
            if (old == kEmpty) {
                const uint32_t xi = atomicAdd(s_nx, 1u);
                if (xi < kXCap) xtouched[xi] = h;
                else *s_over = 1;
                return h;
            }
            if (old == c) return h;
        }
        h = (h + 1 == kXSlots) ? 0 : h + 1;
    }
    *s_over = 1;
    return kEmpty;
}

__global__ void __launch_bounds__(kScoreThreads) score_rows_kernel(ScoreArgs a) {
    PD_DYNAMIC_SMEM(smem_raw);
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* cnt = keys + a.slots;
    uint16_t* touched = reinterpret_cast<uint16_t*>(cnt + a.slots);
    __shared__ WarpScratch ws[kScoreWarps];
    __shared__ uint4 s_desc[2][2];  // two RowDesc buffers
    // per-row control words, double-buffered like s_desc: a row uses [buf]; thread 0 clears [buf ^ 1] at the end of
    // the row, which nobody reads before the closing barrier
    __shared__ uint32_t s_nt2[2], s_nx2[2];
    __shared__ int s_over2[2];
    __shared__ uint32_t s_batch2[2][2];

    const unsigned tid = threadIdx.x;
    const uint32_t slots = a.slots, cap = a.cap;
    uint32_t* xkeys = a.xtab + (size_t)blockIdx.x * (5 * kXSlots);
    uint32_t* xv0 = xkeys + kXSlots;
    uint32_t* xv1 = xv0 + kXSlots;
    uint32_t* xv2 = xv1 + kXSlots;
    uint32_t* xtouched = xv2 + kXSlots;

    for (uint32_t i = tid; i < slots; i += kScoreThreads) {
        keys[i] = kEmpty;
        cnt[i] = 0;
    }
    const uint4* rows4 = reinterpret_cast<const uint4*>(a.rows);
    if (tid == 0) {
        const uint32_t ri = atomicAdd(a.cursor, 1u);
        uint4 d0 = make_uint4(kEmpty, 0u, 0u, 0u), d1 = make_uint4(0u, 0u, 0u, 0u);
        if (ri < a.n_rows) {
            d0 = rows4[2 * (size_t)ri];
            d1 = rows4[2 * (size_t)ri + 1];
        }
        s_desc[0][0] = d0;
        s_desc[0][1] = d1;
        s_nt2[0] = s_nt2[1] = 0;
        s_nx2[0] = s_nx2[1] = 0;
        s_over2[0] = s_over2[1] = 0;
        s_batch2[0][0] = s_batch2[0][1] = s_batch2[1][0] = s_batch2[1][1] = 0;
    }
    __syncthreads();
    unsigned long long pairs = 0;
    int buf = 0;

    for (;;) {
        const uint4 rw0 = s_desc[buf][0], rw1 = s_desc[buf][1];  // (gene, bh_row, fb, fe), (kr, gr, -, -)
        if (rw0.x == kEmpty) break;
        uint32_t& s_nt = s_nt2[buf];
        uint32_t& s_nx = s_nx2[buf];
        int& s_over = s_over2[buf];
        uint32_t* s_batch = s_batch2[buf];
        uint32_t next_ri = 0;
        if (tid == 0) next_ri = atomicAdd(a.cursor, 1u);  // consumed after the accumulate phase
        RowCtx rc;
        rc.r = rw0.x;
        rc.bh_row = rw0.y;
        rc.kr = rw1.x;
        rc.gr = rw1.y;

        // ---- accumulate
        for_each_posting(
            a, rw0.z, rw1.z, rw0.w, ws, s_batch, &s_over,
            // fast: the column is already in its home slot and no k-mer repeats: one load, one atomic
            [&](uint32_t c, uint32_t n, uint32_t m) -> bool {
                const uint32_t h = __umulhi(c * 0x9E3779B1u, slots);
                if ((n | m) > 1u || *(volatile uint32_t*)(keys + h) != c) return false;
                atomicAdd(&cnt[h], 1u);
                return true;
            },
            // slow: first visit of the column in this row, a displaced key, or a repeated k-mer
            [&](uint32_t c, uint32_t n, uint32_t m) {
                uint32_t h = __umulhi(c * 0x9E3779B1u, slots);
                uint32_t k = *(volatile uint32_t*)(keys + h);
                uint32_t probes = 0;
                while (k != c) {
                    if (k == kEmpty) {
                        const uint32_t old = atomicCAS(keys + h, kEmpty, c);
                        if (old == kEmpty) {
                            const uint32_t pos = atomicAdd(&s_nt, 1u);
                            if (pos < cap) touched[pos] = (uint16_t)h;
                            else s_over = 1;
                            break;
                        }
                        if (old == c) break;
                    }
                    h = (h + 1 == slots) ? 0 : h + 1;
                    ++probes;
                    if (probes > slots) s_over = 1;  // full table
                    if ((probes & 15u) == 0 && *(volatile int*)&s_over) {  // row already lost
                        h = kEmpty;
                        break;
                    }
                    k = *(volatile uint32_t*)(keys + h);
                }
                if (h != kEmpty) {
                    atomicAdd(&cnt[h], 1u);
                    if ((n | m) > 1u) {  // corrections for the repeated k-mer go to the side table
                        atomicOr(&cnt[h], kFlag);
                        const uint32_t xs = x_find_or_insert(xkeys, xtouched, c, &s_nx, &s_over);
                        if (xs != kEmpty) {
                            const uint32_t mn = n < m ? n : m;
                            if (mn > 1) atomicAdd(&xv0[xs], mn - 1);
                            if (m > 1) atomicAdd(&xv1[xs], m - 1);
                            if (n > 1) atomicAdd(&xv2[xs], n - 1);
                        }
                    }
                }
            });
        __threadfence_block();
        __syncthreads();

        uint4 nd0 = make_uint4(kEmpty, 0u, 0u, 0u), nd1 = make_uint4(0u, 0u, 0u, 0u);
        if (tid == 0 && next_ri < a.n_rows) {  // consumed at the end of the row
            nd0 = rows4[2 * (size_t)next_ri];
            nd1 = rows4[2 * (size_t)next_ri + 1];
        }

        const uint32_t nx = s_nx < kXCap ? s_nx : kXCap;
        if (s_over) {
            // does not fit: hand the row to the next level, wipe both tables
            if (tid == 0) {
                const unsigned long long o = atomicAdd(a.n_overflow, 1ull);
                uint4* dst = reinterpret_cast<uint4*>(a.overflow_rows + o);
                dst[0] = rw0;
                dst[1] = rw1;
            }
            for (uint32_t i = tid; i < slots; i += kScoreThreads) {
                keys[i] = kEmpty;
                cnt[i] = 0;
            }
            if (s_nx) {
                for (uint32_t i = tid; i < kXSlots; i += kScoreThreads) {
                    xkeys[i] = kEmpty;
                    xv0[i] = 0;
                    xv1[i] = 0;
                    xv2[i] = 0;
                }
                __threadfence();
            }
        } else {
            // ---- finalize + emit over the touched columns; all threads run the same number of steps
            if (nx) __threadfence();
            const uint32_t nt = s_nt;
            for (uint32_t b = 0; b < nt; b += kScoreThreads) {
                const uint32_t i = b + tid;
                bool want = false;
                uint32_t c = kEmpty, gc = 0;
                float score = 0.f, perc = 0.f, trp = 0.f;
                if (i < nt) {
                    const uint32_t h = touched[i];
                    c = keys[h];
                    const uint32_t v = cnt[h];
                    keys[h] = kEmpty;
                    cnt[h] = 0;
                    uint32_t inter = v & ~kFlag, pc = inter, tc = inter;
                    if (v & kFlag) {
                        uint32_t xs = __umulhi(c * 0x85EBCA6Bu, kXSlots);
                        while (*(volatile uint32_t*)(xkeys + xs) != c) xs = (xs + 1 == kXSlots) ? 0 : xs + 1;
                        inter += *(volatile uint32_t*)(xv0 + xs);
                        pc += *(volatile uint32_t*)(xv1 + xs);
                        tc += *(volatile uint32_t*)(xv2 + xs);
                    }
                    if (c != rc.r) {  // identity cell dropped (library.cpp:485-487)
                        pairs++;
                        want = finalize_cell(a, rc, c, inter, pc, tc, &score, &perc, &trp, &gc);
                    }
                }
                emit_cells(a, rc, want, c, score, perc, trp, gc);
            }
            if (nx) {  // side-table slots are released only after every reader is done (linear probing)
                __syncthreads();
                for (uint32_t i = tid; i < nx; i += kScoreThreads) {
                    const uint32_t xs = xtouched[i];
                    xkeys[xs] = kEmpty;
                    xv0[xs] = 0;
                    xv1[xs] = 0;
                    xv2[xs] = 0;
                }
                __threadfence();
            }
        }
        if (tid == 0) {
            s_desc[buf ^ 1][0] = nd0;
            s_desc[buf ^ 1][1] = nd1;
            s_nt2[buf ^ 1] = 0;
            s_nx2[buf ^ 1] = 0;
            s_over2[buf ^ 1] = 0;
            s_batch2[buf ^ 1][0] = 0;
            s_batch2[buf ^ 1][1] = 0;
        }
        __syncthreads();
        buf ^= 1;
    }

    // one atomic per warp for the pair statistic
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, d);
    if ((tid & 31) == 0 && pairs) atomicAdd(a.n_pairs, pairs);
}

// Last resort for rows with more distinct columns than any shared-memory table holds: the reference's own scheme,
// S-sized accumulators in global memory (zero between rows) plus a touched list.
__global__ void __launch_bounds__(kScoreThreads) score_rows_dense_kernel(ScoreArgs a, DenseArgs d) {
    __shared__ WarpScratch ws[kScoreWarps];
    __shared__ uint32_t s_row;
    __shared__ uint32_t s_touched;
    __shared__ int s_stop;
    __shared__ uint32_t s_batch[2];
    const unsigned tid = threadIdx.x;
    uint32_t* inter = d.acc + (size_t)blockIdx.x * 4 * d.S;
    uint32_t* pcv = inter + d.S;
    uint32_t* tcv = pcv + d.S;
    uint32_t* touched = tcv + d.S;
    unsigned long long pairs = 0;
    if (tid == 0) s_stop = 0;

    for (;;) {
        if (tid == 0) {
            s_row = atomicAdd(a.cursor, 1u);
            s_touched = 0;
            s_batch[0] = 0;
            s_batch[1] = 0;
        }
        __syncthreads();
        const uint32_t ri = s_row;
        if (ri >= a.n_rows) break;
        const RowDesc rw = a.rows[ri];
        RowCtx rc;
        rc.r = rw.gene;
        rc.bh_row = rw.bh_row;
        rc.kr = rw.kr;
        rc.gr = rw.gr;
        for_each_posting(
            a, rw.fb, rw.fm, rw.fe, ws, s_batch, &s_stop, [](uint32_t, uint32_t, uint32_t) -> bool { return false; },
            [&](uint32_t c, uint32_t n, uint32_t m) {
                const uint32_t old = atomicAdd(&tcv[c], n);  // counts are >= 1: old == 0 <=> first touch
                if (old == 0) touched[atomicAdd(&s_touched, 1u)] = c;
                atomicAdd(&inter[c], n < m ? n : m);
                atomicAdd(&pcv[c], m);
            });
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
