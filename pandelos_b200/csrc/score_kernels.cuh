// Scoring kernels: the device counterpart of computeScores (reference ig/native/library.cpp:409-527).
//
// One CTA per row gene at a time (persistent CTAs, rows handed out through a global cursor; the next row's
// descriptor and forward entries are fetched behind the current row's work).  The reference's S-sized dense
// accumulators with colour stamps (library.cpp:417-428,467-473) become an open-addressing hash table in shared memory
// keyed by column gene.
//
//   accumulate   inter[c] += min(n, m); pc[c] += m; tc[c] += n                          library.cpp:461-479
//                Written as (1,1,1) + (min(n,m)-1, m-1, n-1): the first part is ONE shared-memory counter per
//                column (one atomic per posting visited); the second part is non-zero only where a k-mer repeats
//                inside a gene (U/N > 0.999: rare) and goes to a small per-CTA side table in global memory (L2).
//   postings     4-byte entries (column gene, bit 31 = "this gene holds the k-mer more than once").  The row's
//                forward entries are staged in shared memory (cp.async, double buffered across rows) and split by the
//                index build into three classes:
//                  short lists  (<= kShortList)  flattened in batches of 32 lists so every lane has a posting,
//                  long lists                    one warp walks one list, 4 x 32 consecutive postings per round,
//                  huge lists   (> kHugeList)    all warps of the CTA stride over the same list.
//                Every round first tries the branch-free case for its 4 postings per lane (column already in its
//                home slot: one LDS + one ATOMS), then a per-lane state machine probes / inserts what is left.
//   finalize     warp-local: each warp compacts the occupied slots of its slice of the table, applies the exact
//                integer validity gate, reserves output space with one atomic per warp and row, then computes the
//                float32 Jaccard only for the cells that pass                             library.cpp:493-505
//   emit         cells with score > 0, SoA in the layout of Scores.java                   library.cpp:506-512, 554-575
//   best hits    BH[row][genome(col)] and colmax[col] by atomic max                        library.cpp:513-515
//
// Rows whose columns do not fit the table are appended to an overflow list and re-run with the largest table; the
// last resort is score_rows_dense_kernel (global S-sized accumulators, as the reference).
#pragma once

#include "pd_rt.h"

namespace pd {
namespace sk {

static const uint32_t kEmpty = 0xFFFFFFFFu;
static const uint32_t kFlag = 0x80000000u;      // counter word: "this column has side-table corrections"
static const uint32_t kMulti = 0x80000000u;     // posting / forward length: multiplicity > 1 (then *_cnt is read)
static const uint32_t kShortList = 64;          // lists up to this length are flattened
static const uint32_t kHugeList = 2048;         // lists longer than this are walked by the whole CTA
static const uint32_t kXSlots = 2048;           // side-table slots per CTA (global memory)
static const uint32_t kXCap = kXSlots * 3 / 4;
static const uint32_t kProbeLimit = 160;        // probes after which a row is declared too big for its table
static const int kItems = 4;                    // postings per lane per round
static const int kDenseThreads = 256;

struct __align__(16) RowDesc {  // 32 B, built on the host per call
    uint32_t gene, bh_row, fb, fe, kr, gr, fm, fh;  // forward entries: [fb, fm) short, [fm, fh) long, [fh, fe) huge lists
};

struct ScoreArgs {
    // index
    const uint32_t* post;      // column gene | kMulti
    const uint32_t* post_cnt;  // multiplicity of every posting (read only where kMulti is set)
    const uint2* fwd;          // (group start, group length | kMulti)
    const uint32_t* fwd_cnt;   // the row's own multiplicity (read only where kMulti is set)
    const uint2* meta;         // (kseq_len, genome)
    // work
    const RowDesc* rows;
    uint32_t n_rows;
    uint32_t* cursor;
    // parameters
    uint32_t G;
    uint32_t k2;      // 2k: perc >= thr  <=>  2k * pc >= K  (exact for K < 2^20, see gate())
    uint32_t hbits;   // log2 of the hash table slots
    uint32_t fcap;    // forward entries staged per segment
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

// ------------------------------------------------------------------------------------------------ small helpers

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
#ifdef PD_EMU
    *reinterpret_cast<uint2*>(smem_dst) = *reinterpret_cast<const uint2*>(gsrc);
#else
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifndef PD_EMU
    asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
}

// Validity gate (library.cpp:497-500) in integers:
//   (float)pc / (float)K >= 1.0f / (2.0f * (float)k)   <=>   2k * pc >= K        for K < 2^20
// (=>) division and reciprocal are correctly rounded and rounding is monotonic; (<=) if 2k*pc <= K - 1 the two reals
// differ by more than 2^-20 relative, far more than the two roundings (2^-24 each) can close.
// Only cells that pass pay for the three float divisions.  inter >= 1 for every touched column, so score > 0
// (library.cpp:505) is the same condition.
__device__ __forceinline__ bool gate(uint32_t k2, uint32_t pc, uint32_t tc, uint32_t kr, uint32_t kc) {
    return (k2 * pc >= kr) || (k2 * tc >= kc);
}

struct Tab {
    uint32_t* keys;
    uint32_t* cnt;
    uint32_t mask;
    uint32_t shift;  // 32 - hbits
    uint32_t limit;  // probe limit
    // side table
    uint32_t* xkeys;
    uint32_t* xv0;
    uint32_t* xv1;
    uint32_t* xv2;
    uint32_t* xtouched;
    uint32_t* s_nx;
    volatile int* s_over;
};

__device__ __forceinline__ uint32_t home_slot(const Tab& t, uint32_t c) { return (c * 0x9E3779B1u) >> t.shift; }

__device__ __forceinline__ uint32_t x_find_or_insert(const Tab& t, uint32_t c) {
    uint32_t h = __umulhi(c * 0x85EBCA6Bu, kXSlots);
    for (uint32_t probe = 0; probe < kXSlots; probe++) {
        const uint32_t k = *(volatile uint32_t*)(t.xkeys + h);
        if (k == c) return h;
        if (k == kEmpty) {
            const uint32_t old = atomicCAS(t.xkeys + h, kEmpty, c);
            if (old == kEmpty) {
                const uint32_t xi = atomicAdd(t.s_nx, 1u);
                if (xi < kXCap) t.xtouched[xi] = h;
                else *t.s_over = 1;
                return h;
            }
            if (old == c) return h;
        }
        h = (h + 1 == kXSlots) ? 0 : h + 1;
    }
    *t.s_over = 1;
    return kEmpty;
}

// The general case of one posting: column c held n times by its gene, m times by the row (n | m > 1: rare).
__device__ __noinline__ void add_general(const Tab& t, uint32_t c, uint32_t n, uint32_t m) {
    uint32_t h = home_slot(t, c);
    uint32_t probes = 0;
    for (;;) {
        uint32_t k = *(volatile uint32_t*)(t.keys + h);
        if (k == kEmpty) {
            const uint32_t old = atomicCAS(t.keys + h, kEmpty, c);
            k = (old == kEmpty) ? c : old;
        }
        if (k == c) break;
        h = (h + 1) & t.mask;
        if (++probes > t.limit) {
            *t.s_over = 1;
            return;
        }
    }
    atomicAdd(&t.cnt[h], 1u);
    if ((n | m) > 1u) {  // corrections for the repeated k-mer go to the side table
        atomicOr(&t.cnt[h], kFlag);
        const uint32_t xs = x_find_or_insert(t, c);
        if (xs != kEmpty) {
            const uint32_t mn = n < m ? n : m;
            if (mn > 1) atomicAdd(&t.xv0[xs], mn - 1);
            if (m > 1) atomicAdd(&t.xv1[xs], m - 1);
            if (n > 1) atomicAdd(&t.xv2[xs], n - 1);
        }
    }
}

// kItems postings per lane, all with n = m = 1 (c[u] == kEmpty: no posting).
//   step 1  branch-free: the column already sits in its home slot -> one shared-memory load, one atomic
//   step 2  what is left (first visit of a column in this row, or a displaced key): a per-lane state machine walks
//           its postings one probe per iteration, so lanes with work left never wait for each other's probes
__device__ __forceinline__ void add_ones(const Tab& t, const uint32_t (&c)[kItems]) {
    uint32_t h[kItems], k[kItems];
#pragma unroll
    for (int u = 0; u < kItems; u++) h[u] = home_slot(t, c[u]);
#pragma unroll
    for (int u = 0; u < kItems; u++) k[u] = *(volatile uint32_t*)(t.keys + h[u]);
    unsigned miss = 0;
#pragma unroll
    for (int u = 0; u < kItems; u++) {
        if (c[u] != kEmpty) {
            if (k[u] == c[u]) atomicAdd(&t.cnt[h[u]], 1u);
            else miss |= 1u << u;
        }
    }
    if (miss) {
        uint32_t cur = kEmpty, hh = 0, probes = 0;
        for (;;) {
            if (cur == kEmpty) {
                if (!miss) break;
                const int u = __ffs((int)miss) - 1;
                miss &= miss - 1;
                cur = c[0];
                hh = h[0];
#pragma unroll
                for (int v = 1; v < kItems; v++)
                    if (u == v) {
                        cur = c[v];
                        hh = h[v];
                    }
                probes = 0;
            }
            uint32_t kk = *(volatile uint32_t*)(t.keys + hh);
            if (kk == kEmpty) {
                const uint32_t old = atomicCAS(t.keys + hh, kEmpty, cur);
                kk = (old == kEmpty) ? cur : old;
            }
            if (kk == cur) {
                atomicAdd(&t.cnt[hh], 1u);
                cur = kEmpty;
            } else {
                hh = (hh + 1) & t.mask;
                if (++probes > t.limit) {
                    *t.s_over = 1;
                    cur = kEmpty;
                }
            }
        }
    }
}

// Per-warp scratch of the flattened walk over short lists
struct WarpScratch {
    uint32_t bits[kShortList + kItems];  // 32 lists x kShortList postings = 2048 marks (+ the words a round reads ahead)
    uint32_t pre[32];
};

// One round of a list walk: postings pl[p0 + 32 u + lane], u < kItems, of a list of gl postings whose k-mer the row
// holds mj times; base = index of pl[0] in the posting array (for post_cnt).
__device__ __forceinline__ void list_round(const ScoreArgs& a, const Tab& t, const uint32_t* __restrict__ pl, uint32_t base,
                                           uint32_t p0, uint32_t gl, uint32_t mj) {
    const unsigned lane = threadIdx.x & 31;
    uint32_t e[kItems];
#pragma unroll
    for (int u = 0; u < kItems; u++) {
        const uint32_t p = p0 + 32u * u + lane;
        e[u] = p < gl ? pl[p] : kEmpty;
    }
    // bit 31: a repeated k-mer (or no posting at all); the row's own repeat count is uniform over the list
    uint32_t any = 0;
#pragma unroll
    for (int u = 0; u < kItems; u++) any |= e[u];
    if ((any & kMulti) || mj > 1) {
#pragma unroll
        for (int u = 0; u < kItems; u++) {
            if (e[u] != kEmpty && ((e[u] & kMulti) || mj > 1)) {
                const uint32_t n = (e[u] & kMulti) ? a.post_cnt[base + p0 + 32u * u + lane] : 1u;
                add_general(t, e[u] & ~kMulti, n, mj);
                e[u] = kEmpty;
            }
        }
    }
    add_ones(t, e);
}

// Accumulates the staged forward entries fbuf[0, n_stage) = forward entries [f0, f0 + n_stage) of the row.
// Classes by position: [0, ns) short, [ns, nl) long, [nl, n_stage) huge.  ctr[0..1] are shared work counters, zero
// at entry.  No ordering is needed between the three parts: every update is an atomic on the table.
template <int THREADS>
__device__ __forceinline__ void accumulate(const ScoreArgs& a, const Tab& t, const uint2* fbuf, uint32_t f0, uint32_t ns, uint32_t nl,
                                           uint32_t n_stage, WarpScratch* ws_all, uint32_t* ctr) {
    constexpr int WARPS = THREADS / 32;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = ws_all[warp];
    const unsigned le = 0xffffffffu >> (31 - lane);  // lanes <= mine
    // ---- short lists: batches of 32, flattened
    for (;;) {
        uint32_t bi = 0;
        if (lane == 0) bi = *t.s_over ? 0x03FFFFFFu : atomicAdd(ctr, 1u);  // one lane polls the stop flag: uniform exit
        bi = __shfl_sync(0xffffffffu, bi, 0);
        const uint32_t b0 = bi * 32;
        if (b0 >= ns) break;
        const uint32_t j = b0 + lane;
        const bool has = j < ns;
        const uint32_t len = has ? (fbuf[j].y & ~kMulti) : 0u;
        uint32_t incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (unsigned)d) incl += o;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t pre = incl - len;
        ws.bits[lane] = 0;
        ws.bits[lane + 32] = 0;
        if (lane < kItems) ws.bits[64 + lane] = 0;
        ws.pre[lane] = pre;
        __syncwarp();
        if (has) atomicOr(&ws.bits[pre >> 5], 1u << (pre & 31));  // lists are >= 2 long: one mark per list
        __syncwarp();
        uint32_t seen = 0;  // marks before this round
        for (uint32_t t0 = 0; t0 < total; t0 += 32 * kItems) {
            uint32_t e[kItems], o[kItems];
#pragma unroll
            for (int u = 0; u < kItems; u++) {
                const uint32_t w = ws.bits[(t0 >> 5) + u];
                o[u] = (seen + __popc(w & le) - 1u) & 31u;
                seen += __popc(w);
            }
            bool special = false;
#pragma unroll
            for (int u = 0; u < kItems; u++) {
                const uint32_t tt = t0 + 32u * u + lane;
                e[u] = kEmpty;
                if (tt < total) {
                    const uint2 fw = fbuf[b0 + o[u]];
                    e[u] = a.post[fw.x + (tt - ws.pre[o[u]])];
                    special |= ((fw.y | e[u]) & kMulti) != 0;
                }
            }
            if (special) {
#pragma unroll
                for (int u = 0; u < kItems; u++) {
                    const uint32_t tt = t0 + 32u * u + lane;
                    if (tt < total) {
                        const uint2 fw = fbuf[b0 + o[u]];
                        if ((fw.y | e[u]) & kMulti) {
                            const uint32_t n = (e[u] & kMulti) ? a.post_cnt[fw.x + (tt - ws.pre[o[u]])] : 1u;
                            const uint32_t m = (fw.y & kMulti) ? a.fwd_cnt[f0 + b0 + o[u]] : 1u;
                            add_general(t, e[u] & ~kMulti, n, m);
                            e[u] = kEmpty;
                        }
                    }
                }
            }
            add_ones(t, e);
        }
        __syncwarp();
    }
    // ---- long lists: one warp per list
    for (;;) {
        uint32_t li = 0;
        if (lane == 0) li = *t.s_over ? 0x7FFFFFFFu : atomicAdd(ctr + 1, 1u);
        li = __shfl_sync(0xffffffffu, li, 0);
        if (li >= nl - ns) break;
        const uint32_t j = ns + li;
        const uint2 fw = fbuf[j];
        const uint32_t mj = (fw.y & kMulti) ? a.fwd_cnt[f0 + j] : 1u;
        const uint32_t gl = fw.y & ~kMulti;
        const uint32_t* pl = a.post + fw.x;
        for (uint32_t p0 = 0; p0 < gl; p0 += 32 * kItems) list_round(a, t, pl, fw.x, p0, gl, mj);
    }
    // ---- huge lists: the whole CTA strides over each
    for (uint32_t j = nl; j < n_stage; j++) {
        if (__any_sync(0xffffffffu, *t.s_over != 0)) break;
        const uint2 fw = fbuf[j];
        const uint32_t mj = (fw.y & kMulti) ? a.fwd_cnt[f0 + j] : 1u;
        const uint32_t gl = fw.y & ~kMulti;
        const uint32_t* pl = a.post + fw.x;
        for (uint32_t p0 = warp * (32 * kItems); p0 < gl; p0 += WARPS * 32 * kItems) {
            list_round(a, t, pl, fw.x, p0, gl, mj);
            if ((p0 & 0x3FFFu) < WARPS * 32 * kItems && __any_sync(0xffffffffu, *t.s_over != 0)) break;  // poll now and then
        }
    }
}

// (inter, pc, tc) of a table slot
__device__ __forceinline__ void slot_sums(const Tab& t, uint32_t c, uint32_t v, uint32_t* inter, uint32_t* pc, uint32_t* tc) {
    uint32_t i = v & ~kFlag, p = i, q = i;
    if (v & kFlag) {
        uint32_t xs = __umulhi(c * 0x85EBCA6Bu, kXSlots);
        while (*(volatile uint32_t*)(t.xkeys + xs) != c) xs = (xs + 1 == kXSlots) ? 0 : xs + 1;
        i += *(volatile uint32_t*)(t.xv0 + xs);
        p += *(volatile uint32_t*)(t.xv1 + xs);
        q += *(volatile uint32_t*)(t.xv2 + xs);
    }
    *inter = i;
    *pc = p;
    *tc = q;
}

__device__ __forceinline__ void emit_cell(const ScoreArgs& a, const RowCtx& rc, unsigned long long idx, uint32_t c, uint32_t gc,
                                          uint32_t kc, uint32_t inter, uint32_t pc, uint32_t tc) {
    const int uni = (int)rc.kr + (int)kc - (int)inter;                                       // library.cpp:494-496
    const float perc = __fdiv_rn(__int2float_rn((int)pc), __int2float_rn((int)rc.kr));        // :497
    const float tr_perc = __fdiv_rn(__int2float_rn((int)tc), __int2float_rn((int)kc));        // :498
    const float score = __fdiv_rn(__int2float_rn((int)inter), __int2float_rn(uni));           // :501 (valid -> * 1.0f)
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

// ------------------------------------------------------------------------------------------------ main kernel

template <int THREADS>
__global__ void __launch_bounds__(THREADS) score_rows_kernel(ScoreArgs a) {
    constexpr int WARPS = THREADS / 32;
    PD_DYNAMIC_SMEM(smem_raw);
    const uint32_t H = 1u << a.hbits;
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* cnt = keys + H;
    uint2* fwdbuf = reinterpret_cast<uint2*>(cnt + H);                                  // 2 x fcap
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(fwdbuf + 2 * (size_t)a.fcap);     // WARPS
    uint16_t* cand = reinterpret_cast<uint16_t*>(ws + WARPS);                           // H
    __shared__ uint4 s_desc[2][2];  // two RowDesc buffers
    // per-row control words, double-buffered like s_desc: a row uses [buf]; thread 0 clears [buf ^ 1] while the row
    // runs (its last readers finished before the row's opening barrier)
    __shared__ uint32_t s_nx2[2];
    __shared__ int s_over2[2];
    __shared__ uint32_t s_ctr2[2][2];

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t fcap = a.fcap;
    uint32_t* xbase = a.xtab + (size_t)blockIdx.x * (5 * kXSlots);

    Tab t;
    t.keys = keys;
    t.cnt = cnt;
    t.mask = H - 1;
    t.shift = 32 - a.hbits;
    t.limit = H < kProbeLimit ? H : kProbeLimit;
    t.xkeys = xbase;
    t.xv0 = xbase + kXSlots;
    t.xv1 = xbase + 2 * kXSlots;
    t.xv2 = xbase + 3 * kXSlots;
    t.xtouched = xbase + 4 * kXSlots;

    for (uint32_t i = tid; i < H; i += THREADS) {
        keys[i] = kEmpty;
        cnt[i] = 0;
    }
    const uint4* rows4 = reinterpret_cast<const uint4*>(a.rows);
    uint32_t idx_next = 0;  // thread 0: row claimed for the iteration after this one
    if (tid == 0) {
        const uint32_t ri = atomicAdd(a.cursor, 1u);
        idx_next = atomicAdd(a.cursor, 1u);
        uint4 d0 = make_uint4(kEmpty, 0u, 0u, 0u), d1 = make_uint4(0u, 0u, 0u, 0u);
        if (ri < a.n_rows) {
            d0 = rows4[2 * (size_t)ri];
            d1 = rows4[2 * (size_t)ri + 1];
        }
        s_desc[0][0] = d0;
        s_desc[0][1] = d1;
        s_nx2[0] = s_nx2[1] = 0;
        s_over2[0] = s_over2[1] = 0;
        s_ctr2[0][0] = s_ctr2[0][1] = s_ctr2[1][0] = s_ctr2[1][1] = 0;
    }
    __syncthreads();
    {
        const uint4 d0 = s_desc[0][0];
        if (d0.x != kEmpty) {
            const uint32_t f1 = d0.w - d0.z < fcap ? d0.w : d0.z + fcap;
            for (uint32_t f = d0.z + tid; f < f1; f += THREADS) cp_async8(&fwdbuf[f - d0.z], &a.fwd[f]);
        }
    }
    unsigned long long pairs = 0;
    int buf = 0;

    for (;;) {
        cp_async_wait_all();
        __syncthreads();  // the row's descriptor and first forward segment are staged; the table is clean
        const uint4 rw0 = s_desc[buf][0], rw1 = s_desc[buf][1];  // (gene, bh_row, fb, fe), (kr, gr, fm, fh)
        if (rw0.x == kEmpty) break;
        t.s_nx = &s_nx2[buf];
        t.s_over = &s_over2[buf];
        uint4 nd0 = make_uint4(kEmpty, 0u, 0u, 0u), nd1 = make_uint4(0u, 0u, 0u, 0u);
        uint32_t idx_after = 0;
        if (tid == 0) {  // both are consumed after the accumulate phase
            if (idx_next < a.n_rows) {
                nd0 = rows4[2 * (size_t)idx_next];
                nd1 = rows4[2 * (size_t)idx_next + 1];
            }
            idx_after = atomicAdd(a.cursor, 1u);
        }
        RowCtx rc;
        rc.r = rw0.x;
        rc.bh_row = rw0.y;
        rc.kr = rw1.x;
        rc.gr = rw1.y;
        const uint32_t fb = rw0.z, fe = rw0.w, fm = rw1.z, fh = rw1.w;
        uint2* fbuf = fwdbuf + (size_t)buf * fcap;

        // ---- accumulate, one staged segment of forward entries at a time
        for (uint32_t f0 = fb;;) {
            const uint32_t f1 = fe - f0 < fcap ? fe : f0 + fcap;
            const uint32_t ns = fm > f0 ? (fm < f1 ? fm - f0 : f1 - f0) : 0u;
            const uint32_t nl = fh > f0 ? (fh < f1 ? fh - f0 : f1 - f0) : 0u;
            accumulate<THREADS>(a, t, fbuf, f0, ns, nl, f1 - f0, ws, s_ctr2[buf]);
            if (f1 >= fe) break;
            __syncthreads();
            f0 = f1;
            const uint32_t f2 = fe - f0 < fcap ? fe : f0 + fcap;
            for (uint32_t f = f0 + tid; f < f2; f += THREADS) fbuf[f - f0] = a.fwd[f];
            if (tid == 0) s_ctr2[buf][0] = s_ctr2[buf][1] = 0;
            __syncthreads();
        }
        if (tid == 0) {
            s_desc[buf ^ 1][0] = nd0;
            s_desc[buf ^ 1][1] = nd1;
            s_nx2[buf ^ 1] = 0;
            s_over2[buf ^ 1] = 0;
            s_ctr2[buf ^ 1][0] = 0;
            s_ctr2[buf ^ 1][1] = 0;
            idx_next = idx_after;
        }
        __threadfence_block();
        __syncthreads();  // the table is complete

        {  // stage the next row's first forward segment behind the finalize pass
            const uint4 d0 = s_desc[buf ^ 1][0];
            if (d0.x != kEmpty) {
                uint2* nbuf = fwdbuf + (size_t)(buf ^ 1) * fcap;
                const uint32_t f1 = d0.w - d0.z < fcap ? d0.w : d0.z + fcap;
                for (uint32_t f = d0.z + tid; f < f1; f += THREADS) cp_async8(&nbuf[f - d0.z], &a.fwd[f]);
            }
        }
        const bool over = *t.s_over != 0;
        const uint32_t nx = *t.s_nx < kXCap ? *t.s_nx : kXCap;
        // each warp owns a slice of the table
        const uint32_t spw = H >= 32u * WARPS ? H / WARPS : 32u;
        const uint32_t lo = warp * spw;
        if (over) {
            // does not fit: hand the row to the next level, wipe both tables
            if (tid == 0) {
                const unsigned long long o = atomicAdd(a.n_overflow, 1ull);
                uint4* dst = reinterpret_cast<uint4*>(a.overflow_rows + o);
                dst[0] = rw0;
                dst[1] = rw1;
            }
            for (uint32_t i = tid; i < H; i += THREADS) {
                keys[i] = kEmpty;
                cnt[i] = 0;
            }
            if (*t.s_nx) {
                for (uint32_t i = tid; i < kXSlots; i += THREADS) {
                    t.xkeys[i] = kEmpty;
                    t.xv0[i] = 0;
                    t.xv1[i] = 0;
                    t.xv2[i] = 0;
                }
                __threadfence();
            }
        } else {
            if (nx) __threadfence();
            if (lo < H) {
                uint16_t* cw = cand + lo;
                // ---- occupied slots of the slice, compacted
                uint32_t n = 0;
                for (uint32_t s0 = 0; s0 < spw; s0 += 32) {
                    const uint32_t s = lo + s0 + lane;
                    const bool occ = keys[s] != kEmpty;
                    const unsigned mb = __ballot_sync(0xffffffffu, occ);
                    if (occ) cw[n + __popc(mb & lt)] = (uint16_t)s;
                    n += __popc(mb);
                }
                __syncwarp();
                // ---- validity gate; the slots that fail are released at once, the others move to the front
                uint32_t nv = 0;
                for (uint32_t i0 = 0; i0 < n; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    bool valid = false;
                    uint32_t s = 0;
                    if (i < n) {
                        s = cw[i];
                        const uint32_t c = keys[s];
                        if (c != rc.r) {  // identity cell dropped (library.cpp:485-487)
                            pairs++;
                            uint32_t inter, pc, tc;
                            slot_sums(t, c, cnt[s], &inter, &pc, &tc);
                            valid = gate(a.k2, pc, tc, rc.kr, a.meta[c].x);
                        }
                        if (!valid) {
                            keys[s] = kEmpty;
                            cnt[s] = 0;
                        }
                    }
                    const unsigned mb = __ballot_sync(0xffffffffu, valid);
                    if (valid) cw[nv + __popc(mb & lt)] = (uint16_t)s;
                    nv += __popc(mb);
                }
                __syncwarp();
                // ---- emit: one reservation per warp and row, dense coalesced stores
                if (nv) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(a.n_cells, (unsigned long long)nv);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    for (uint32_t i = lane; i < nv; i += 32) {
                        const uint32_t s = cw[i];
                        const uint32_t c = keys[s];
                        const uint32_t v = cnt[s];
                        keys[s] = kEmpty;
                        cnt[s] = 0;
                        uint32_t inter, pc, tc;
                        slot_sums(t, c, v, &inter, &pc, &tc);
                        const uint2 mc = a.meta[c];
                        emit_cell(a, rc, base + i, c, mc.y, mc.x, inter, pc, tc);
                    }
                }
            }
            if (nx) {  // side-table slots are released only after every reader is done (linear probing)
                __syncthreads();
                for (uint32_t i = tid; i < nx; i += THREADS) {
                    const uint32_t xs = t.xtouched[i];
                    t.xkeys[xs] = kEmpty;
                    t.xv0[xs] = 0;
                    t.xv1[xs] = 0;
                    t.xv2[xs] = 0;
                }
                __threadfence();
            }
        }
        buf ^= 1;
    }

    // one atomic per warp for the pair statistic
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, d);
    if (lane == 0 && pairs) atomicAdd(a.n_pairs, pairs);
}

inline size_t score_smem_bytes(uint32_t hbits, uint32_t fcap, int threads) {
    const size_t H = (size_t)1 << hbits;
    return H * 8 + (size_t)fcap * 16 + (size_t)(threads / 32) * sizeof(WarpScratch) + H * 2 + 16;
}

// Last resort for rows with more distinct columns than any shared-memory table holds: the reference's own scheme,
// S-sized accumulators in global memory (zero between rows) plus a touched list.  One warp per forward entry.
__global__ void __launch_bounds__(kDenseThreads) score_rows_dense_kernel(ScoreArgs a, DenseArgs d) {
    __shared__ uint32_t s_row;
    __shared__ uint32_t s_touched;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* inter = d.acc + (size_t)blockIdx.x * 4 * d.S;
    uint32_t* pcv = inter + d.S;
    uint32_t* tcv = pcv + d.S;
    uint32_t* touched = tcv + d.S;
    unsigned long long pairs = 0;

    for (;;) {
        if (tid == 0) {
            s_row = atomicAdd(a.cursor, 1u);
            s_touched = 0;
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
        for (uint32_t f = rw.fb + warp; f < rw.fe; f += kDenseThreads / 32) {
            const uint2 fw = a.fwd[f];
            const uint32_t m = (fw.y & kMulti) ? a.fwd_cnt[f] : 1u;
            const uint32_t gl = fw.y & ~kMulti;
            for (uint32_t p = lane; p < gl; p += 32) {
                const uint32_t e = a.post[fw.x + p];
                const uint32_t c = e & ~kMulti;
                const uint32_t n = (e & kMulti) ? a.post_cnt[fw.x + p] : 1u;
                const uint32_t old = atomicAdd(&tcv[c], n);  // counts are >= 1: old == 0 <=> first touch
                if (old == 0) touched[atomicAdd(&s_touched, 1u)] = c;
                atomicAdd(&inter[c], n < m ? n : m);
                atomicAdd(&pcv[c], m);
            }
        }
        __threadfence();
        __syncthreads();
        const uint32_t nt = s_touched;
        for (uint32_t b = 0; b < nt; b += kDenseThreads) {
            const uint32_t i = b + tid;
            bool want = false;
            uint32_t c = kEmpty, in = 0, pc = 0, tc = 0;
            uint2 mc = make_uint2(0u, 0u);
            if (i < nt) {
                c = *(volatile uint32_t*)&touched[i];
                in = *(volatile uint32_t*)&inter[c];
                pc = *(volatile uint32_t*)&pcv[c];
                tc = *(volatile uint32_t*)&tcv[c];
                inter[c] = 0;
                pcv[c] = 0;
                tcv[c] = 0;
                if (c != rc.r) {
                    pairs++;
                    mc = a.meta[c];
                    want = gate(a.k2, pc, tc, rc.kr, mc.x);
                }
            }
            // warp-aggregated append
            const unsigned mb = __ballot_sync(0xffffffffu, want);
            if (mb) {
                const unsigned leader = __ffs((int)mb) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(a.n_cells, (unsigned long long)__popc(mb));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (want) emit_cell(a, rc, base + __popc(mb & ((1u << lane) - 1u)), c, mc.y, mc.x, in, pc, tc);
            }
        }
        __threadfence();
        __syncthreads();
    }
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, dd);
    if (lane == 0 && pairs) atomicAdd(a.n_pairs, pairs);
}

}  // namespace sk
}  // namespace pd
