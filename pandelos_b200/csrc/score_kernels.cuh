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
//                index build into three classes (long and huge lists are walked first: they carry the homologs):
//                  short lists  (<= kShortList)  flattened in batches of 32 lists so every lane has a posting,
//                  long lists                    one warp walks one list, 8 x 32 consecutive postings per round,
//                                                the next round's loads issued while this round is accumulated,
//                  huge lists   (> kHugeList)    all warps of the CTA stride over the same list.
//   table        two tiers of (key, count) slots in shared memory.  Tier 1 is direct mapped by an ORDER-PRESERVING
//                function of the column, slot = floor(c * T1 / S): one 4-byte shared load + one shared atomic per
//                posting, and a row's frequent columns (its homologs, one per genome, far apart in gene order) each
//                keep a slot for the whole row.  Postings whose tier-1 slot belongs to another column go to a
//                per-warp queue and are probed / inserted into tier 2 (linear probing) 32 at a time, all lanes busy.
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
static const uint32_t kNone = 0x7FFFFFFFu;      // "no posting" in a lane's item (gene ids are < 2^31 - 1)
#ifndef PD_KITEMS
#define PD_KITEMS 8
#endif
static const int kItems = PD_KITEMS;            // postings per lane per round of a long list
static const int kItemsA = 4;                   // ... of the flattened short lists
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
    const uint32_t* n_rows_dev;  // non-null: rows per level live on the device (row_keys_kernel); this launch takes
    uint32_t level;              // the rows of level `level`: [sum of the counts before it, + n_rows_dev[level])
    const unsigned long long* n_rows_ovf;  // non-null (retry launch): the row count is the overflow counter of the launches before
    uint32_t* cursor;
    // parameters
    uint32_t G;
    uint32_t S;       // genes (debug bounds checks)
    uint32_t k2;      // 2k: perc >= thr  <=>  2k * pc >= K  (exact for K < 2^20, see gate())
    uint32_t hbits;   // log2 of the tier-2 slots H
    uint32_t hmul;    // tier-2 home slot of column c = floor(c * H / S) = umulhi(c, hmul)
    uint32_t hmask;   // H - 1
    uint32_t plimit;  // tier-2 probe limit
    uint32_t nslots;  // T1 + H
    uint32_t t1;      // tier-1 slots (a multiple of 32)
    uint32_t t1mul;   // tier-1 slot of column c = floor(c * T1 / S) = umulhi(c, t1mul): ORDER PRESERVING on purpose, see Tab
    uint32_t fcap;    // forward entries staged per segment
    // outputs
    float* o_score;
    float* o_perc;
    float* o_trperc;
    int32_t* o_row;
    int32_t* o_col;
    int32_t* o_g1;
    int32_t* o_g2;
    uint32_t g1_bhrow;               // non-zero: o_g1 receives the row's index inside the call instead of its genome
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

#if defined(PD_DEBUG_BOUNDS) && !defined(PD_EMU)
#define PD_CHECK(cond, id, v)                                                                                        \
    do {                                                                                                             \
        if (!(cond)) {                                                                                               \
            printf("PD_CHECK %d failed v=%u blk=%d thr=%d\n", id, (unsigned)(v), (int)blockIdx.x, (int)threadIdx.x); \
            __trap();                                                                                                \
        }                                                                                                            \
    } while (0)
#else
#define PD_CHECK(cond, id, v) \
    do {                      \
    } while (0)
#endif

static const uint32_t kDenseWordsPerGene = 5;
struct DenseArgs {
    uint32_t S;
    uint32_t* acc;  // per CTA: hits[S], three correction arrays [S], touched[S]; the first four all zero between rows
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
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
#ifdef PD_EMU
    *reinterpret_cast<uint4*>(smem_dst) = *reinterpret_cast<const uint4*>(gsrc);
#else
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifndef PD_EMU
    asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
}

// Shared-memory accesses of the hot loop by 32-bit shared-window address (no generic-address arithmetic per access).
#ifdef PD_EMU
typedef uintptr_t saddr_t;
__device__ __forceinline__ saddr_t smem_addr(const void* p) { return reinterpret_cast<uintptr_t>(p); }
__device__ __forceinline__ uint32_t lds_u32(saddr_t a) { return *reinterpret_cast<volatile uint32_t*>(a); }
__device__ __forceinline__ uint32_t atoms_cas(saddr_t a, uint32_t cmp, uint32_t val) { return atomicCAS(reinterpret_cast<uint32_t*>(a), cmp, val); }
__device__ __forceinline__ void reds_inc(saddr_t a) { atomicAdd(reinterpret_cast<uint32_t*>(a), 1u); }
#else
typedef uint32_t saddr_t;
__device__ __forceinline__ saddr_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t lds_u32(saddr_t a) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t atoms_cas(saddr_t a, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;\n" : "=r"(old) : "r"(a), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ void reds_inc(saddr_t a) { asm volatile("red.shared.add.u32 [%0], 1;\n" ::"r"(a) : "memory"); }
#endif

// Validity gate (library.cpp:497-500) in integers:
//   (float)pc / (float)K >= 1.0f / (2.0f * (float)k)   <=>   2k * pc >= K        for K < 2^20
// (=>) division and reciprocal are correctly rounded and rounding is monotonic; (<=) if 2k*pc <= K - 1 the two reals
// differ by more than 2^-20 relative, far more than the two roundings (2^-24 each) can close.
// Only cells that pass pay for the three float divisions.  inter >= 1 for every touched column, so score > 0
// (library.cpp:505) is the same condition.
__device__ __forceinline__ bool gate(uint32_t k2, uint32_t pc, uint32_t tc, uint32_t kr, uint32_t kc) {
    return (k2 * pc >= kr) || (k2 * tc >= kc);
}

// The accumulator has two tiers, both (key, count) arrays in shared memory:
//   tier 1  T1 slots, DIRECT MAPPED by an order-preserving function of the column, slot = floor(c * T1 / S).  Posting
//           lists are sorted by gene and a row's homologs sit in different genomes, i.e. far apart in gene order: the
//           lanes of a warp, which hold consecutive postings, hit ascending, well separated slots (no bank conflicts),
//           and a homolog column keeps its slot for the whole row.  One 4-byte shared load + one shared atomic.
//   tier 2  H slots, linear probing.  Takes the columns whose tier-1 slot belongs to another column.  Postings that
//           miss tier 1 go to a per-warp queue and are probed / inserted here 32 at a time, every lane busy.
// A column lives in exactly one place: a tier-1 slot is claimed once (CAS on empty) and never changes owner during a
// row, so a column that lost its slot goes to tier 2 every time.  Slots are numbered [0, T1) and [T1, T1 + H).
struct RowCtl {  // per-row control words in shared memory
    uint32_t nx;      // side-table entries in use
    int over;         // the row does not fit: stop, hand it to the next level
    uint32_t ctr[2];  // work counters: batches of short lists, long lists
};

// Launch-uniform geometry (tier sizes, slot functions) is read from the kernel arguments (constant bank), not carried
// in registers; Tab holds only what is per CTA / per row.
struct Tab {
    uint32_t* keys;   // nslots keys, then nslots counts (generic address, finalize)
    saddr_t keys_sa;  // the same by shared-window address (hot loop)
    RowCtl* ctl;
    __device__ __forceinline__ volatile int* over() const { return &ctl->over; }
};
__device__ __forceinline__ saddr_t cnt_sa(const ScoreArgs& a, const Tab& t) { return t.keys_sa + a.nslots * 4u; }
__device__ __forceinline__ uint32_t* cnt_of(const ScoreArgs& a, const Tab& t) { return t.keys + a.nslots; }

// side table of this CTA: keys, d_inter, d_pc, d_tc, touched list (kXSlots each)
struct XTab {
    uint32_t* base;
    RowCtl* ctl;
    __device__ __forceinline__ uint32_t* keys() const { return base; }
    __device__ __forceinline__ uint32_t* v(int i) const { return base + (1 + i) * kXSlots; }
    __device__ __forceinline__ uint32_t* touched() const { return base + 4 * kXSlots; }
};
__device__ __forceinline__ XTab xtab_of(const ScoreArgs& a, const Tab& t) {
    XTab x;
    x.base = a.xtab + (size_t)blockIdx.x * (5 * kXSlots);
    x.ctl = t.ctl;
    return x;
}

// Tier-1 slot of column c: the order-preserving index floor(c * T1 / S), one multiply.  c must be a gene (< S): the hot
// loop pads with the row's own gene, whose cell is dropped anyway.  (Round 2 also tried the index rotated so that the bank
// follows the genome number: no effect on the kernel time — shared-memory conflicts are not what bounds it — and three
// more instructions per posting, so it is gone.)
__device__ __forceinline__ uint32_t t1_slot_of(uint32_t c, uint32_t t1mul) { return __umulhi(c, t1mul); }
__device__ __forceinline__ uint32_t t1_slot(const ScoreArgs& a, uint32_t c) { return t1_slot_of(c, a.t1mul); }
__device__ __forceinline__ uint32_t t2_home(const ScoreArgs& a, uint32_t c) { return min(__umulhi(c, a.hmul), a.hmask); }

__device__ __forceinline__ uint32_t x_find_or_insert(const XTab& x, uint32_t c) {
    uint32_t h = __umulhi(c * 0x85EBCA6Bu, kXSlots);
    for (uint32_t probe = 0; probe < kXSlots; probe++) {
        const uint32_t k = *(volatile uint32_t*)(x.keys() + h);
        if (k == c) return h;
        if (k == kEmpty) {
            const uint32_t old = atomicCAS(x.keys() + h, kEmpty, c);
            if (old == kEmpty) {
                const uint32_t xi = atomicAdd(&x.ctl->nx, 1u);
                if (xi < kXCap) x.touched()[xi] = h;
                else x.ctl->over = 1;
                return h;
            }
            if (old == c) return h;
        }
        h = (h + 1 == kXSlots) ? 0 : h + 1;
    }
    x.ctl->over = 1;
    return kEmpty;
}

// tier 2: finds column c or claims the first free slot of its probe path; returns the slot (T1-based), kEmpty if the
// row does not fit
__device__ __forceinline__ uint32_t t2_find_or_insert(const ScoreArgs& a, const Tab& t, uint32_t c) {
    uint32_t hh = t2_home(a, c), probes = 0;
    for (;;) {
        PD_CHECK(a.t1 + hh < a.nslots, 2, hh);
        const saddr_t addr = t.keys_sa + (a.t1 + hh) * 4u;
        uint32_t kk = lds_u32(addr);
        if (kk == kEmpty) {
            const uint32_t old = atoms_cas(addr, kEmpty, c);
            kk = (old == kEmpty) ? c : old;
        }
        if (kk == c) return a.t1 + hh;
        hh = (hh + 1) & a.hmask;
        if (++probes > a.plimit) {
            *t.over() = 1;
            return kEmpty;
        }
    }
}

// The general case of one posting: column c held n times by its gene, m times by the row (n | m > 1: rare).
// Never inlined, and handed everything it needs BY VALUE: a reference to the kernel arguments or to the hot loop's
// state would force them into local memory.
struct GenArgs {
    uint32_t* keys;
    saddr_t keys_sa;
    RowCtl* ctl;
    uint32_t* xbase;
    uint32_t t1, t1mul, hmask, hmul, plimit, nslots;
};
__device__ __noinline__ void add_general(const GenArgs g, uint32_t c, uint32_t n, uint32_t m) {
    // the slot of column c in either tier (claimed if new)
    uint32_t h = t1_slot_of(c, g.t1mul);
    {
        const saddr_t a1 = g.keys_sa + h * 4u;
        uint32_t k = lds_u32(a1);
        if (k == kEmpty) {
            const uint32_t old = atoms_cas(a1, kEmpty, c);
            k = (old == kEmpty) ? c : old;
        }
        if (k != c) {
            uint32_t hh = min(__umulhi(c, g.hmul), g.hmask), probes = 0;
            for (;;) {
                const saddr_t addr = g.keys_sa + (g.t1 + hh) * 4u;
                uint32_t kk = lds_u32(addr);
                if (kk == kEmpty) {
                    const uint32_t old = atoms_cas(addr, kEmpty, c);
                    kk = (old == kEmpty) ? c : old;
                }
                if (kk == c) break;
                hh = (hh + 1) & g.hmask;
                if (++probes > g.plimit) {
                    g.ctl->over = 1;
                    return;
                }
            }
            h = g.t1 + hh;
        }
    }
    uint32_t* cnt = g.keys + g.nslots;
    PD_CHECK(h < g.nslots, 7, h);
    atomicAdd(&cnt[h], 1u);
    if ((n | m) > 1u) {  // corrections for the repeated k-mer go to the side table
        atomicOr(&cnt[h], kFlag);
        XTab x;
        x.base = g.xbase;
        x.ctl = g.ctl;
        const uint32_t xs = x_find_or_insert(x, c);
        if (xs != kEmpty) {
            const uint32_t mn = n < m ? n : m;
            if (mn > 1) atomicAdd(&x.v(0)[xs], mn - 1);
            if (m > 1) atomicAdd(&x.v(1)[xs], m - 1);
            if (n > 1) atomicAdd(&x.v(2)[xs], n - 1);
        }
    }
}
__device__ __forceinline__ GenArgs gen_args(const ScoreArgs& a, const Tab& t) {
    GenArgs g;
    g.keys = t.keys;
    g.keys_sa = t.keys_sa;
    g.ctl = t.ctl;
    g.xbase = a.xtab + (size_t)blockIdx.x * (5 * kXSlots);
    g.t1 = a.t1;
    g.t1mul = a.t1mul;
    g.hmask = a.hmask;
    g.hmul = a.hmul;
    g.plimit = a.plimit;
    g.nslots = a.nslots;
    return g;
}

// ---- accumulate, the common case: postings with n = m = 1.  One "item" is one posting per lane; lanes past the end of
// a list hold the row's own gene (its cell is dropped by finalize), so the hot path has no "no posting" case.
//
// A posting whose tier-1 slot does not hold its column (a few per cent of them: first touches, second paralogs, chance
// hits) is NOT handled where it is found: with 64 postings per pair of items almost every pair has one, and a branch
// taken for one or two lanes costs the whole warp its issue slots.  Each LANE appends its misses to its own column of a
// per-warp buffer (kLaneQueue rows of 32 words: one predicated store, no vote, no atomic); every four items one vote
// asks whether some lane is about to run out of rows, and only then the warp resolves everything queued: claim the
// tier-1 slot if it is still free, count there if the column owns it, else probe / insert into tier 2.
static const uint32_t kLaneQueue = 8;   // rows per lane
static const uint32_t kLaneDrain = 5;   // drain when a lane holds this many: at most 4 + 4 between two checks
struct LaneQueue {
    saddr_t base;  // this lane's word of row 0
    saddr_t wr;    // where this lane's next miss goes (base + 128 * queued)
};

__device__ __forceinline__ void sts_u32(saddr_t a, uint32_t v) {
#ifdef PD_EMU
    *reinterpret_cast<volatile uint32_t*>(a) = v;
#else
    asm volatile("st.volatile.shared.u32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory");
#endif
}

// one posting: count it if tier 1 holds its column, else queue it.  No branch: the increment of a miss is redirected to
// the lane's own dump word (the row after its queue rows), the queue store is predicated.
__device__ __forceinline__ void item_add(const ScoreArgs& a, const Tab& t, LaneQueue& lq, uint32_t c) {
    const uint32_t s = t1_slot(a, c);
    PD_CHECK(s < a.t1 && c < a.S, 1, c);
    const saddr_t ka = t.keys_sa + s * 4u;
    const uint32_t k = lds_u32(ka);
#ifdef PD_EMU
    if (k == c) reds_inc(ka + a.nslots * 4u);
    else {
        sts_u32(lq.wr, c);
        lq.wr += 128u;
    }
#else
    const saddr_t hit_at = ka + a.nslots * 4u, dump_at = lq.base + kLaneQueue * 128u;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 ad, nw;\n"
        "setp.eq.u32 p, %1, %2;\n"
        "selp.u32 ad, %3, %4, p;\n"
        "red.shared.add.u32 [ad], 1;\n"
        "@!p st.volatile.shared.u32 [%0], %2;\n"
        "add.u32 nw, %0, 128;\n"
        "selp.u32 %0, %0, nw, p;\n"
        "}\n"
        : "+r"(lq.wr)
        : "r"(k), "r"(c), "r"(hit_at), "r"(dump_at)
        : "memory");
#endif
}

// resolves everything the lanes have queued (warp-uniform call; n = this lane's count).  The lanes' columns are first
// packed into one dense list (exclusive prefix of the counts) so that the resolve loop runs with every lane busy instead
// of as many times as the fullest lane has entries.
__device__ __noinline__ void queue_drain(const GenArgs g, saddr_t base, uint32_t n) {
    const unsigned lane = threadIdx.x & 31;
    uint32_t v[kLaneQueue];
#pragma unroll
    for (uint32_t i = 0; i < kLaneQueue; i++) v[i] = i < n ? lds_u32(base + i * 128u) : 0u;
    uint32_t incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const saddr_t flat = base - lane * 4u + (incl - n) * 4u;  // the rows as one array of 32 * kLaneQueue words
    __syncwarp();
#pragma unroll
    for (uint32_t i = 0; i < kLaneQueue; i++)
        if (i < n) sts_u32(flat + i * 4u, v[i]);
    __syncwarp();
    for (uint32_t j = lane; j < total; j += 32) {
        const uint32_t cur = lds_u32(base + (j - lane) * 4u);
        const uint32_t s1 = t1_slot_of(cur, g.t1mul);
        const saddr_t a1 = g.keys_sa + s1 * 4u;
        uint32_t k1 = lds_u32(a1);
        if (k1 == kEmpty) {
            const uint32_t old = atoms_cas(a1, kEmpty, cur);
            k1 = (old == kEmpty) ? cur : old;
        }
        uint32_t slot = s1;
        if (k1 != cur) {
            uint32_t hh = min(__umulhi(cur, g.hmul), g.hmask), probes = 0;
            for (;;) {
                const saddr_t addr = g.keys_sa + (g.t1 + hh) * 4u;
                uint32_t kk = lds_u32(addr);
                if (kk == kEmpty) {
                    const uint32_t old = atoms_cas(addr, kEmpty, cur);
                    kk = (old == kEmpty) ? cur : old;
                }
                if (kk == cur) break;
                hh = (hh + 1) & g.hmask;
                if (++probes > g.plimit) {
                    g.ctl->over = 1;
                    hh = kEmpty;
                    break;
                }
            }
            if (hh == kEmpty) continue;
            slot = g.t1 + hh;
        }
        reds_inc(g.keys_sa + (g.nslots + slot) * 4u);
    }
    __syncwarp();
}
__device__ __forceinline__ void queue_check(const ScoreArgs& a, const Tab& t, LaneQueue& lq, uint32_t threshold) {
    if (__any_sync(0xffffffffu, lq.wr - lq.base >= threshold * 128u)) {
        queue_drain(gen_args(a, t), lq.base, (lq.wr - lq.base) >> 7);
        lq.wr = lq.base;
    }
}

// Per-warp scratch of the flattened walk over short lists
struct WarpScratch {
    uint32_t bits[kShortList + kItemsA];  // 32 lists x kShortList postings = 2048 marks (+ the words a round reads ahead)
    uint32_t pre[32];
    uint32_t queue[kLaneQueue * 32];  // row i: the i-th queued miss of every lane
    uint32_t dump[32];                // row kLaneQueue: where a lane's increment goes when its posting missed (item_add)
    uint32_t pad[4];
};

// A list walk in rounds of kItems x 32 consecutive postings: lane l holds postings p0 + 32 u + l, u < kItems, in e[u]
// (the row's own gene `self` past the end of the list).  round_step consumes e pair by pair and refills each pair at
// once with the NEXT round's postings (list `nq`, `nrem` postings from this lane's first to that list's end; nrem <= 0:
// nothing), so kItems loads per warp stay in flight while the table is updated.  cnt = postings in this round
// (uniform), the row holds this list's k-mer mj times; pos = index of this round's first posting in the posting array
// (for post_cnt).
__device__ __forceinline__ void round_step(const ScoreArgs& a, const Tab& t, LaneQueue& lq, uint32_t (&e)[kItems], uint32_t cnt,
                                           uint32_t pos, uint32_t mj, const uint32_t* __restrict__ nq, int nrem, uint32_t self) {
    const unsigned lane = threadIdx.x & 31;
    if (kItems % 4 == 0 && cnt >= 32u * kItems && mj == 1) {
        // a full round of a k-mer the row holds once (nearly all rounds of the conserved lists): four postings per lane at
        // a time, no per-pair checks; a repeated posting (bit 31) is a per-lane side step
        const bool next_full = __all_sync(0xffffffffu, nrem > 32 * (kItems - 1));  // the next round is full too: loads need no predicate
#pragma unroll
        for (int u = 0; u < kItems; u += 4) {
            uint32_t c[4];
            if (next_full) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    c[i] = e[u + i];
                    e[u + i] = nq[32 * (u + i)];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    c[i] = e[u + i];
                    e[u + i] = nrem > 32 * (u + i) ? nq[32 * (u + i)] : self;
                }
            }
            if ((c[0] | c[1] | c[2] | c[3]) & kMulti) {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (c[i] & kMulti) {
                        add_general(gen_args(a, t), c[i] & ~kMulti, a.post_cnt[pos + 32u * (u + i) + lane], 1u);
                        c[i] = self;
                    }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) item_add(a, t, lq, c[i]);
            queue_check(a, t, lq, kLaneDrain);
        }
        return;
    }
#pragma unroll
    for (int u = 0; u < kItems; u += 2) {
        uint32_t c0 = e[u], c1 = e[u + 1];
        e[u] = nrem > 32 * u ? nq[32 * u] : self;
        e[u + 1] = nrem > 32 * (u + 1) ? nq[32 * (u + 1)] : self;
        if (32u * u < cnt) {  // uniform: the tail of a list does not pay for empty items
            // bit 31: a repeated k-mer; the row's own repeat count is uniform over the list
            if (((c0 | c1) & kMulti) || mj > 1) {
                if (32u * u + lane < cnt && ((c0 & kMulti) || mj > 1)) {
                    add_general(gen_args(a, t), c0 & ~kMulti, (c0 & kMulti) ? a.post_cnt[pos + 32u * u + lane] : 1u, mj);
                    c0 = self;
                }
                if (32u * (u + 1) + lane < cnt && ((c1 & kMulti) || mj > 1)) {
                    add_general(gen_args(a, t), c1 & ~kMulti, (c1 & kMulti) ? a.post_cnt[pos + 32u * (u + 1) + lane] : 1u, mj);
                    c1 = self;
                }
            }
            item_add(a, t, lq, c0);
            item_add(a, t, lq, c1);
        }
        if ((u & 2) != 0) queue_check(a, t, lq, kLaneDrain);  // every four items
    }
}

// Accumulates the staged forward entries fbuf[0, n_stage) = forward entries [f0, f0 + n_stage) of the row `self`.
// Classes by position: [0, ns) short, [ns, nl) long, [nl, n_stage) huge.  ctr[0..1] are shared work counters, zero
// at entry.  No ordering is needed between the three parts (every update is an atomic on the table); the long lists
// go first so that the row's frequent columns — its homologs — claim the tier-1 slots before the chance hits of the
// short lists arrive.
template <int THREADS>
__device__ __forceinline__ void accumulate(const ScoreArgs& a, const Tab& t_in, const uint2* fbuf, uint32_t f0, uint32_t ns, uint32_t nl,
                                           uint32_t n_stage, WarpScratch* ws_all, uint32_t* ctr, uint32_t self) {
    constexpr int WARPS = THREADS / 32;
    Tab t = t_in;
#ifndef PD_EMU
    // the table's shared-window address stays in a register: left to itself the compiler re-derives it (special register
    // read + three integer ops) for every pair of postings
    asm volatile("" : "+r"(t.keys_sa));
#endif
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = ws_all[warp];
    LaneQueue lq;
    lq.base = smem_addr(ws.queue + lane);
    lq.wr = lq.base;
    const unsigned le = 0xffffffffu >> (31 - lane);  // lanes <= mine
    // ---- long lists: one warp per list
    {
        uint32_t e[kItems];
        uint32_t gs = 0, gl = 0, mj = 1, p0 = 0;
        auto claim = [&](uint32_t& cgs, uint32_t& cgl, uint32_t& cmj) -> bool {
            uint32_t li = 0;
            if (lane == 0) li = *t.over() ? 0x7FFFFFFFu : atomicAdd(ctr + 1, 1u);  // one lane polls the stop flag
            li = __shfl_sync(0xffffffffu, li, 0);
            if (li >= nl - ns) return false;
            const uint2 fw = fbuf[ns + li];
            cmj = (fw.y & kMulti) ? a.fwd_cnt[f0 + ns + li] : 1u;
            cgl = fw.y & ~kMulti;
            cgs = fw.x;
            return true;
        };
        bool have = claim(gs, gl, mj);
        if (have) {
#pragma unroll
            for (int u = 0; u < kItems; u++) e[u] = 32u * u + lane < gl ? a.post[gs + 32u * u + lane] : self;
        }
        while (have) {
            uint32_t ngs = gs, ngl = gl, nmj = mj, np0 = p0 + 32 * kItems;
            bool nhave = true;
            if (np0 >= gl) {
                nhave = claim(ngs, ngl, nmj);
                np0 = 0;
            }
            const int nrem = nhave ? (int)(ngl - np0) - (int)lane : 0;
            round_step(a, t, lq, e, gl - p0, gs + p0, mj, a.post + ngs + np0 + lane, nrem, self);
            gs = ngs;
            gl = ngl;
            mj = nmj;
            p0 = np0;
            have = nhave;
        }
    }
    // ---- huge lists: the whole CTA strides over each
    for (uint32_t j = nl; j < n_stage; j++) {
        if (__any_sync(0xffffffffu, *t.over() != 0)) break;
        const uint2 fw = fbuf[j];
        const uint32_t mj = (fw.y & kMulti) ? a.fwd_cnt[f0 + j] : 1u;
        const uint32_t gl = fw.y & ~kMulti;
        const uint32_t stride = WARPS * 32 * kItems;
        uint32_t p0 = warp * (32 * kItems);
        if (p0 >= gl) continue;
        uint32_t e[kItems];
#pragma unroll
        for (int u = 0; u < kItems; u++) e[u] = p0 + 32u * u + lane < gl ? a.post[fw.x + p0 + 32u * u + lane] : self;
        for (;;) {
            const uint32_t np0 = p0 + stride;
            const bool more = np0 < gl && !((np0 & 0x7FFFu) < stride && __any_sync(0xffffffffu, *t.over() != 0));  // poll now and then
            const int nrem = more ? (int)(gl - np0) - (int)lane : 0;
            round_step(a, t, lq, e, gl - p0, fw.x + p0, mj, a.post + fw.x + np0 + lane, nrem, self);
            if (!more) break;
            p0 = np0;
        }
    }
    // ---- short lists: batches of 32, flattened
    for (;;) {
        uint32_t bi = 0;
        if (lane == 0) bi = *t.over() ? 0x03FFFFFFu : atomicAdd(ctr, 1u);  // one lane polls the stop flag: uniform exit
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
        if (lane < kItemsA) ws.bits[64 + lane] = 0;
        ws.pre[lane] = pre;
        __syncwarp();
        if (has) atomicOr(&ws.bits[pre >> 5], 1u << (pre & 31));  // lists are >= 2 long: one mark per list
        __syncwarp();
        uint32_t seen = 0;  // marks before this round
        for (uint32_t t0 = 0; t0 < total; t0 += 32 * kItemsA) {
            uint32_t e[kItemsA], o[kItemsA];
#pragma unroll
            for (int u = 0; u < kItemsA; u++) {
                const uint32_t w = ws.bits[(t0 >> 5) + u];
                o[u] = (seen + __popc(w & le) - 1u) & 31u;
                seen += __popc(w);
            }
            uint32_t special = 0;
#pragma unroll
            for (int u = 0; u < kItemsA; u++) {
                const uint32_t tt = t0 + 32u * u + lane;
                e[u] = self;
                if (tt < total) {
                    const uint2 fw = fbuf[b0 + o[u]];
                    e[u] = a.post[fw.x + (tt - ws.pre[o[u]])];
                    special |= fw.y | e[u];
                }
            }
            if (special & kMulti) {
#pragma unroll
                for (int u = 0; u < kItemsA; u++) {
                    const uint32_t tt = t0 + 32u * u + lane;
                    if (tt < total) {
                        const uint2 fw = fbuf[b0 + o[u]];
                        if ((fw.y | e[u]) & kMulti) {
                            const uint32_t n = (e[u] & kMulti) ? a.post_cnt[fw.x + (tt - ws.pre[o[u]])] : 1u;
                            const uint32_t m = (fw.y & kMulti) ? a.fwd_cnt[f0 + b0 + o[u]] : 1u;
                            add_general(gen_args(a, t), e[u] & ~kMulti, n, m);
                            e[u] = self;
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kItemsA; u++) {
                if (t0 + 32u * u >= total) break;
                item_add(a, t, lq, e[u]);
            }
            queue_check(a, t, lq, kLaneDrain);
        }
        __syncwarp();
    }
    queue_check(a, t, lq, 1);
}

// (inter, pc, tc) of a table slot
__device__ __forceinline__ void slot_sums(const XTab& x, uint32_t c, uint32_t v, uint32_t* inter, uint32_t* pc, uint32_t* tc) {
    uint32_t i = v & ~kFlag, p = i, q = i;
    if (v & kFlag) {
        uint32_t xs = __umulhi(c * 0x85EBCA6Bu, kXSlots);
        while (*(volatile uint32_t*)(x.keys() + xs) != c) xs = (xs + 1 == kXSlots) ? 0 : xs + 1;
        i += *(volatile uint32_t*)(x.v(0) + xs);
        p += *(volatile uint32_t*)(x.v(1) + xs);
        q += *(volatile uint32_t*)(x.v(2) + xs);
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
        a.o_g1[idx] = (int32_t)(a.g1_bhrow ? rc.bh_row : rc.gr);
        a.o_g2[idx] = (int32_t)gc;
    }
    PD_CHECK(gc < a.G && c < a.S, 6, gc);
    // scores are positive floats: their bit patterns order like the values
    atomicMax(&a.bh[(size_t)rc.bh_row * a.G + gc], __float_as_uint(score));
    if (a.colmax) atomicMax(&a.colmax[c], __float_as_uint(score));
}

// ------------------------------------------------------------------------------------------------ main kernel

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) score_rows_kernel(ScoreArgs a) {
    constexpr int WARPS = THREADS / 32;
    PD_DYNAMIC_SMEM(smem_raw);
    const uint32_t H = a.nslots;  // slots of both tiers
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* cnt = keys + H;
    uint2* fwdbuf = reinterpret_cast<uint2*>(cnt + H);                                  // 2 x fcap
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(fwdbuf + 2 * (size_t)a.fcap);     // WARPS
    uint16_t* cand = reinterpret_cast<uint16_t*>(ws + WARPS);                           // one per slot, in warp slices
    __shared__ uint4 s_desc[2][2];  // two RowDesc buffers
    // per-row control words, double-buffered like s_desc: a row uses [buf]; thread 0 clears [buf ^ 1] while the row
    // runs (its last readers finished before the row's opening barrier)
    __shared__ RowCtl s_ctl[2];

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t fcap = a.fcap;

    Tab t;
    t.keys = keys;
    t.keys_sa = smem_addr(keys);

    for (uint32_t i = tid; i < H; i += THREADS) {
        keys[i] = kEmpty;
        cnt[i] = 0;
    }
    uint32_t n_rows = a.n_rows, row0 = 0;
    if (a.n_rows_dev) {
        n_rows = a.n_rows_dev[a.level];
        for (uint32_t l = 0; l < a.level; l++) row0 += a.n_rows_dev[l];
    }
    if (a.n_rows_ovf) n_rows = (uint32_t)*a.n_rows_ovf;
    const uint4* rows4 = reinterpret_cast<const uint4*>(a.rows + row0);
    uint32_t idx_next = 0;  // thread 0: row claimed for the iteration after this one
    if (tid == 0) {
        const uint32_t ri = atomicAdd(a.cursor, 1u);
        idx_next = atomicAdd(a.cursor, 1u);
        uint4 d0 = make_uint4(kEmpty, 0u, 0u, 0u), d1 = make_uint4(0u, 0u, 0u, 0u);
        if (ri < n_rows) {
            d0 = rows4[2 * (size_t)ri];
            d1 = rows4[2 * (size_t)ri + 1];
        }
        s_desc[0][0] = d0;
        s_desc[0][1] = d1;
        s_ctl[0].nx = s_ctl[1].nx = 0;
        s_ctl[0].over = s_ctl[1].over = 0;
        s_ctl[0].ctr[0] = s_ctl[0].ctr[1] = s_ctl[1].ctr[0] = s_ctl[1].ctr[1] = 0;
    }
    __syncthreads();
    {
        const uint4 d0 = s_desc[0][0];
        if (d0.x != kEmpty) {
            const uint32_t f1 = d0.w - d0.z < fcap ? d0.w : d0.z + fcap;
            for (uint32_t f = d0.z + tid; f < f1; f += THREADS) cp_async8(&fwdbuf[f - d0.z], &a.fwd[f]);
        }
    }
    int buf = 0;

    for (;;) {
        cp_async_wait_all();
        __syncthreads();  // the row's descriptor and first forward segment are staged; the table is clean
        // Nothing of the row's descriptor or of the look-ahead is kept in registers across the accumulate phase (the hot
        // loop runs at the register cap): the descriptor is re-read from shared memory afterwards, the next row's
        // descriptor travels by cp.async straight into the other buffer.
        if (s_desc[buf][0].x == kEmpty) break;
        t.ctl = &s_ctl[buf];
        uint32_t idx_after = 0;
        if (tid == 0) {
            if (idx_next < n_rows) {
                cp_async16(&s_desc[buf ^ 1][0], &rows4[2 * (size_t)idx_next]);
                cp_async16(&s_desc[buf ^ 1][1], &rows4[2 * (size_t)idx_next + 1]);
            } else {
                s_desc[buf ^ 1][0] = make_uint4(kEmpty, 0u, 0u, 0u);
            }
            idx_after = atomicAdd(a.cursor, 1u);  // consumed after the accumulate phase
        }
        uint2* fbuf = fwdbuf + (size_t)buf * fcap;

        // ---- accumulate, one staged segment of forward entries at a time
        {
            const uint32_t fb = s_desc[buf][0].z, fe = s_desc[buf][0].w, fm = s_desc[buf][1].z, fh = s_desc[buf][1].w;
            for (uint32_t f0 = fb;;) {
                const uint32_t f1 = fe - f0 < fcap ? fe : f0 + fcap;
                const uint32_t ns = fm > f0 ? (fm < f1 ? fm - f0 : f1 - f0) : 0u;
                const uint32_t nl = fh > f0 ? (fh < f1 ? fh - f0 : f1 - f0) : 0u;
                accumulate<THREADS>(a, t, fbuf, f0, ns, nl, f1 - f0, ws, s_ctl[buf].ctr, s_desc[buf][0].x);
                if (f1 >= fe) break;
                __syncthreads();
                f0 = f1;
                const uint32_t f2 = fe - f0 < fcap ? fe : f0 + fcap;
                for (uint32_t f = f0 + tid; f < f2; f += THREADS) fbuf[f - f0] = a.fwd[f];
                if (tid == 0) s_ctl[buf].ctr[0] = s_ctl[buf].ctr[1] = 0;
                __syncthreads();
            }
        }
        if (tid == 0) {
            s_ctl[buf ^ 1].nx = 0;
            s_ctl[buf ^ 1].over = 0;
            s_ctl[buf ^ 1].ctr[0] = 0;
            s_ctl[buf ^ 1].ctr[1] = 0;
            idx_next = idx_after;
        }
        cp_async_wait_all();  // thread 0: the next row's descriptor has landed
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
        const uint4 rw0 = s_desc[buf][0], rw1 = s_desc[buf][1];  // (gene, bh_row, fb, fe), (kr, gr, fm, fh)
        RowCtx rc;
        rc.r = rw0.x;
        rc.bh_row = rw0.y;
        rc.kr = rw1.x;
        rc.gr = rw1.y;
        const bool over = *t.over() != 0;
        const uint32_t nx = t.ctl->nx < kXCap ? t.ctl->nx : kXCap;
        uint32_t pairs = 0;  // candidate cells of this row seen by this lane
        XTab xt;
        xt.base = a.xtab + (size_t)blockIdx.x * (5 * kXSlots);
        // each warp owns a slice of the table (both tiers, slots [0, H)); H is a multiple of 32
        const uint32_t spw = ((H / 32 + WARPS - 1) / WARPS) * 32;
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
            if (t.ctl->nx) {
                for (uint32_t i = tid; i < kXSlots; i += THREADS) {
                    xt.keys()[i] = kEmpty;
                    xt.v(0)[i] = 0;
                    xt.v(1)[i] = 0;
                    xt.v(2)[i] = 0;
                }
                __threadfence();
            }
        } else {
            if (nx) __threadfence();
            if (lo < H) {
                uint16_t* cw = cand + lo;
                // ---- occupied slots of the slice, compacted
                uint32_t n = 0;
                for (uint32_t s0 = 0; s0 < spw && lo + s0 < H; s0 += 32) {
                    const uint32_t s = lo + s0 + lane;
                    const bool occ = keys[s] != kEmpty;
                    const unsigned mb = __ballot_sync(0xffffffffu, occ);
                    PD_CHECK(s < H && n + 32 <= spw + 32, 4, s);
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
                        PD_CHECK(s < H && c < a.S, 5, c);
                        if (c != rc.r) {  // identity cell dropped (library.cpp:485-487)
                            pairs++;
                            uint32_t inter, pc, tc;
                            slot_sums(xt, c, cnt[s], &inter, &pc, &tc);
                            // the row-side test needs no gather: homologs pass it, and only what fails it looks up the column's length
                            valid = (a.k2 * pc >= rc.kr) || gate(a.k2, pc, tc, rc.kr, a.meta[c].x);
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
                        slot_sums(xt, c, v, &inter, &pc, &tc);
                        const uint2 mc = a.meta[c];
                        emit_cell(a, rc, base + i, c, mc.y, mc.x, inter, pc, tc);
                    }
                }
            }
            if (nx) {  // side-table slots are released only after every reader is done (linear probing)
                __syncthreads();
                for (uint32_t i = tid; i < nx; i += THREADS) {
                    const uint32_t xs = xt.touched()[i];
                    xt.keys()[xs] = kEmpty;
                    xt.v(0)[xs] = 0;
                    xt.v(1)[xs] = 0;
                    xt.v(2)[xs] = 0;
                }
                __threadfence();
            }
        }
        // one atomic per warp and row for the pair statistic (kept out of the registers of the accumulate phase)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, d);
        if (lane == 0 && pairs) atomicAdd(a.n_pairs, (unsigned long long)pairs);
        buf ^= 1;
    }
}

inline size_t score_smem_bytes(uint32_t t1, uint32_t hbits, uint32_t fcap, int threads) {
    const size_t H = (size_t)t1 + ((size_t)1 << hbits);
    return H * 8 + (size_t)fcap * 16 + (size_t)(threads / 32) * (sizeof(WarpScratch) + 64) + H * 2 + 16;
}

// Last resort for rows with more distinct columns than any shared-memory table holds: the reference's own scheme,
// S-sized accumulators in global memory (zero between rows) plus a touched list.  One warp per forward entry.
__global__ void __launch_bounds__(kDenseThreads) score_rows_dense_kernel(ScoreArgs a, DenseArgs d) {
    __shared__ uint32_t s_row;
    __shared__ uint32_t s_touched;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // per CTA: hits[S] (postings seen per column) and three correction arrays for postings with a multiplicity above one
    // on either side (rare: U/N > 0.999) — one atomic per posting instead of three — and the touched list
    uint32_t* hits = d.acc + (size_t)blockIdx.x * kDenseWordsPerGene * d.S;
    uint32_t* xin = hits + d.S;   // sum of min(n, m) - 1
    uint32_t* xpc = xin + d.S;    // sum of m - 1
    uint32_t* xtc = xpc + d.S;    // sum of n - 1
    uint32_t* touched = xtc + d.S;
    unsigned long long pairs = 0;
    const uint32_t n_rows = a.n_rows_dev ? *a.n_rows_dev : a.n_rows;

    for (;;) {
        if (tid == 0) {
            s_row = atomicAdd(a.cursor, 1u);
            s_touched = 0;
        }
        __syncthreads();
        const uint32_t ri = s_row;
        if (ri >= n_rows) break;
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
                const uint32_t old = atomicAdd(&hits[c], 1u);  // old == 0 <=> first touch
                if (old == 0) touched[atomicAdd(&s_touched, 1u)] = c;
                if ((n | m) > 1u) {
                    atomicAdd(&xin[c], (n < m ? n : m) - 1u);
                    atomicAdd(&xpc[c], m - 1u);
                    atomicAdd(&xtc[c], n - 1u);
                }
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
                const uint32_t h = *(volatile uint32_t*)&hits[c];
                in = h + *(volatile uint32_t*)&xin[c];
                pc = h + *(volatile uint32_t*)&xpc[c];
                tc = h + *(volatile uint32_t*)&xtc[c];
                hits[c] = 0;
                xin[c] = 0;
                xpc[c] = 0;
                xtc[c] = 0;
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

// The same for indices small enough that one counter per gene fits in shared memory (S <= kDenseSmemGenes, 16-bit
// counters, two genes per word): the wide rows of small indices — a row sharing random k-mers with a third of all
// genes — are the rule there, and global atomics make them cost 60 % of the scoring time.  A posting is one
// fire-and-forget shared-memory add; no touched list, the finalize pass walks all S counters (a few thousand words per
// thread block pass).  Postings with a multiplicity above one on either side add their corrections to global arrays
// (3 x S words per CTA, all zero between rows), read only for rows that made any.  A column is hit at most once per
// distinct k-mer of the row (twice in the tail-merged group), so 16 bits hold any row with fewer than 65,534 k-mers;
// longer rows are left to the global-memory kernel (the host sends them there).
static const int kDenseSmemThreads = 1024;
static const uint32_t kDenseSmemGenes = 112 * 1024;  // 2 B each: 224 KB
static const uint32_t kDenseSmemMaxK = 65534;

__global__ void __launch_bounds__(kDenseSmemThreads, 1) score_rows_dense_smem_kernel(ScoreArgs a, DenseArgs d) {
    PD_DYNAMIC_SMEM(smem_raw);
    uint32_t* h = reinterpret_cast<uint32_t*>(smem_raw);  // (S + 1) / 2 words
    __shared__ uint32_t s_row;
    __shared__ uint32_t s_corrected;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t words = (d.S + 1) / 2;
    uint32_t* xin = d.acc + (size_t)blockIdx.x * 3 * d.S;  // sum of min(n, m) - 1
    uint32_t* xpc = xin + d.S;                             // sum of m - 1
    uint32_t* xtc = xpc + d.S;                             // sum of n - 1
    unsigned long long pairs = 0;
    const uint32_t n_rows = a.n_rows_dev ? *a.n_rows_dev : a.n_rows;
    for (uint32_t w = tid; w < words; w += kDenseSmemThreads) h[w] = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_row = atomicAdd(a.cursor, 1u);
            s_corrected = 0;
        }
        __syncthreads();
        const uint32_t ri = s_row;
        if (ri >= n_rows) break;
        const RowDesc rw = a.rows[ri];
        RowCtx rc;
        rc.r = rw.gene;
        rc.bh_row = rw.bh_row;
        rc.kr = rw.kr;
        rc.gr = rw.gr;
        // short posting lists: one thread per list; long and huge lists: one warp per list, lanes striding the postings
        for (uint32_t f = rw.fb + tid; f < rw.fm; f += kDenseSmemThreads) {
            const uint2 fw = a.fwd[f];
            const uint32_t gl = fw.y & ~kMulti;
            const uint32_t m = (fw.y & kMulti) ? a.fwd_cnt[f] : 1u;
            for (uint32_t p = 0; p < gl; p++) {
                const uint32_t e = a.post[fw.x + p];
                const uint32_t c = e & ~kMulti;
                atomicAdd(&h[c >> 1], 1u << ((c & 1u) * 16u));
                if ((e & kMulti) || m > 1u) {
                    const uint32_t n = (e & kMulti) ? a.post_cnt[fw.x + p] : 1u;
                    atomicAdd(&xin[c], (n < m ? n : m) - 1u);
                    atomicAdd(&xpc[c], m - 1u);
                    atomicAdd(&xtc[c], n - 1u);
                    s_corrected = 1u;
                }
            }
        }
        for (uint32_t f = rw.fm + warp; f < rw.fe; f += kDenseSmemThreads / 32) {
            const uint2 fw = a.fwd[f];
            const uint32_t m = (fw.y & kMulti) ? a.fwd_cnt[f] : 1u;
            const uint32_t gl = fw.y & ~kMulti;
            for (uint32_t p = lane; p < gl; p += 32) {
                const uint32_t e = a.post[fw.x + p];
                const uint32_t c = e & ~kMulti;
                atomicAdd(&h[c >> 1], 1u << ((c & 1u) * 16u));
                if ((e & kMulti) || m > 1u) {
                    const uint32_t n = (e & kMulti) ? a.post_cnt[fw.x + p] : 1u;
                    atomicAdd(&xin[c], (n < m ? n : m) - 1u);
                    atomicAdd(&xpc[c], m - 1u);
                    atomicAdd(&xtc[c], n - 1u);
                    s_corrected = 1u;
                }
            }
        }
        __threadfence();
        __syncthreads();
        const bool corrected = s_corrected != 0;
        // finalize: every thread walks its words; both genes of a word, then the word is cleared
        for (uint32_t w0 = 0; w0 < words; w0 += kDenseSmemThreads) {
            const uint32_t w = w0 + tid;
            const uint32_t v = w < words ? h[w] : 0u;
            if (v) h[w] = 0;
#pragma unroll
            for (uint32_t half = 0; half < 2; half++) {
                const uint32_t hits = half ? (v >> 16) : (v & 0xFFFFu);
                const uint32_t c = 2 * w + half;
                bool want = false;
                uint32_t in = hits, pc = hits, tc = hits;
                uint2 mc = make_uint2(0u, 0u);
                if (hits) {
                    if (corrected) {
                        const uint32_t xi = *(volatile uint32_t*)&xin[c], xp = *(volatile uint32_t*)&xpc[c], xt = *(volatile uint32_t*)&xtc[c];
                        if (xi | xp | xt) {
                            in += xi;
                            pc += xp;
                            tc += xt;
                            xin[c] = 0;
                            xpc[c] = 0;
                            xtc[c] = 0;
                        }
                    }
                    if (c != rc.r) {  // identity cell dropped (library.cpp:485-487)
                        pairs++;
                        mc = a.meta[c];
                        want = gate(a.k2, pc, tc, rc.kr, mc.x);
                    }
                }
                const unsigned mb = __ballot_sync(0xffffffffu, want);
                if (mb) {  // warp-aggregated append
                    const unsigned leader = __ffs((int)mb) - 1;
                    unsigned long long base = 0;
                    if (lane == leader) base = atomicAdd(a.n_cells, (unsigned long long)__popc(mb));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (want) emit_cell(a, rc, base + __popc(mb & ((1u << lane) - 1u)), c, mc.y, mc.x, in, pc, tc);
                }
            }
        }
        if (corrected) __threadfence();
    }
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, dd);
    if (lane == 0 && pairs) atomicAdd(a.n_pairs, pairs);
}

// ------------------------------------------------------------------------------------------------ row lists

// The rows of one scoring call are sorted on the device by (table level, family key) and turned into descriptors.
//   row_keys_kernel   key_i = level << 62 | fam_key << 31 | i ; level by the row's bound on distinct columns,
//                     min(total_visited, S); counts rows per level and the call's posting / forward-entry totals
//   (radix sort of the keys on bits [31, 64))
//   row_desc_kernel   descriptor j from sorted key j: levels end up contiguous, rows of a family adjacent
// Row i is gene genes[i] (or gene_base + i) and writes best hits to row i of the call's table.
struct ClassifyArgs {
    uint32_t n;
    const uint32_t* genes;
    uint32_t gene_base;
    uint32_t S;
    const unsigned long long* visited;
    const uint32_t* fam_key;
    const uint32_t* fwd_ptr;
    const unsigned long long* cls;   // forward entries per list class, 3 x cls_bits
    uint32_t cls_bits;
    const uint2* meta;
    unsigned long long max_cols[3];
    uint32_t by_cost;                // non-zero: inside a level, heaviest rows first instead of family order (see run_rows)
    uint32_t* counts;                // [3] rows per level
    unsigned long long* stats;       // [0] += postings the rows visit, [1] += their forward entries
};

__global__ void __launch_bounds__(256) row_keys_kernel(ClassifyArgs a, uint64_t* __restrict__ keys) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    unsigned long long lk = 0, fe = 0;
    if (i < a.n) {
        const uint32_t g = a.genes ? a.genes[i] : a.gene_base + i;
        const unsigned long long v = a.visited[g];
        const unsigned long long cols = v < a.S ? v : a.S;
        const unsigned long long level = cols <= a.max_cols[0] ? 0 : (cols <= a.max_cols[1] ? 1 : 2);
        const unsigned long long mid = a.by_cost ? 0x7FFFFFFFull - (v < 0x7FFFFFFFull ? v : 0x7FFFFFFFull)
                                                 : (unsigned long long)(a.fam_key[g] & 0x7FFFFFFFu);
        keys[i] = (level << 62) | (mid << 31) | i;
        atomicAdd(&a.counts[level], 1u);
        lk = v;
        fe = a.fwd_ptr[g + 1] - a.fwd_ptr[g];
    }
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) {
        lk += __shfl_xor_sync(0xffffffffu, lk, dd);
        fe += __shfl_xor_sync(0xffffffffu, fe, dd);
    }
    if ((threadIdx.x & 31) == 0 && fe) {
        atomicAdd(&a.stats[0], lk);
        atomicAdd(&a.stats[1], fe);
    }
}

__global__ void __launch_bounds__(256) row_desc_kernel(ClassifyArgs a, const uint64_t* __restrict__ keys, RowDesc* __restrict__ rows) {
    const uint32_t j = blockIdx.x * 256u + threadIdx.x;
    if (j >= a.n) return;
    const uint32_t i = (uint32_t)(keys[j] & 0x7FFFFFFFull);
    const uint32_t g = a.genes ? a.genes[i] : a.gene_base + i;
    const unsigned long long cl = a.cls[g];
    const unsigned long long m = (1ull << a.cls_bits) - 1ull;
    const uint2 mg = a.meta[g];
    RowDesc d;
    d.gene = g;
    d.bh_row = i;
    d.fb = a.fwd_ptr[g];
    d.fe = a.fwd_ptr[g + 1];
    d.fm = d.fb + (uint32_t)(cl & m);
    d.fh = d.fe - (uint32_t)((cl >> (2 * a.cls_bits)) & m);
    d.kr = mg.x;
    d.gr = mg.y;
    rows[j] = d;
}

}  // namespace sk
}  // namespace pd
