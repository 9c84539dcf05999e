// Host orchestration of the B200 Pangenes engine: index build (preprocessSequences, reference
// ig/native/library.cpp:189-371) and scoring calls (computeScores, library.cpp:409-604).  All arithmetic runs in
// the kernels of index_kernels.cuh / score_kernels.cuh / prims.cuh; this file sizes buffers, orders launches on one
// stream per context and moves results.  There is no CPU implementation of any step.
#include "engine.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <map>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "filter_kernels.cuh"
#include "index_kernels.cuh"
#include "prims.cuh"
#include "score_kernels.cuh"
#include "sort_kernels.cuh"

namespace pd {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
const std::string& last_error() { return g_last_error; }

namespace {

inline unsigned blocks_for(uint64_t n, unsigned per = 256) { return (unsigned)((n + per - 1) / per); }
inline int bits_for(uint64_t v) {  // bits needed to represent values 0..v
    int b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}

struct Timer {
    rt::event_t a, b;
    rt::stream_t st;
    explicit Timer(rt::stream_t s) : st(s) {
        a = rt::event_create();
        b = rt::event_create();
    }
    ~Timer() {
        rt::event_destroy(a);
        rt::event_destroy(b);
    }
    void start() { rt::event_record(a, st); }
    void stop() { rt::event_record(b, st); }
    double ms() { return rt::event_ms(a, b); }
};

}  // namespace

// ------------------------------------------------------------------------------------------------ contexts

struct ScoreContext {
    Index* ix = nullptr;
    rt::stream_t st;
    rt::event_t ev_call0, ev_call1, ev_k0, ev_k1;
    // cell buffers (device, SoA in the layout of Scores.java) and their host mirrors
    uint64_t cap = 0;
    rt::DevBuf<float> d_score, d_perc, d_trperc;
    rt::DevBuf<int32_t> d_row, d_col, d_g1, d_g2;
    rt::DevBuf<uint32_t> d_bh, d_colmax;
    rt::DevBuf<unsigned long long> d_counters;  // 8
    rt::DevBuf<uint32_t> d_cursors;             // 16
    rt::DevBuf<sk::RowDesc> d_rows, d_ovf;
    rt::DevBuf<uint32_t> d_dense, d_xtab, d_sorttmp;
    // per-genome network filter (genome_edges)
    rt::DevBuf<uint32_t> d_imax, d_rowthr, d_esrc, d_edst;
    rt::DevBuf<float> d_escore;
    rt::DevBuf<uint8_t> d_flag;
    rt::PinBuf<uint32_t> h_esrc, h_edst;
    rt::PinBuf<float> h_escore;
    uint64_t edge_cap = 0;
    bool g1_bhrow = false;
    rt::DevBuf<uint64_t> d_rowkeys;
    uint32_t xtab_ctas = 0;
    uint32_t dense_S = 0;  // gene count the dense accumulators are laid out (and clean) for
    uint32_t dense_layout = 0;  // 1: global-memory kernel, 2: shared-memory kernel (corrections only)
    rt::PinBuf<unsigned long long> h_counters;
    rt::PinBuf<sk::RowDesc> h_rows;
    rt::PinBuf<float> h_score, h_perc, h_trperc, h_bh, h_colmax;
    rt::PinBuf<int32_t> h_row, h_col, h_g1, h_g2, h_map;
    pd_score_stats stats;
    uint32_t map_r0 = 0, map_rows = 0;  // rows currently set in h_map
    bool host_cells = false;            // the pinned cell arrays follow `cap`
    unsigned long long seq = 0;         // last completion number asked of the stream (publish / wait)

    explicit ScoreContext(Index* i) : ix(i) {
        st = rt::stream_create();
        ev_call0 = rt::event_create();
        ev_call1 = rt::event_create();
        ev_k0 = rt::event_create();
        ev_k1 = rt::event_create();
        d_counters.alloc(8);
        d_cursors.alloc(16);
        h_counters.ensure(16);
        memset(h_counters.p, 0, 16 * sizeof(unsigned long long));
        memset(&stats, 0, sizeof(stats));
    }
    ~ScoreContext() {
        rt::event_destroy(ev_call0);
        rt::event_destroy(ev_call1);
        rt::event_destroy(ev_k0);
        rt::event_destroy(ev_k1);
        rt::stream_destroy(st);
    }
    void ensure_cells(uint64_t n) {
        if (n <= cap) return;
        d_score.alloc(n); d_perc.alloc(n); d_trperc.alloc(n);
        d_row.alloc(n); d_col.alloc(n); d_g1.alloc(n); d_g2.alloc(n);
        cap = n;
        if (host_cells) ensure_host_cells();
    }
    // pinned images of the cell arrays, sized with the device arrays so that they are reallocated as rarely
    void ensure_host_cells() {
        host_cells = true;
        const size_t n = std::max<uint64_t>(cap, 1);
        h_score.ensure(n); h_perc.ensure(n); h_trperc.ensure(n);
        h_row.ensure(n); h_col.ensure(n); h_g1.ensure(n); h_g2.ensure(n);
    }
};

// PD_TRACE=1: host wall-clock per phase of the per-genome calls, summed over calls and printed when the index dies
// fills a pinned result array from the host with streaming stores (no read-for-ownership of lines that are only written)
static void fill_i32_stream(int32_t* p, uint64_t n, int32_t v) {
#if defined(__SSE2__) && !defined(PD_EMU)
    uint64_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(p + i) & 15u)) p[i++] = v;
    const __m128i vv = _mm_set1_epi32(v);
    for (; i + 16 <= n; i += 16) {
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i), vv);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i + 4), vv);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i + 8), vv);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + i + 12), vv);
    }
    _mm_sfence();
    for (; i < n; i++) p[i] = v;
#else
    std::fill(p, p + n, v);
#endif
}

// takes a stage token unless the stage is switched off
struct StageLock {
    std::unique_lock<std::mutex> lk;
    StageLock(std::mutex& m, bool on) : lk(m, std::defer_lock) { if (on) lk.lock(); }
    void unlock() { if (lk.owns_lock()) lk.unlock(); }
};

namespace trace {
enum { kAcquire, kSortLaunch, kLaunch, kLevelsWait, kRetryWait, kTokenWait, kCopyEnqueue, kCopyWait, kTail, kCalls, kN };
static std::atomic<uint64_t> ns[kN];
static const char* const names[kN] = {"acquire", "sort_launch", "launch", "levels_wait", "retry_wait", "token_wait", "copy_enqueue", "copy_wait", "tail", "calls"};
static bool on() {
    static const bool v = getenv("PD_TRACE") != nullptr;
    return v;
}
struct Clock {
    std::chrono::steady_clock::time_point t;
    Clock() : t(std::chrono::steady_clock::now()) {}
    void lap(int what) {
        if (!on()) return;
        const auto n = std::chrono::steady_clock::now();
        ns[what] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(n - t).count();
        t = n;
    }
};
static void report() {
    if (!on()) return;
    fprintf(stderr, "[pd trace] %llu calls;", (unsigned long long)ns[kCalls].load());
    for (int i = 0; i < kCalls; i++) fprintf(stderr, " %s %.1f ms;", names[i], (double)ns[i].exchange(0) * 1e-6);
    fprintf(stderr, "\n");
    ns[kCalls] = 0;
#ifndef PD_EMU
    for (int d = 0; d < 16; d++) {  // the block cache since the last report: a steady state should show no cudaMalloc at all
        uint64_t hits, misses, bytes, mns;
        rt::cache_stats(d, &hits, &misses, &bytes, &mns);
        if (hits | misses)
            fprintf(stderr, "[pd trace] device %d block cache: %llu hits, %llu cudaMalloc (%.1f MB, %.1f ms)\n", d, (unsigned long long)hits,
                    (unsigned long long)misses, (double)bytes / 1048576.0, (double)mns * 1e-6);
    }
#endif
}
}  // namespace trace

// Score contexts (stream, events, per-call device and pinned result buffers) hold nothing of the index they serve, so
// they outlive it: a dying index parks its contexts here and the next index on that device takes them over with their
// buffers already grown.  Without this every new index pays the cudaMalloc / cudaMallocHost warm-up of its first calls
// again (measured: hundreds of milliseconds over the first few hundred calls).  Parked contexts are never destroyed at
// process exit: the CUDA runtime may be gone by then.
namespace {
struct ContextPool {
    std::mutex mu;
    std::vector<ScoreContext*> parked[16];
};
ContextPool& context_pool() {
    static ContextPool* p = new ContextPool;
    return *p;
}
const size_t kParkedPerDevice = 8;
}  // namespace

Index::~Index() {
    delete shard;
    trace::report();
    ContextPool& pool = context_pool();
    for (ScoreContext* c : all_ctx) {
        c->ix = nullptr;
        // a context the caller still holds (pd_scores / pd_edges not released) is not parked: its buffers are in use and
        // its flat_map rows are still set; it is deleted when the caller releases it (release_context)
        if (std::find(free_ctx.begin(), free_ctx.end(), c) == free_ctx.end()) continue;
        std::unique_lock<std::mutex> lk(pool.mu);
        std::vector<ScoreContext*>& v = pool.parked[device & 15];
        if (v.size() < kParkedPerDevice) {
            v.push_back(c);
        } else {
            lk.unlock();
            delete c;
        }
    }
}

// drops everything the library keeps between indices: parked score contexts, cached device and pinned host blocks
void trim_memory() {
    std::vector<ScoreContext*> drop;
    {
        ContextPool& pool = context_pool();
        std::lock_guard<std::mutex> lk(pool.mu);
        for (auto& v : pool.parked) {
            drop.insert(drop.end(), v.begin(), v.end());
            v.clear();
        }
    }
    for (ScoreContext* c : drop) delete c;
    rt::trim_all();
}

ScoreContext* Index::acquire() {
    std::unique_lock<std::mutex> lk(mu);
    const size_t limit = opt.contexts > 0 ? (size_t)opt.contexts : 4;
    for (;;) {
        if (!free_ctx.empty()) {
            ScoreContext* c = free_ctx.back();
            free_ctx.pop_back();
            return c;
        }
        if (all_ctx.size() < limit) {
            ScoreContext* c = nullptr;
            {
                ContextPool& pool = context_pool();
                std::lock_guard<std::mutex> pk(pool.mu);
                std::vector<ScoreContext*>& v = pool.parked[device & 15];
                if (!v.empty()) {
                    c = v.back();
                    v.pop_back();
                }
            }
            if (c) c->ix = this;
            else c = new ScoreContext(this);
            all_ctx.push_back(c);
            return c;
        }
        cv.wait(lk);
    }
}

void Index::release_context(ScoreContext* c) {
    if (!c) return;
    if (c->ix) c->ix->release(c);
    else delete c;   // its index is gone (pd_free before pd_scores_release): nothing to hand it back to
}

void Index::release(ScoreContext* c) {
    for (uint32_t i = 0; i < c->map_rows; i++) c->h_map.p[genome_rows[c->map_r0 + i]] = INT32_MAX;
    c->map_rows = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        free_ctx.push_back(c);
    }
    cv.notify_one();
}

// ------------------------------------------------------------------------------------------------ build

void Index::build(const uint8_t* residues, bool residues_on_device, const uint64_t* offsets, const uint32_t* genome_ids,
                  uint32_t S, int32_t k, const pd_options* o) {
    memset(&info, 0, sizeof(info));
    memset(&opt, 0, sizeof(opt));
    opt.device = -1;
    if (o) opt = *o;
    if (k <= 0) throw Error(PD_ERR_INVALID, "K value must be greater than 0.");  // library.cpp:90-93
    // pd_options.table_on_device: `offsets` and `genome_ids` are device pointers too (with device residues): a host that keeps
    // the whole input resident in HBM between builds pays no copy of the 12 S bytes of the gene table either
    const bool tab_dev = opt.table_on_device != 0 && residues_on_device;
    if (!offsets || (!genome_ids && S)) throw Error(PD_ERR_INVALID, "null input pointer");
    if (rt::device_count() <= 0) throw Error(PD_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU path");
    if (opt.device >= 0) rt::set_device(opt.device);
    device = rt::current_device();
    if (const char* e = getenv("PD_STAGE_TOKENS")) stage_tokens = atoi(e);
    sms = rt::sm_count();
    smem_optin = rt::max_optin_smem();

    uint64_t off_ends[2] = {0, 0};   // offsets[0], offsets[S]
    if (tab_dev) {
        rt::stream_t s0 = rt::stream_create();
        rt::d2h(&off_ends[0], offsets, sizeof(uint64_t), s0);
        rt::d2h(&off_ends[1], offsets + S, sizeof(uint64_t), s0);
        rt::sync(s0);
        rt::stream_destroy(s0);
    } else {
        off_ends[0] = offsets[0];
        off_ends[1] = offsets[S];
    }
    if (!residues && S && off_ends[1] != off_ends[0]) throw Error(PD_ERR_INVALID, "null input pointer");
    if (off_ends[0] != 0) throw Error(PD_ERR_INVALID, "offsets[0] must be 0");
    if (S >= 0x7FFFFFFFu) throw Error(PD_ERR_UNSUPPORTED, "2^31 or more genes");
    const uint64_t total = off_ends[1];
    info.S = S;
    info.k = k;
    thr = 1.0f / (2.0f * (float)k);

    struct StreamGuard {   // every exit path, the throwing ones too, gives the stream back
        rt::stream_t s;
        StreamGuard() : s(rt::stream_create()) {}
        ~StreamGuard() { rt::stream_destroy(s); }
    } st_guard;
    rt::stream_t st = st_guard.s;
    uint64_t launches = 0;
    Timer t_all(st), t_h2d(st), t_hist(st), t_enc(st), t_sort(st), t_grp(st), t_fwd(st);
    t_all.start();

    // ---- residues and gene table to HBM; per gene: k-mer count, (count, genome), input checks (library.cpp:242-262)
    rt::DevBuf<uint8_t> d_res_own;
    const uint8_t* d_res = residues;
    t_h2d.start();
    if (!residues_on_device) {
        d_res_own.alloc(total + 16);
        rt::h2d(d_res_own.p, residues, total, st);
        d_res = d_res_own.p;
    }
    struct TableRef {   // the gene table on the device: a copy made here, or the caller's own arrays
        rt::DevBuf<uint64_t> off_own;
        rt::DevBuf<uint32_t> gid_own;
        const uint64_t* off = nullptr;
        const uint32_t* gid = nullptr;
    } tab;
    struct { const uint64_t* p; } d_gene_off;
    struct { const uint32_t* p; } d_gid;
    if (tab_dev) {
        d_gene_off.p = offsets;
        d_gid.p = genome_ids;
    } else {
        tab.off_own.alloc((size_t)S + 1);
        tab.gid_own.alloc(std::max<size_t>(S, 1));
        d_gene_off.p = tab.off_own.p;
        d_gid.p = tab.gid_own.p;
    }
    rt::DevBuf<uint32_t> d_kseq((size_t)S + 1), d_key_off((size_t)S + 1), d_flags(4), d_total(4);
    rt::DevBuf<unsigned long long> d_n64(2);
    rt::zero(d_n64.p, 2 * sizeof(unsigned long long), st);
    if (!tab_dev) {
        rt::h2d(tab.off_own.p, offsets, sizeof(uint64_t) * ((size_t)S + 1), st);
        rt::h2d(tab.gid_own.p, genome_ids, sizeof(uint32_t) * S, st);
    }
    rt::zero(d_flags.p, 4 * sizeof(uint32_t), st);
    meta.alloc(std::max<size_t>(S, 1));
    PD_LAUNCH(ik::gene_meta_kernel, blocks_for((uint64_t)S + 1), 256, 0, st, d_gene_off.p, d_gid.p, S, (int)k,
              d_kseq.p, meta.p, d_flags.p, d_n64.p);
    launches++;
    rt::DevBuf<uint32_t> scratch(prims::scan_tmp_words((uint64_t)S + 1) + 16);
    prims::exclusive_scan_u32(d_kseq.p, d_key_off.p, (uint64_t)S + 1, scratch.p, nullptr, st, &launches);
    uint32_t h_flags[4] = {0, 0, 0, 0}, h_N = 0;
    rt::d2h(h_flags, d_flags.p, sizeof(h_flags), st);
    rt::d2h(&h_N, d_key_off.p + S, sizeof(uint32_t), st);
    unsigned long long h_n64 = 0;   // the exact k-mer total: the 32-bit scan above wraps at 2^32
    rt::d2h(&h_n64, d_n64.p, sizeof(h_n64), st);
    t_h2d.stop();

    // ---- alphabet (library.cpp:216-230, 96-100)
    t_hist.start();
    rt::DevBuf<unsigned long long> d_hist(256);
    rt::zero(d_hist.p, 256 * sizeof(unsigned long long), st);
    if (total) {
        unsigned grid = std::min<uint64_t>((uint64_t)sms * 8, (total + 4095) / 4096);
        PD_LAUNCH(ik::byte_hist_kernel, std::max(1u, grid), 256, 0, st, d_res, total, d_hist.p);
        launches++;
    }
    unsigned long long h_hist[256];
    rt::d2h(h_hist, d_hist.p, sizeof(h_hist), st);
    rt::sync(st);
    t_hist.stop();
    if (h_flags[0] & 1u) throw Error(PD_ERR_INVALID, "offsets must be ascending");
    if (h_flags[0] & 2u) throw Error(PD_ERR_UNSUPPORTED, "a gene of 2^20 or more residues");
    if (h_n64 >= (1ull << 31)) throw Error(PD_ERR_UNSUPPORTED, "2^31 or more k-mers (the reference's own int index limit)");
    const uint64_t N = h_N;
    const uint32_t G = S ? h_flags[1] + 1 : 0;  // library.cpp:242
    info.G = G;
    info.N = N;
    ik::ValTable vt;
    memset(&vt, 0, sizeof(vt));
    uint32_t base = 0;
    for (int b = 0; b < 256; b++)
        if (h_hist[b]) vt.v[b] = (uint8_t)base++;
    info.base = base;

    // base^k must be exactly representable; the reference would switch to its Rabin hash (library.cpp:103-119)
    unsigned __int128 pw = 1;
    for (int i = 0; i < k; i++) {
        pw *= (base ? base : 1);
        if (pw >> 63) throw Error(PD_ERR_UNSUPPORTED, "base^k >= 2^63: the reference's hash fallback is out of scope");
    }
    const int rank_bits = std::max(1, bits_for((uint64_t)pw - 1));
    const int seq_bits = std::max(1, bits_for(S ? (uint64_t)S - 1 : 0));
    if (rank_bits + seq_bits > 63) throw Error(PD_ERR_UNSUPPORTED, "k-mer rank and gene id do not fit one 64-bit sort key");
    info.rank_bits = rank_bits;
    info.seq_bits = seq_bits;

    fwd_ptr.alloc((size_t)S + 1);
    cls.alloc(std::max<size_t>(S, 1));
    d_visited.alloc(std::max<size_t>(S, 1));
    fam_key.alloc(std::max<size_t>(S, 1));
    rt::fill_byte(fam_key.p, 0x7F, sizeof(uint32_t) * S, st);
    rt::zero(fwd_ptr.p, sizeof(uint32_t) * ((size_t)S + 1), st);
    rt::zero(cls.p, sizeof(unsigned long long) * S, st);
    rt::zero(d_visited.p, sizeof(unsigned long long) * S, st);
    uint32_t U = 0, n_groups = 0, R = 0;
    unsigned long long lookups = 0;

    if (shard_world > 1) {
        // ================================================================ sharded build: this rank's slice of the rank space
        if (opt.keep_sorted) throw Error(PD_ERR_INVALID, "keep_sorted is not available in a sharded build");
        if (N == 0 || S == 0) throw Error(PD_ERR_UNSUPPORTED, "sharded build of an empty input");
        if (shard_world > (uint32_t)sortk::kMaxSlices) throw Error(PD_ERR_UNSUPPORTED, "sharded build: more than 32 ranks");
        shard = new Shard;
        Shard& sh = *shard;
        // Slice bounds in the rank space, identical on every rank: cut at multiples of base^(k-2) (a pair of leading
        // letters) so that every slice holds about 1 / world of the k-mers if letters were independent — balance of the
        // sort only, any cut is correct.  Counts of leading pairs are estimated from the alphabet histogram.
        uint64_t cuts[sortk::kMaxSlices + 1];
        {
            const int lead = k >= 2 ? 2 : 1;
            std::vector<double> f(base);
            {
                uint32_t v = 0;
                for (int b = 0; b < 256; b++)
                    if (h_hist[b]) f[v++] = (double)h_hist[b] / (double)total;
            }
            const uint64_t unit = (uint64_t)pw / (lead == 2 ? (uint64_t)base * base : (uint64_t)base);
            const uint64_t bins = lead == 2 ? (uint64_t)base * base : base;
            uint64_t b = 0;
            double cum = 0;
            cuts[0] = 0;
            for (uint32_t r = 1; r < shard_world; r++) {   // first bin of slice r
                const double want = (double)r / (double)shard_world;
                while (b < bins) {
                    const double p = lead == 2 ? f[b / base] * f[b % base] : f[b];
                    if (cum + p / 2 >= want) break;
                    cum += p;
                    b++;
                }
                cuts[r] = b * unit;
            }
            cuts[shard_world] = (uint64_t)pw;
        }
        // ---- this rank's SHARE of the genes (about 1 / world of the residues): its k-mers are made once, grouped by
        // destination slice in one stable pass, and handed to the caller for the all-to-all
        uint32_t h_pos[2] = {0, 0};
        {
            const uint64_t t0 = total / shard_world * shard_rank + (total % shard_world) * shard_rank / shard_world;
            const uint64_t t1 = shard_rank + 1 == shard_world ? total : total / shard_world * (shard_rank + 1) + (total % shard_world) * (shard_rank + 1) / shard_world;
            // first genes at or after residues t0 / t1 (lower bound over the offsets), and the key positions they start at
            rt::DevBuf<uint32_t> d_pos(4);
            PD_LAUNCH(ik::share_bounds_kernel, 1, 32, 0, st, d_gene_off.p, (const uint32_t*)d_key_off.p, S, shard_rank == 0 ? 0ull : t0,
                      shard_rank + 1 == shard_world ? ~0ull : t1, d_pos.p);
            launches++;
            rt::d2h(h_pos, d_pos.p, sizeof(h_pos), st);
            rt::sync(st);
        }
        const uint64_t pos_lo = h_pos[0], pos_hi = h_pos[1], n_share = pos_hi - pos_lo;
        t_enc.start();
        const uint32_t stiles = sortk::tiles_of(N);
        sh.keys_a.alloc(std::max<uint64_t>(n_share, 1));
        rt::DevBuf<uint32_t> tile_gene((size_t)stiles + 1), sort_ctl((size_t)2 * sortk::kRadix + 32);
        uint32_t* d_hist = sort_ctl.p;
        uint32_t* d_bins = d_hist + sortk::kRadix;
        uint32_t* d_tot = d_bins + sortk::kRadix;
        rt::zero(sort_ctl.p, sort_ctl.bytes(), st);
        PD_LAUNCH(sortk::tile_gene_kernel, blocks_for(std::max<uint32_t>(S, 1)), 256, 0, st, (const uint32_t*)d_key_off.p, S, stiles, tile_gene.p);
        launches++;
        memset(sh.send_counts, 0, sizeof(sh.send_counts));
        if (n_share) {
            sortk::EncodeSrc es;
            memset(&es, 0, sizeof(es));
            es.res = d_res; es.gene_off = d_gene_off.p; es.key_off = d_key_off.p; es.tile_gene = tile_gene.p;
            es.S = S; es.k = (int)k; es.base = base; es.mult = (uint64_t)(pw / (base ? base : 1)); es.seq_bits = seq_bits;
            es.lo = 0; es.hi = 1ull << 63; es.N = N; es.pos_lo = pos_lo; es.pos_hi = pos_hi; es.vt = vt;
            es.n_slices = shard_world;
            for (uint32_t r = 0; r <= shard_world; r++) es.cuts[r] = cuts[r];
            const bool r32 = rank_bits <= 32;
            const uint32_t tile0 = (uint32_t)(pos_lo / sortk::kTile), tile1 = (uint32_t)((pos_hi + sortk::kTile - 1) / sortk::kTile);
            const uint32_t ntiles = tile1 - tile0;
            rt::DevBuf<uint32_t> status((size_t)ntiles * sortk::kRadix);
            rt::zero(status.p, status.bytes(), st);
            const unsigned hgrid = std::min<uint32_t>(ntiles, (uint32_t)sms * 8);
            if (r32) PD_LAUNCH((sortk::kmer_hist_kernel<uint32_t, true>), hgrid, sortk::kThreads, 0, st, es, tile0, ntiles, 1, d_hist);
            else PD_LAUNCH((sortk::kmer_hist_kernel<uint64_t, true>), hgrid, sortk::kThreads, 0, st, es, tile0, ntiles, 1, d_hist);
            PD_LAUNCH(sortk::digit_bins_kernel, 1, sortk::kRadix, 0, st, (const uint32_t*)d_hist, 1, d_bins, d_tot);
            sortk::SweepArgs sa;
            sa.keys = nullptr; sa.out = sh.keys_a.p; sa.n = N; sa.shift = 0; sa.bins = d_bins; sa.status = status.p; sa.counter = d_tot + 8; sa.tile0 = tile0;
            const size_t smem = sortk::sweep_smem_bytes();
            void (*kfn)(sortk::SweepArgs, sortk::EncodeSrc) = r32 ? sortk::onesweep_kernel<2, uint32_t> : sortk::onesweep_kernel<2, uint64_t>;
            rt::allow_smem(kfn, smem);
            PD_LAUNCH(kfn, ntiles, sortk::kThreads, smem, st, sa, es);
            launches += 3;
            uint32_t h_cnt[sortk::kMaxSlices];
            rt::d2h(h_cnt, d_hist, sizeof(uint32_t) * shard_world, st);
            t_enc.stop();
            t_all.stop();
            rt::sync(st);   // (also: status, sort_ctl, tile_gene die here)
            for (uint32_t r = 0; r < shard_world; r++) sh.send_counts[r] = h_cnt[r];
        } else {
            t_enc.stop();
            t_all.stop();
            rt::sync(st);
        }
        sh.n_share = n_share;
        sh.rank_bits = rank_bits;
        sh.launches = launches;
        sh.ms = t_all.ms();
        info.build_ms[0] = total ? t_hist.ms() : 0;
        info.build_ms[1] = t_enc.ms();
        info.build_ms[5] = sh.ms;
        info.build_ms[6] = t_h2d.ms();
        info.build_ms[7] = (double)launches;
        info.max_kseq = h_flags[2];
        return;
    }

    if (N > 0) {
        // ---- k-mer keys + sort by rank in one sweep per digit (library.cpp:234-265, 270-278): the keys are made from the
        // residues twice (digit counts, first pass) instead of being written and re-read in gene order; stability keeps
        // genes ascending inside a rank
        rt::DevBuf<uint64_t> keys_a(N), keys_b(N);
        const uint32_t stiles = sortk::tiles_of(N);
        const int passes = sortk::passes_of(rank_bits);
        rt::DevBuf<uint32_t> tile_gene((size_t)stiles + 1), sort_ctl((size_t)2 * passes * sortk::kRadix + 32);
        rt::DevBuf<uint32_t> sort_status((size_t)passes * stiles * sortk::kRadix);
        uint64_t* sorted = nullptr;
        {
            t_enc.start();
            uint32_t* d_hist = sort_ctl.p;
            uint32_t* d_bins = d_hist + passes * sortk::kRadix;
            uint32_t* d_tot = d_bins + passes * sortk::kRadix;  // [0] keys kept, [8 + p] tile tickets
            rt::zero(sort_ctl.p, sort_ctl.bytes(), st);
            rt::zero(sort_status.p, sort_status.bytes(), st);
            PD_LAUNCH(sortk::tile_gene_kernel, blocks_for(std::max<uint32_t>(S, 1)), 256, 0, st, (const uint32_t*)d_key_off.p, S, stiles, tile_gene.p);
            sortk::EncodeSrc es;
            memset(&es, 0, sizeof(es));
            es.res = d_res;
            es.gene_off = d_gene_off.p;
            es.key_off = d_key_off.p;
            es.tile_gene = tile_gene.p;
            es.S = S;
            es.k = (int)k;
            es.base = base;
            es.mult = (uint64_t)(pw / (base ? base : 1));
            es.seq_bits = seq_bits;
            es.lo = 0;
            es.hi = 1ull << 63;
            es.N = N;
            es.pos_lo = 0;
            es.pos_hi = N;
            es.n_slices = 1;
            es.vt = vt;
            const bool r32 = rank_bits <= 32;
            const unsigned hgrid = std::min<uint32_t>(stiles, (uint32_t)sms * 8);
            if (r32) PD_LAUNCH((sortk::kmer_hist_kernel<uint32_t, false>), hgrid, sortk::kThreads, 0, st, es, 0u, stiles, passes, d_hist);
            else PD_LAUNCH((sortk::kmer_hist_kernel<uint64_t, false>), hgrid, sortk::kThreads, 0, st, es, 0u, stiles, passes, d_hist);
            PD_LAUNCH(sortk::digit_bins_kernel, 1, sortk::kRadix, 0, st, (const uint32_t*)d_hist, passes, d_bins, d_tot);
            launches += 3;
            t_enc.stop();

            t_sort.start();
            const size_t smem = sortk::sweep_smem_bytes();
            uint64_t* src = nullptr;
            uint64_t* dst = keys_a.p;
            for (int p = 0; p < passes; p++) {
                sortk::SweepArgs sa;
                sa.keys = src;
                sa.out = dst;
                sa.n = N;
                sa.shift = seq_bits + 8 * p;
                sa.bins = d_bins + p * sortk::kRadix;
                sa.status = sort_status.p + (size_t)p * stiles * sortk::kRadix;
                sa.counter = d_tot + 8 + p;
                sa.tile0 = 0;
                void (*kfn)(sortk::SweepArgs, sortk::EncodeSrc) = sortk::onesweep_kernel<0, uint32_t>;
                if (p == 0) kfn = r32 ? sortk::onesweep_kernel<1, uint32_t> : sortk::onesweep_kernel<1, uint64_t>;
                rt::allow_smem(kfn, smem);
                PD_LAUNCH(kfn, stiles, sortk::kThreads, smem, st, sa, es);
                launches++;
                src = dst;
                dst = (dst == keys_a.p) ? keys_b.p : keys_a.p;
            }
            sorted = src;
            t_sort.stop();
        }

        // ---- count dedup (library.cpp:280-287) and rank groups incl. the tail merge (library.cpp:297-306)
        t_grp.start();
        const uint32_t tiles = (uint32_t)((N + ik::kEntTile - 1) / ik::kEntTile);
        rt::DevBuf<uint32_t> tile_h((size_t)tiles + 1), tile_g((size_t)tiles + 1), d_spur(4);
        rt::zero(d_spur.p, 4 * sizeof(uint32_t), st);
        PD_LAUNCH(ik::entry_count_kernel, tiles, ik::kEntThreads, 0, st, (const uint64_t*)sorted, N, seq_bits, tile_h.p, tile_g.p);
        launches++;
        scratch.ensure(prims::scan_tmp_words(tiles) + 16);
        prims::exclusive_scan_u32(tile_h.p, tile_h.p, tiles, scratch.p, d_total.p, st, &launches);
        prims::exclusive_scan_u32(tile_g.p, tile_g.p, tiles, scratch.p, d_total.p + 1, st, &launches);
        uint32_t h_tot[2] = {0, 0};
        rt::d2h(h_tot, d_total.p, sizeof(h_tot), st);
        rt::sync(st);
        U = h_tot[0];
        const uint32_t g_counted = h_tot[1];
        post.alloc(U);
        post_cnt.alloc(U);
        ent_gid.alloc(U);
        grp_head.alloc((size_t)g_counted + 1);
        if (opt.keep_sorted) ent_rank.alloc(U);
        ik::ShardOut so;
        memset(&so, 0, sizeof(so));
        so.tail_merge = 1;
        PD_LAUNCH(ik::entry_apply_kernel, tiles, ik::kEntThreads, 0, st, (const uint64_t*)sorted, N, seq_bits, (const uint32_t*)tile_h.p,
                  (const uint32_t*)tile_g.p, U, post.p, post_cnt.p, ent_gid.p, grp_head.p, opt.keep_sorted ? ent_rank.p : (uint64_t*)nullptr,
                  d_spur.p, so);
        PD_LAUNCH(ik::group_tail_kernel, 1, 32, 0, st, grp_head.p, g_counted, (const uint32_t*)d_spur.p, U);
        launches += 2;
        t_grp.stop();

        // ---- forward lists by counting sort on the gene, and the cost model (library.cpp:308-330)
        t_fwd.start();
        rt::DevBuf<unsigned long long> d_lookups(2);
        rt::DevBuf<uint32_t> gene_tot((size_t)S + 1), cur3(std::max<size_t>((size_t)3 * S, 1));
        rt::zero(d_lookups.p, 2 * sizeof(unsigned long long), st);
        PD_LAUNCH(ik::fwd_count_kernel, blocks_for(U, 256 * ik::kFwdItems), 256, 0, st, (const uint32_t*)post.p, (const uint32_t*)ent_gid.p,
                  (const uint32_t*)grp_head.p, U, sk::kShortList, sk::kHugeList, cls.p, (unsigned long long*)nullptr);
        PD_LAUNCH(ik::fwd_totals_kernel, blocks_for((uint64_t)S + 1), 256, 0, st, (const unsigned long long*)cls.p, S, 0u, S, gene_tot.p);
        launches += 2;
        scratch.ensure(prims::scan_tmp_words((uint64_t)S + 1) + 16);
        prims::exclusive_scan_u32(gene_tot.p, fwd_ptr.p, (uint64_t)S + 1, scratch.p, nullptr, st, &launches);
        uint32_t h_R = 0;
        rt::d2h(&h_R, fwd_ptr.p + S, sizeof(uint32_t), st);
        rt::sync(st);
        R = h_R;
        fwd.alloc(std::max<size_t>(R, 1));
        fwd_cnt.alloc(std::max<size_t>(R, 1));
        if (R) {
            // transpose in two steps: records partitioned by gene bucket, then placed bucket after bucket
            const uint32_t bshift = (uint32_t)std::max(0, seq_bits - 8);
            rt::DevBuf<uint4> records(R);
            rt::DevBuf<uint32_t> bucket_cur(ik::kMaxBuckets);
            rt::zero(bucket_cur.p, sizeof(uint32_t) * ik::kMaxBuckets, st);
            PD_LAUNCH(ik::fwd_partition_kernel, blocks_for(U, ik::kPartTile), ik::kPartThreads, 0, st, (const uint32_t*)post.p,
                      (const uint32_t*)post_cnt.p, (const uint32_t*)ent_gid.p, (const uint32_t*)grp_head.p, U, S, sk::kShortList, sk::kHugeList,
                      bshift, (const uint32_t*)fwd_ptr.p, bucket_cur.p, records.p, 0u, S, (const uint32_t*)nullptr, (const uint32_t*)nullptr);
            PD_LAUNCH(ik::fwd_cursor_init_kernel, blocks_for(S), 256, 0, st, (const unsigned long long*)cls.p, (const uint32_t*)fwd_ptr.p, S,
                      cur3.p);
            PD_LAUNCH(ik::fwd_place_kernel, blocks_for(R, 256 * ik::kFwdItems), 256, 0, st, (const uint4*)records.p, R, cur3.p, fwd.p, fwd_cnt.p);
            launches += 3;
        }
        PD_LAUNCH(ik::gene_visited_kernel, blocks_for((uint64_t)S * 32), 256, 0, st, (const uint2*)fwd.p, (const uint32_t*)fwd_ptr.p, 0u, S,
                  sk::kShortList, d_visited.p, fam_key.p, d_lookups.p);
        launches += 1;
        uint32_t h_spur = 0;
        rt::d2h(&h_spur, d_spur.p, sizeof(uint32_t), st);
        rt::d2h(&lookups, d_lookups.p, sizeof(unsigned long long), st);
        t_fwd.stop();
        rt::sync(st);
        n_groups = g_counted - h_spur;
        if (!opt.keep_sorted) {
            ent_gid.release();
            grp_head.release();
        }
    } else {
        post.alloc(1);
        post_cnt.alloc(1);
        fwd.alloc(1);
        fwd_cnt.alloc(1);
    }
    t_all.stop();
    rt::sync(st);
    info.U = U;
    info.groups = n_groups;
    info.R = R;
    info.lookups = lookups;
    info.max_kseq = h_flags[2];
    info.build_ms[0] = total ? t_hist.ms() : 0;
    if (N) {
        info.build_ms[1] = t_enc.ms();
        info.build_ms[2] = t_sort.ms();
        info.build_ms[3] = t_grp.ms();
        info.build_ms[4] = t_fwd.ms();
    }
    info.build_ms[5] = t_all.ms();
    info.build_ms[6] = t_h2d.ms();
    info.build_ms[7] = (double)launches;

    if (opt.verbose) {
        // the reference's cost report (library.cpp:347-370); "Total cost" is the work-unit ground truth
        printf("------------\nCOMPUTATIONAL COSTS: \nTotal cost: %llu lookups\nLinear ratio: %g\n", (unsigned long long)lookups,
               N ? (double)((float)lookups / (float)N) : 0.0);
        printf("Index: %u genes, %u genomes, k=%d, base=%u, %llu k-mers, %llu postings, %llu shared (gene,k-mer) pairs\n", S, G, k, base,
               (unsigned long long)N, (unsigned long long)info.U, (unsigned long long)info.R);
        printf("------------\n\n");
        fflush(stdout);
    }
}

// ------------------------------------------------------------------------------------------------ sharded build, later steps

// Step 1b: room for the keys of this rank's slice, which the caller's all-to-all delivers in source-rank order (= gene order).
uint64_t* Index::shard_recv(uint64_t n_recv) {
    if (!shard) throw Error(PD_ERR_INVALID, "not a sharded build in progress");
    rt::set_device(device);
    if (n_recv >= (1ull << 31)) throw Error(PD_ERR_UNSUPPORTED, "2^31 or more k-mers in one slice");
    shard->n_recv = n_recv;
    shard->keys_b.alloc(std::max<uint64_t>(n_recv, 1));
    return shard->keys_b.p;
}

// Step 2: sort the slice by rank (stable: genes stay ascending inside a rank), count-dedup, rank groups, this slice's part
// of the per-gene counts.
void Index::shard_sort(pd_shard_info* out) {
    if (!shard || !shard->keys_b.p) throw Error(PD_ERR_INVALID, "pd_shard_recv has not been called");
    rt::set_device(device);
    Shard& sh = *shard;
    const uint32_t S = info.S;
    const int seq_bits = (int)info.seq_bits;
    const uint64_t Nr = sh.n_recv;
    if (Nr < 2) {
        // A (nearly) empty slice — an input too small for this many ranks.  Throwing here would leave the other ranks waiting in
        // their next collective: the slice is reported as unusable (multi = all ones) and every rank refuses together after
        // the counts have been exchanged (multigpu.build_sharded; pd_shard_buffers refuses too).
        sh.U_r = 0;
        sh.M_r = 0;
        sh.unusable = true;
        sh.gene_counts.alloc((size_t)2 * std::max<uint32_t>(S, 1));
        rt::stream_t st0 = rt::stream_create();
        rt::zero(sh.gene_counts.p, sh.gene_counts.bytes(), st0);
        rt::sync(st0);
        rt::stream_destroy(st0);
        out->entries = 0;
        out->multi = ~0ull;
        out->kmers = Nr;
        out->d_gene_counts = reinterpret_cast<uint64_t*>(sh.gene_counts.p);
        return;
    }
    rt::stream_t st = rt::stream_create();
    uint64_t launches = 0;
    Timer t_all(st), t_sort(st), t_grp(st);
    t_all.start();
    t_sort.start();
    sh.keys_a.alloc(Nr);   // the send buffer has done its job: the second buffer of the ping-pong
    const int passes = sortk::passes_of(sh.rank_bits);
    const uint32_t rtiles = sortk::tiles_of(Nr);
    rt::DevBuf<uint32_t> sort_ctl((size_t)2 * passes * sortk::kRadix + 32), sort_status((size_t)passes * rtiles * sortk::kRadix);
    uint32_t* d_hist = sort_ctl.p;
    uint32_t* d_bins = d_hist + passes * sortk::kRadix;
    uint32_t* d_tot = d_bins + passes * sortk::kRadix;
    rt::zero(sort_ctl.p, sort_ctl.bytes(), st);
    rt::zero(sort_status.p, sort_status.bytes(), st);
    PD_LAUNCH(sortk::keys_hist_kernel, std::min<uint32_t>(rtiles, (uint32_t)sms * 8), sortk::kThreads, 0, st, (const uint64_t*)sh.keys_b.p, Nr, seq_bits, passes,
              d_hist);
    PD_LAUNCH(sortk::digit_bins_kernel, 1, sortk::kRadix, 0, st, (const uint32_t*)d_hist, passes, d_bins, d_tot);
    launches += 2;
    uint64_t* src = sh.keys_b.p;
    uint64_t* dst = sh.keys_a.p;
    {
        sortk::EncodeSrc es;
        memset(&es, 0, sizeof(es));
        const size_t smem = sortk::sweep_smem_bytes();
        void (*kfn)(sortk::SweepArgs, sortk::EncodeSrc) = sortk::onesweep_kernel<0, uint32_t>;
        rt::allow_smem(kfn, smem);
        for (int p = 0; p < passes; p++) {
            sortk::SweepArgs sa;
            sa.keys = src; sa.out = dst; sa.n = Nr; sa.shift = seq_bits + 8 * p;
            sa.bins = d_bins + p * sortk::kRadix;
            sa.status = sort_status.p + (size_t)p * rtiles * sortk::kRadix;
            sa.counter = d_tot + 8 + p;
            sa.tile0 = 0;
            PD_LAUNCH(kfn, rtiles, sortk::kThreads, smem, st, sa, es);
            launches++;
            std::swap(src, dst);
        }
    }
    const uint64_t* sorted = src;
    t_sort.stop();

    // count dedup + rank groups of the slice (group boundaries are rank boundaries: no group straddles two slices)
    t_grp.start();
    const uint32_t tiles = (uint32_t)((Nr + ik::kEntTile - 1) / ik::kEntTile);
    rt::DevBuf<uint32_t> tile_h((size_t)tiles + 1), tile_g((size_t)tiles + 1), d_spur(4), d_total(4), scratch(prims::scan_tmp_words(tiles) + 16);
    rt::zero(d_spur.p, 4 * sizeof(uint32_t), st);
    PD_LAUNCH(ik::entry_count_kernel, tiles, ik::kEntThreads, 0, st, sorted, Nr, seq_bits, tile_h.p, tile_g.p);
    prims::exclusive_scan_u32(tile_h.p, tile_h.p, tiles, scratch.p, d_total.p, st, &launches);
    prims::exclusive_scan_u32(tile_g.p, tile_g.p, tiles, scratch.p, d_total.p + 1, st, &launches);
    uint32_t h_tot[2] = {0, 0};
    rt::d2h(h_tot, d_total.p, sizeof(h_tot), st);
    rt::sync(st);
    const uint32_t Ur = h_tot[0], g_counted = h_tot[1];
    sh.U_r = Ur;
    sh.post_slice.alloc(Ur);
    sh.heads_slice.alloc((size_t)Ur / 32 + 2);
    const uint32_t multi_cap = (uint32_t)std::min<uint64_t>((uint64_t)Ur, (uint64_t)(Nr - Ur) + 16);  // every repeated entry merged at least one key
    sh.multi_slice.alloc((size_t)2 * multi_cap + 2);
    rt::DevBuf<uint32_t> cnt_tmp(Ur), gid_tmp(Ur), head_tmp((size_t)g_counted + 1);
    rt::zero(sh.heads_slice.p, sh.heads_slice.bytes(), st);
    ik::ShardOut so;
    memset(&so, 0, sizeof(so));
    so.head_bits = sh.heads_slice.p;
    so.multi = sh.multi_slice.p;
    so.n_multi = d_spur.p + 1;
    so.multi_cap = multi_cap;
    so.tail_merge = shard_rank + 1 == shard_world ? 1u : 0u;
    PD_LAUNCH(ik::entry_apply_kernel, tiles, ik::kEntThreads, 0, st, sorted, Nr, seq_bits, (const uint32_t*)tile_h.p, (const uint32_t*)tile_g.p, Ur,
              sh.post_slice.p, cnt_tmp.p, gid_tmp.p, head_tmp.p, (uint64_t*)nullptr, d_spur.p, so);
    PD_LAUNCH(ik::group_tail_kernel, 1, 32, 0, st, head_tmp.p, g_counted, (const uint32_t*)d_spur.p, Ur);
    // this slice's part of the per-gene list-class counts and of total_visited (library.cpp:327)
    sh.gene_counts.alloc((size_t)2 * S);
    rt::zero(sh.gene_counts.p, sh.gene_counts.bytes(), st);
    PD_LAUNCH(ik::fwd_count_kernel, blocks_for(Ur, 256 * ik::kFwdItems), 256, 0, st, (const uint32_t*)sh.post_slice.p, (const uint32_t*)gid_tmp.p,
              (const uint32_t*)head_tmp.p, Ur, sk::kShortList, sk::kHugeList, sh.gene_counts.p, sh.gene_counts.p + S);
    uint32_t h_sp[2] = {0, 0};
    rt::d2h(h_sp, d_spur.p, sizeof(h_sp), st);
    launches += 4;
    t_grp.stop();
    t_all.stop();
    rt::sync(st);
    if (h_sp[1] > multi_cap) throw Error(PD_ERR_CUDA, "sharded build: repeated-entry list overflow");
    sh.M_r = h_sp[1];
    sh.keys_a.release();
    sh.keys_b.release();
    info.U = Ur;   // until pd_shard_finish: this slice only
    info.build_ms[2] = t_sort.ms();
    info.build_ms[3] = t_grp.ms();
    info.build_ms[5] += t_all.ms();
    info.build_ms[7] += (double)launches;
    rt::stream_destroy(st);
    out->entries = sh.U_r;
    out->multi = sh.M_r;
    out->kmers = Nr;
    out->d_gene_counts = reinterpret_cast<uint64_t*>(sh.gene_counts.p);
}


void Index::shard_buffers(uint64_t max_entries, uint64_t max_multi, pd_shard_arrays* out) {
    if (!shard) throw Error(PD_ERR_INVALID, "not a sharded build in progress");
    rt::set_device(device);
    Shard& sh = *shard;
    if (sh.unusable) throw Error(PD_ERR_UNSUPPORTED, "sharded build: a rank's slice of the k-mer ranks is (nearly) empty — use fewer ranks for this input");
    if (max_entries < sh.U_r || (max_multi && max_multi < sh.M_r)) throw Error(PD_ERR_INVALID, "segment smaller than this rank's slice");
    sh.seg = (max_entries + 4095) / 4096 * 4096;
    sh.mseg = max_multi;   // 0: the repeated-entry lists get their array later (shard_multi), once their sizes are known
    if (sh.seg * shard_world >= 0x7FFFFFFFull) throw Error(PD_ERR_UNSUPPORTED, "2^31 or more (padded) entries");
    rt::stream_t st = rt::stream_create();
    post.alloc((size_t)sh.seg * shard_world);
    sh.heads_all.alloc((size_t)sh.seg / 32 * shard_world);
    if (sh.mseg) sh.multi_all.alloc((size_t)2 * sh.mseg * shard_world);
    uint32_t* my_post = post.p + (size_t)sh.seg * shard_rank;
    uint32_t* my_heads = sh.heads_all.p + (size_t)sh.seg / 32 * shard_rank;
    rt::zero(my_heads, sizeof(uint32_t) * (sh.seg / 32), st);
    rt::d2d(my_post, sh.post_slice.p, sizeof(uint32_t) * sh.U_r, st);
    rt::d2d(my_heads, sh.heads_slice.p, sizeof(uint32_t) * (((size_t)sh.U_r + 31) / 32), st);
    if (sh.seg > sh.U_r)
        PD_LAUNCH(ik::shard_pad_kernel, blocks_for(sh.seg - sh.U_r), 256, 0, st, my_post, my_heads, sh.U_r, (uint32_t)sh.seg);
    if (sh.mseg) rt::d2d(sh.multi_all.p + (size_t)2 * sh.mseg * shard_rank, sh.multi_slice.p, sizeof(uint32_t) * 2 * sh.M_r, st);
    rt::sync(st);
    rt::stream_destroy(st);
    sh.post_slice.release();
    sh.heads_slice.release();
    if (sh.mseg) sh.multi_slice.release();
    out->d_post = post.p;
    out->d_heads = sh.heads_all.p;
    out->d_multi = sh.mseg ? sh.multi_all.p : nullptr;
    out->seg = sh.seg;
    out->mseg = sh.mseg;
}

// The array of the repeated-entry lists when pd_shard_buffers was called before their sizes were known (max_multi = 0).
uint32_t* Index::shard_multi(uint64_t max_multi) {
    if (!shard || !shard->seg) throw Error(PD_ERR_INVALID, "pd_shard_buffers has not been called");
    rt::set_device(device);
    Shard& sh = *shard;
    if (max_multi < sh.M_r) throw Error(PD_ERR_INVALID, "segment smaller than this rank's list");
    sh.mseg = std::max<uint64_t>(max_multi, 1);
    sh.multi_all.alloc((size_t)2 * sh.mseg * shard_world);
    rt::stream_t st = rt::stream_create();
    rt::d2d(sh.multi_all.p + (size_t)2 * sh.mseg * shard_rank, sh.multi_slice.p, sizeof(uint32_t) * 2 * sh.M_r, st);
    rt::sync(st);
    rt::stream_destroy(st);
    sh.multi_slice.release();
    return sh.multi_all.p;
}

// Step 3a: everything that does not need the gathered postings — multiplicities, group structure from the head bits,
// per-gene counts, the query partition.  The caller may run it while the all-gather of the postings is still in flight.
void Index::shard_groups(const uint64_t* entries_of_rank, const uint64_t* multi_of_rank, uint32_t* bounds) {
    if (!shard || !shard->seg) throw Error(PD_ERR_INVALID, "pd_shard_buffers has not been called");
    if (shard->grouped) throw Error(PD_ERR_INVALID, "pd_shard_groups called twice");
    if (!shard->mseg) throw Error(PD_ERR_INVALID, "pd_shard_multi has not been called");
    rt::set_device(device);
    Shard& sh = *shard;
    const uint32_t S = info.S, G = info.G, W = shard_world;
    const uint64_t E = sh.seg * W;   // entries incl. padding
    rt::stream_t st = rt::stream_create();
    Timer t_fin(st);
    t_fin.start();
    uint64_t launches = 0;
    uint64_t U_all = 0;
    std::vector<uint32_t> h_nm(W);
    for (uint32_t r = 0; r < W; r++) {
        if (entries_of_rank[r] > sh.seg || multi_of_rank[r] > sh.mseg) throw Error(PD_ERR_INVALID, "a rank's slice exceeds the segment size");
        U_all += entries_of_rank[r];
        h_nm[r] = (uint32_t)multi_of_rank[r];
    }
    // ---- multiplicities of the repeated entries of all slices
    post_cnt.alloc((size_t)E);
    {
        rt::DevBuf<uint32_t> d_nm(W);
        rt::h2d(d_nm.p, h_nm.data(), sizeof(uint32_t) * W, st);
        PD_LAUNCH(ik::shard_multi_kernel, blocks_for(sh.mseg * W), 256, 0, st, (const uint32_t*)sh.multi_all.p, (const uint32_t*)d_nm.p, (uint32_t)sh.mseg,
                  (uint32_t)sh.seg, W, post_cnt.p);
        launches++;
        rt::sync(st);  // d_nm, h_nm
    }
    // ---- every entry's group, every group's first entry, from the head bits
    const uint64_t words = E / 32;
    const uint32_t htiles = (uint32_t)((words + ik::kHeadTileWords - 1) / ik::kHeadTileWords);
    sh.tile_heads.alloc((size_t)htiles + 1);
    rt::DevBuf<uint32_t>& tile_heads = sh.tile_heads;
    rt::DevBuf<uint32_t> d_tot(4), scratch(prims::scan_tmp_words(std::max<uint64_t>(htiles, (uint64_t)S + 1)) + 16);
    PD_LAUNCH(ik::head_count_kernel, htiles, ik::kHeadTileWords, 0, st, (const uint32_t*)sh.heads_all.p, words, tile_heads.p);
    prims::exclusive_scan_u32(tile_heads.p, tile_heads.p, htiles, scratch.p, d_tot.p, st, &launches);
    uint32_t h_heads = 0;
    rt::d2h(&h_heads, d_tot.p, sizeof(uint32_t), st);
    rt::sync(st);
    grp_head.alloc((size_t)h_heads + 1);
    rt::zero(d_tot.p, 4 * sizeof(uint32_t), st);
    PD_LAUNCH(ik::head_apply_kernel, htiles, ik::kHeadTileWords, 0, st, (const uint32_t*)sh.heads_all.p, words, (const uint32_t*)tile_heads.p, grp_head.p);
    PD_LAUNCH(ik::group_tail_kernel, 1, 32, 0, st, grp_head.p, h_heads, (const uint32_t*)d_tot.p, (uint32_t)E);
    launches += 3;

    // ---- per-gene class counts and total_visited: all-reduced by the caller; the query partition from per-genome sums
    rt::d2d(cls.p, sh.gene_counts.p, sizeof(unsigned long long) * S, st);
    rt::d2d(d_visited.p, sh.gene_counts.p + S, sizeof(unsigned long long) * S, st);
    rt::DevBuf<unsigned long long> d_cost((size_t)G + 2);
    rt::DevBuf<uint32_t> d_first((size_t)2 * G + 2);   // first gene, gene count per genome
    rt::zero(d_cost.p, d_cost.bytes(), st);
    rt::fill_byte(d_first.p, 0xFF, sizeof(uint32_t) * G, st);
    rt::zero(d_first.p + G, sizeof(uint32_t) * ((size_t)G + 2), st);
    PD_LAUNCH(ik::genome_cost_kernel, blocks_for(S), 256, 0, st, (const unsigned long long*)d_visited.p, (const uint2*)meta.p, S, d_cost.p, d_cost.p + G,
              d_first.p, d_first.p + G);
    launches++;
    std::vector<unsigned long long> h_cost((size_t)G + 2);
    std::vector<uint32_t> h_first((size_t)2 * G + 2);
    rt::d2h(h_cost.data(), d_cost.p, sizeof(unsigned long long) * ((size_t)G + 1), st);
    rt::d2h(h_first.data(), d_first.p, sizeof(uint32_t) * 2 * G, st);
    rt::sync(st);
    std::vector<uint32_t> gstart((size_t)G + 1, 0);   // genes of a genome contiguous and genomes ascending  <=>  first = prefix sum of count
    {
        bool ok = true;
        uint32_t at = 0;
        for (uint32_t g = 0; g < G && ok; g++) {
            const uint32_t n = h_first[G + g];
            ok = n == 0 || h_first[g] == at;
            gstart[g] = at;
            at += n;
        }
        gstart[G] = at;
        if (!ok || at != S) throw Error(PD_ERR_UNSUPPORTED, "sharded build: genes of a genome must be contiguous and genomes ascending");
    }
    unsigned long long all_cost = 0;
    for (uint32_t g = 0; g < G; g++) all_cost += h_cost[g];
    std::vector<uint32_t> gcut(W + 1, 0);
    gcut[W] = G;
    {
        unsigned long long cum = 0;
        uint32_t g = 0;
        for (uint32_t p = 1; p < W; p++) {
            const unsigned long long target = all_cost / W * p + (all_cost % W) * p / W;
            while (g < G && cum + h_cost[g] / 2 < target) cum += h_cost[g++];   // nearest genome boundary
            gcut[p] = std::max(g, gcut[p - 1]);
        }
    }
    for (uint32_t p = 0; p <= W; p++) bounds[p] = gstart[gcut[p]];
    own_g0 = gcut[shard_rank];
    own_g1 = gcut[shard_rank + 1];
    own_row0 = bounds[shard_rank];
    own_row1 = bounds[shard_rank + 1];

    uint64_t pad = 0;
    for (uint32_t r = 0; r < W; r++) pad += sh.seg - entries_of_rank[r];
    t_fin.stop();
    rt::sync(st);
    info.U = U_all;
    info.groups = (uint64_t)h_heads - pad;
    info.lookups = h_cost[G];
    sh.fin_ms = t_fin.ms();
    sh.fin_launches = launches;
    sh.grouped = true;
    rt::stream_destroy(st);
}

// Step 3b: forward lists of this rank's rows from the gathered postings.
void Index::shard_finish() {
    if (!shard || !shard->grouped) throw Error(PD_ERR_INVALID, "pd_shard_groups has not been called");
    rt::set_device(device);
    Shard& sh = *shard;
    const uint32_t S = info.S, G = info.G, W = shard_world;
    const uint64_t E = sh.seg * W;
    rt::stream_t st = rt::stream_create();
    Timer t_fin(st);
    t_fin.start();
    uint64_t launches = sh.fin_launches;
    rt::DevBuf<uint32_t>& tile_heads = sh.tile_heads;
    rt::DevBuf<uint32_t> scratch(prims::scan_tmp_words((uint64_t)S + 1) + 16);
    // ---- forward lists of this rank's rows only (library.cpp:308-330)
    rt::DevBuf<uint32_t> gene_tot((size_t)S + 1), cur3(std::max<size_t>((size_t)3 * S, 1));
    PD_LAUNCH(ik::fwd_totals_kernel, blocks_for((uint64_t)S + 1), 256, 0, st, (const unsigned long long*)cls.p, S, own_row0, own_row1, gene_tot.p);
    prims::exclusive_scan_u32(gene_tot.p, fwd_ptr.p, (uint64_t)S + 1, scratch.p, nullptr, st, &launches);
    uint32_t h_R = 0;
    rt::d2h(&h_R, fwd_ptr.p + S, sizeof(uint32_t), st);
    rt::sync(st);
    const uint32_t R = h_R;
    fwd.alloc(std::max<size_t>(R, 1));
    fwd_cnt.alloc(std::max<size_t>(R, 1));
    if (R) {
        const uint32_t bshift = (uint32_t)std::max(0, (int)info.seq_bits - 8);
        rt::DevBuf<uint4> records(R);
        rt::DevBuf<uint32_t> bucket_cur(ik::kMaxBuckets);
        rt::zero(bucket_cur.p, sizeof(uint32_t) * ik::kMaxBuckets, st);
        PD_LAUNCH(ik::fwd_partition_kernel, blocks_for(E, ik::kPartTile), ik::kPartThreads, 0, st, (const uint32_t*)post.p, (const uint32_t*)post_cnt.p,
                  (const uint32_t*)nullptr, (const uint32_t*)grp_head.p, (uint32_t)E, S, sk::kShortList, sk::kHugeList, bshift, (const uint32_t*)fwd_ptr.p,
                  bucket_cur.p, records.p, own_row0, own_row1, (const uint32_t*)sh.heads_all.p, (const uint32_t*)tile_heads.p);
        PD_LAUNCH(ik::fwd_cursor_init_kernel, blocks_for(S), 256, 0, st, (const unsigned long long*)cls.p, (const uint32_t*)fwd_ptr.p, S, cur3.p);
        PD_LAUNCH(ik::fwd_place_kernel, blocks_for(R, 256 * ik::kFwdItems), 256, 0, st, (const uint4*)records.p, R, cur3.p, fwd.p, fwd_cnt.p);
        launches += 3;
        rt::sync(st);  // records, bucket_cur
    }
    if (own_row1 > own_row0)
        PD_LAUNCH(ik::gene_visited_kernel, blocks_for((uint64_t)(own_row1 - own_row0) * 32), 256, 0, st, (const uint2*)fwd.p, (const uint32_t*)fwd_ptr.p,
                  own_row0, own_row1, sk::kShortList, (unsigned long long*)nullptr, fam_key.p, (unsigned long long*)nullptr);
    launches += 3;
    t_fin.stop();
    rt::sync(st);
    info.R = R;
    info.build_ms[4] = t_fin.ms() + sh.fin_ms;
    info.build_ms[5] += t_fin.ms() + sh.fin_ms;
    info.build_ms[7] += (double)launches;
    grp_head.release();
    rt::stream_destroy(st);
    delete shard;
    shard = nullptr;
    if (opt.verbose) {
        printf("------------\nCOMPUTATIONAL COSTS: \nTotal cost: %llu lookups\nLinear ratio: %g\n", (unsigned long long)info.lookups,
               info.N ? (double)((float)info.lookups / (float)info.N) : 0.0);
        printf("Index: %u genes, %u genomes, k=%d, %llu k-mers, %llu postings; rank %u of %u holds forward lists of genes [%u, %u): %llu entries\n", S, G,
               info.k, (unsigned long long)info.N, (unsigned long long)info.U, shard_rank, W, own_row0, own_row1, (unsigned long long)info.R);
        printf("------------\n\n");
        fflush(stdout);
    }
}

// Host copies of the per-gene tables, on first use (pd_gene_stats, pd_partition_rows, pd_compute_scores).
void Index::host_mirrors() {
    std::lock_guard<std::mutex> lk(mirror_mu);
    if (have_mirrors) return;
    rt::set_device(device);
    const uint32_t S = info.S;
    kseq.assign(S, 0);
    genome_of.assign(S, 0);
    visited.assign(S, 0);
    if (S) {
        std::vector<uint2> m(S);
        rt::stream_t st = rt::stream_create();
        rt::d2h(m.data(), meta.p, sizeof(uint2) * S, st);
        static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "u64");
        rt::d2h(visited.data(), d_visited.p, sizeof(uint64_t) * S, st);
        rt::sync(st);
        rt::stream_destroy(st);
        for (uint32_t s = 0; s < S; s++) {
            kseq[s] = m[s].x;
            genome_of[s] = m[s].y;
        }
    }
    have_mirrors = true;
}

// genes of each genome in input order (genome_sequences, library.cpp:245), host and device.  Made on the device — a
// stable sort of (genome << 32 | gene) on the genome bits — because the host version (three passes over S genes with
// scattered writes, after copying the gene table back) cost tens of milliseconds in front of the first computeScores.
namespace {
__global__ void __launch_bounds__(256) genome_key_kernel(const uint2* __restrict__ meta, uint32_t S, uint64_t* __restrict__ keys,
                                                          uint32_t* __restrict__ counts) {
    const uint32_t s = blockIdx.x * 256u + threadIdx.x;
    const bool valid = s < S;
    const uint32_t g = valid ? meta[s].y : 0xFFFFFFFFu;
    if (valid) keys[s] = ((uint64_t)g << 32) | s;
    const unsigned peers = __match_any_sync(0xffffffffu, g);  // genes of a genome sit together: one add per run
    if (valid && (threadIdx.x & 31u) == (unsigned)(__ffs((int)peers) - 1)) atomicAdd(&counts[g], (uint32_t)__popc(peers));
}
__global__ void __launch_bounds__(256) genome_rows_kernel(const uint64_t* __restrict__ sorted, uint32_t S, const uint32_t* __restrict__ gptr,
                                                           uint32_t* __restrict__ rows, uint32_t* __restrict__ local_of) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= S) return;
    const uint64_t key = sorted[i];
    const uint32_t s = (uint32_t)key, g = (uint32_t)(key >> 32);
    rows[i] = s;
    local_of[s] = i - gptr[g];
}
}  // namespace

void Index::genome_lists() {
    std::lock_guard<std::mutex> lk(mirror_mu);
    if (have_genome_lists) return;
    rt::set_device(device);
    const uint32_t S = info.S, G = info.G;
    genome_ptr.assign((size_t)G + 1, 0);
    d_genome_rows.alloc(std::max<size_t>(S, 1));
    d_local_of.alloc(std::max<size_t>(S, 1));
    h_genome_rows.ensure(std::max<size_t>(S, 1));
    genome_rows = h_genome_rows.p;
    if (S) {
        rt::stream_t st = rt::stream_create();
        rt::DevBuf<uint64_t> keys((size_t)2 * S);
        rt::DevBuf<uint32_t> gptr((size_t)G + 1), tmp(prims::radix_tmp_words(S) + prims::scan_tmp_words((uint64_t)G + 1) + 32);
        rt::zero(gptr.p, sizeof(uint32_t) * ((size_t)G + 1), st);
        PD_LAUNCH(genome_key_kernel, (S + 255) / 256, 256, 0, st, (const uint2*)meta.p, S, keys.p, gptr.p);
        prims::exclusive_scan_u32(gptr.p, gptr.p, (uint64_t)G + 1, tmp.p, nullptr, st);
        int gbits = 1;
        while (gbits < 32 && (1ull << gbits) < (uint64_t)G) gbits++;
        const uint64_t* sorted = prims::radix_sort_u64(keys.p, keys.p + S, S, 32, 32 + gbits, tmp.p + prims::scan_tmp_words((uint64_t)G + 1) + 16, st);
        PD_LAUNCH(genome_rows_kernel, (S + 255) / 256, 256, 0, st, sorted, S, (const uint32_t*)gptr.p, d_genome_rows.p, d_local_of.p);
        rt::d2h(h_genome_rows.p, d_genome_rows.p, sizeof(uint32_t) * S, st);
        rt::d2h(genome_ptr.data(), gptr.p, sizeof(uint32_t) * ((size_t)G + 1), st);
        rt::sync(st);
        rt::stream_destroy(st);
    }
    have_genome_lists = true;
}

// ------------------------------------------------------------------------------------------------ scoring

namespace {

static const int kMaxCtasPerSm = 16;
static const int kLevels = 4;  // three first-try levels by row size + the retry level

__global__ void __launch_bounds__(256) fill_u32_kernel(uint32_t* p, uint32_t n, uint32_t v) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i < n) p[i] = v;
}

// Zero fill of per-call buffers by a kernel.  cudaMemsetAsync of a large range may be carried out by a copy engine, where
// it queues behind the result arrays another context is sending to the host; a kernel only needs SMs.
__global__ void __launch_bounds__(256) zero_words_kernel(uint32_t* p, uint64_t words) {
    const uint64_t stride = (uint64_t)gridDim.x * 256u;
    uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    const uint64_t quads = (reinterpret_cast<uintptr_t>(p) & 15u) == 0 ? words / 4 : 0;
    uint4* q = reinterpret_cast<uint4*>(p);
    for (uint64_t j = i; j < quads; j += stride) q[j] = make_uint4(0, 0, 0, 0);
    for (uint64_t j = quads * 4 + i; j < words; j += stride) p[j] = 0;
}
static void zero_words(void* p, uint64_t bytes, rt::stream_t st) {
    if (!bytes) return;
    const uint64_t words = bytes / 4;  // every buffer here holds 4- or 8-byte elements
    const unsigned grid = (unsigned)std::min<uint64_t>((words / 4 + 255) / 256 + 1, 148 * 8);
    PD_LAUNCH(zero_words_kernel, grid, 256, 0, st, static_cast<uint32_t*>(p), words);
}

// Job counters and completion flags go to the host by stores into pinned (device-visible) host memory, and the host
// waits by polling that memory (ScoreContext::wait).  Two things measured on the per-genome path with several calling
// threads made this necessary: a cudaMemcpyAsync of the counters queues on the device-to-host copy engine behind the
// result arrays another call is sending, and a thread blocked in cudaStreamSynchronize slows down the kernel launches
// of the other threads by an order of magnitude (driver lock).  Polling memory involves no driver call.
static const uint32_t kSeqSlot = 15;
__global__ void publish_counters_kernel(unsigned long long* host, const unsigned long long* dev, uint32_t first, uint32_t count,
                                        unsigned long long seq) {
    if (threadIdx.x < count) host[first + threadIdx.x] = dev[first + threadIdx.x];
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile unsigned long long*>(host + kSeqSlot) = seq;
        __threadfence_system();
    }
}

static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield");
#endif
}
// queues the counters [first, first + count) and a completion number behind the work already on the context's stream
static void publish(ScoreContext& c, uint32_t first, uint32_t count) {
    c.seq++;
    PD_LAUNCH(publish_counters_kernel, 1, 32, 0, c.st, c.h_counters.p, c.d_counters.p, first, count, c.seq);
}
// waits for the last publish() by polling the completion number.  The stream itself is only asked (for failures) every
// 20 ms: cudaStreamQuery takes the driver lock that the launches of the other calling threads need.
static void wait(ScoreContext& c) {
    volatile unsigned long long* f = c.h_counters.p + kSeqSlot;
    auto last = std::chrono::steady_clock::now();
    for (uint32_t spins = 1; *f != c.seq; spins++) {
        for (int i = 0; i < 16; i++) cpu_relax();
        if ((spins & 255u) == 0) {
            std::this_thread::yield();  // a box with fewer cores than waiting threads must still make progress
            const auto now = std::chrono::steady_clock::now();
            if (now - last > std::chrono::milliseconds(20)) {
                last = now;
                if (rt::stream_idle(c.st) && *f != c.seq) throw Error(PD_ERR_CUDA, "completion number never arrived");
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
}

// side tables (score_kernels.cuh): keys empty, sums zero
__global__ void __launch_bounds__(256) xtab_init_kernel(uint32_t* xtab, uint32_t ctas) {
    const uint64_t n = (uint64_t)ctas * 5 * sk::kXSlots;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256)
        xtab[i] = (i % (5 * sk::kXSlots)) < sk::kXSlots ? sk::kEmpty : 0u;
}

// One table configuration of score_rows_kernel.  A row goes to the first level whose `max_cols` covers the row's
// bound on distinct columns, min(total_visited, S); rows that overflow their table are re-run on the last level and
// then, if need be, on the dense global path.
struct Level {
    uint32_t t1;        // tier-1 (direct-mapped) slots, a multiple of 32
    uint32_t hbits;     // log2 tier-2 (probing) slots
    int threads;        // CTA size: 128, 256 or 512
    uint32_t fcap;      // forward entries staged per segment
    uint64_t max_cols;  // first-try rows: bound on distinct columns
};

inline uint32_t log2_floor(uint64_t v) {
    uint32_t b = 0;
    while (v > 1) {
        v >>= 1;
        b++;
    }
    return b;
}

// Defaults: three first-try table sizes by row size and a large retry table.  pd_options.hash_log2 replaces the
// first-try ladder by one small table of that many tier-2 slots; PD_SMEM_TOP (tests) caps the bytes of the retry table.
inline bool small_index(const Index& ix) { return ix.info.S <= (1u << 18); }

inline void levels_of(const Index& ix, Level lv[kLevels]) {
    lv[0] = {512, 9, 128, 256, 256};
    lv[1] = {2048, 11, 256, 512, 1536};
    // 4992 tier-1 slots: what two CTAs per SM leave room for (112.9 KB each).  Against 4096 it halves the rows that
    // overflow and are scored a second time (21 K -> 10 K of 3.8 M on the 1,000-genome input) and takes the whole job's
    // scoring from 749 to 729 ms; the landscape around it is bumpy (tools/levels_sweep.py, profiles/r02_levels_sweep.txt).
    lv[2] = {4992, 11, 512, 1024, ~0ull};
    lv[3] = {2048, 14, 1024, 1024, 0};  // few, heavy rows, one CTA per SM: twice the warps per row shortens the tail
    // Small indices (tens of genomes, k = 4..5): a row shares random k-mers with thousands of the few ten thousand
    // genes, so its columns are many relative to S and the tier-1 slots (S / T1 genes each) collide.  A larger table
    // for the top level cuts the rows that have to be scored twice from 1/3 to 1/20 (xanthomonas14: 63 -> 44 ms,
    // ecoli10: 15.5 -> 10 ms); at S in the millions the columns of a row are a vanishing part of S and the two smaller
    // CTAs per SM are faster.
    if (small_index(ix)) lv[2] = {8192, 12, 1024, 1024, ~0ull};
    if (const char* e = getenv("PD_LEVELS")) {  // tuning: "t1:hbits:threads:fcap:maxcols,..." for the four levels
        unsigned t1, h, t, f;
        unsigned long long m;
        int i = 0;
        while (i < kLevels && sscanf(e, "%u:%u:%u:%u:%llu", &t1, &h, &t, &f, &m) == 5) {
            lv[i++] = {t1, h, (int)t, f, m};
            e = strchr(e, ',');
            if (!e) break;
            e++;
        }
        lv[2].max_cols = ~0ull;
    }
    if (ix.opt.hash_log2 > 0) {
        const uint32_t hb = std::min<uint32_t>(std::max<uint32_t>((uint32_t)ix.opt.hash_log2, 5), 14);
        const uint32_t t1 = std::min<uint32_t>(1u << hb, 4096);
        lv[0] = {t1, hb, 128, 256, 0};   // unused
        lv[1] = {t1, hb, 128, 256, 0};   // unused
        lv[2] = {t1, hb, hb >= 12 ? 256 : 128, 256, ~0ull};
        lv[3] = {2048, std::max<uint32_t>(hb, 14), 1024, 1024, 0};
    }
    if (const char* e = getenv("PD_SMEM_TOP")) {  // tests: shrink the retry table to force the dense path
        const size_t v = (size_t)atoll(e);
        if (v >= 64) {
            const uint32_t hb = std::max<uint32_t>(5, log2_floor(v / 8) - 1);
            if (hb < lv[3].hbits) lv[3] = {32, hb, 128, 256, 0};
        }
    }
    for (int i = 0; i < kLevels; i++) {
        lv[i].t1 = std::max<uint32_t>(32, lv[i].t1 / 32 * 32);
        while (sk::score_smem_bytes(lv[i].t1, lv[i].hbits, lv[i].fcap, lv[i].threads) > ix.smem_optin && lv[i].hbits > 5) lv[i].hbits--;
    }
}

template <int THREADS>
void launch_rows_t(ScoreContext& c, sk::ScoreArgs& a, size_t smem) {
    Index& ix = *c.ix;
    // the shared-memory opt-in and the occupancy of a (kernel, shared memory, device) combination never change: ask once
    static std::mutex cache_mu;
    static std::map<std::pair<int, size_t>, int> cache;
    int occ;
    {
        std::lock_guard<std::mutex> lk(cache_mu);
        auto it = cache.find(std::make_pair(ix.device, smem));
        if (it == cache.end()) {
            rt::allow_smem(sk::score_rows_kernel<THREADS>, smem);
            const int o = std::min(kMaxCtasPerSm, std::max(1, rt::occupancy(sk::score_rows_kernel<THREADS>, THREADS, smem)));
            it = cache.emplace(std::make_pair(ix.device, smem), o).first;
        }
        occ = it->second;
    }
    unsigned grid = (unsigned)std::min<uint64_t>(a.n_rows, (uint64_t)ix.sms * occ);
    PD_LAUNCH(sk::score_rows_kernel<THREADS>, grid, THREADS, smem, c.st, a);
}

void launch_rows(ScoreContext& c, sk::ScoreArgs a, const Level& lv, int cursor_id) {
    if (a.n_rows == 0) return;
    a.hbits = lv.hbits;
    a.hmask = (1u << lv.hbits) - 1;
    a.plimit = std::min<uint32_t>(1u << lv.hbits, sk::kProbeLimit);
    a.nslots = lv.t1 + (1u << lv.hbits);
    a.t1 = lv.t1;
    {   // order-preserving slot functions: umulhi(c, mul) = floor(c * slots / S) < slots for every c < S
        const uint64_t S = std::max<uint64_t>(c.ix->info.S, 1);
        a.hmul = (uint32_t)std::min<uint64_t>(((1ull << lv.hbits) << 32) / S, 0xFFFFFFFFull);
        a.t1mul = (uint32_t)std::min<uint64_t>(((uint64_t)lv.t1 << 32) / S, 0xFFFFFFFFull);
    }
    a.fcap = lv.fcap;
    a.cursor = c.d_cursors.p + cursor_id;
    const size_t smem = sk::score_smem_bytes(lv.t1, lv.hbits, lv.fcap, lv.threads);
    if (lv.threads >= 1024) launch_rows_t<1024>(c, a, smem);
    else if (lv.threads >= 512) launch_rows_t<512>(c, a, smem);
    else if (lv.threads >= 256) launch_rows_t<256>(c, a, smem);
    else launch_rows_t<128>(c, a, smem);
    c.stats.launches++;
}

}  // namespace

// Scores the `n` rows gene(i) = genes[i] (device array) or gene_base + i; row i writes best hits to row i of d_bh.
// Returns the number of non-zero cells; cells beyond c.cap are counted but not stored (caller grows and re-runs).
// by_cost: order the rows of a level by descending posting-list volume instead of by family.  Family order makes
// neighbouring rows share posting lists in L2 and is what a block of many genomes wants; the rows of ONE genome belong
// to different families, so there a per-genome call starts its heaviest rows first and shortens the tail of the launch.
static uint64_t run_rows(ScoreContext& c, uint32_t n, const uint32_t* d_genes, uint32_t gene_base, uint32_t* d_bh, uint32_t* d_colmax,
                         uint64_t* pairs, uint64_t* lookups, uint64_t* fwd_entries, bool by_cost) {
    Index& ix = *c.ix;
    trace::Clock tc;
    Level lv[kLevels];
    levels_of(ix, lv);
    const uint32_t want_ctas = (uint32_t)ix.sms * kMaxCtasPerSm;
    if (c.xtab_ctas < want_ctas) {
        c.d_xtab.alloc((size_t)want_ctas * 5 * sk::kXSlots);
        PD_LAUNCH(xtab_init_kernel, (unsigned)ix.sms * 4, 256, 0, c.st, c.d_xtab.p, want_ctas);
        c.xtab_ctas = want_ctas;
    }
    zero_words(c.d_counters.p, 8 * sizeof(unsigned long long), c.st);
    zero_words(c.d_cursors.p, 16 * sizeof(uint32_t), c.st);
    c.d_rows.grow(n);
    c.d_ovf.grow((size_t)2 * n);

    // ---- row descriptors on the device: sorted by (first-try level, family key)
    sk::ClassifyArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.n = n;
    ca.genes = d_genes;
    ca.gene_base = gene_base;
    ca.S = ix.info.S;
    ca.visited = ix.d_visited.p;
    ca.fam_key = ix.fam_key.p;
    ca.fwd_ptr = ix.fwd_ptr.p;
    ca.cls = ix.cls.p;
    ca.cls_bits = ik::kClsBits;
    ca.meta = ix.meta.p;
    for (int l = 0; l < 3; l++) ca.max_cols[l] = lv[l].max_cols;
    ca.counts = c.d_cursors.p + 8;  // [8..10]
    ca.stats = c.d_counters.p + 4;  // [4], [5]
    {
        static const char* const order = getenv("PD_ROW_ORDER");  // tuning: "family" | "cost"
        ca.by_cost = order ? (strcmp(order, "cost") == 0) : by_cost;
    }
    c.d_rowkeys.grow((size_t)2 * n);
    c.d_sorttmp.grow(prims::radix_tmp_words(n) + 16);
    PD_LAUNCH(sk::row_keys_kernel, blocks_for(n), 256, 0, c.st, ca, c.d_rowkeys.p);
    uint64_t nl = 0;
    const uint64_t* sorted_keys = prims::radix_sort_u64(c.d_rowkeys.p, c.d_rowkeys.p + n, n, 31, 64, c.d_sorttmp.p, c.st, &nl);
    PD_LAUNCH(sk::row_desc_kernel, blocks_for(n), 256, 0, c.st, ca, sorted_keys, c.d_rows.p);
    c.stats.launches += 2 + nl;
    tc.lap(trace::kSortLaunch);

    sk::ScoreArgs a;
    memset(&a, 0, sizeof(a));
    a.post = ix.post.p; a.post_cnt = ix.post_cnt.p; a.fwd = ix.fwd.p; a.fwd_cnt = ix.fwd_cnt.p; a.meta = ix.meta.p;
    a.G = ix.info.G; a.S = ix.info.S; a.k2 = 2u * (uint32_t)ix.info.k;
    a.o_score = c.d_score.p; a.o_perc = c.d_perc.p; a.o_trperc = c.d_trperc.p;
    a.o_row = c.d_row.p; a.o_col = c.d_col.p; a.o_g1 = c.d_g1.p; a.o_g2 = c.d_g2.p;
    a.cell_cap = c.cap;
    a.g1_bhrow = c.g1_bhrow ? 1u : 0u;
    a.n_cells = c.d_counters.p + 0;
    a.n_pairs = c.d_counters.p + 1;
    a.bh = d_bh;
    a.colmax = d_colmax;
    a.xtab = c.d_xtab.p;

    const bool skip_retry = small_index(ix) && ix.opt.hash_log2 == 0 && !getenv("PD_LEVELS");
    rt::event_record(c.ev_k0, c.st);
    for (int level = 0; level < kLevels - 1; level++) {
        sk::ScoreArgs b = a;
        b.rows = c.d_rows.p;
        b.n_rows = n;  // upper bound: sizes the grid
        b.n_rows_dev = c.d_cursors.p + 8;
        b.level = (uint32_t)level;
        // small indices: a row that overflows the (already large) top table has a fifth of all genes as columns and never
        // fits the retry table either (measured: every retried row of ecoli10 / xanthomonas14 fell through) — straight
        // to the dense kernel's list then
        b.overflow_rows = c.d_ovf.p + (skip_retry ? n : 0);
        b.n_overflow = c.d_counters.p + (skip_retry ? 3 : 2);
        launch_rows(c, b, lv[level], level);
    }
    // retry level: rows that overflowed their first table.  Their count stays on the device (no host round trip in
    // the middle of the job: a counter read-back would queue behind the result arrays another call is sending)
    if (!skip_retry) {
        sk::ScoreArgs b = a;
        b.rows = c.d_ovf.p;
        b.n_rows = n;  // upper bound: sizes the grid
        b.n_rows_ovf = c.d_counters.p + 2;
        b.overflow_rows = c.d_ovf.p + n;
        b.n_overflow = c.d_counters.p + 3;
        launch_rows(c, b, lv[kLevels - 1], kLevels - 1);
    }
    publish(c, 0, 8);
    rt::event_record(c.ev_k1, c.st);
    tc.lap(trace::kLaunch);
    wait(c);
    tc.lap(trace::kLevelsWait);
    c.stats.kernel_ms += rt::event_ms(c.ev_k0, c.ev_k1);
    c.stats.retry_rows += c.h_counters.p[2];
    // last resort: one counter per gene — in shared memory where the index is small enough, else in global memory
    if (const uint32_t nr = (uint32_t)c.h_counters.p[3]) {
        const uint32_t S = ix.info.S;
        const size_t smem_bytes = (size_t)((S + 1) / 2) * 4;
        const char* const dense_env = getenv("PD_DENSE");  // tests: "global" keeps the global-memory kernel covered
        const bool in_smem = S <= sk::kDenseSmemGenes && ix.info.max_kseq < sk::kDenseSmemMaxK && smem_bytes + 1024 <= ix.smem_optin &&
                             !(dense_env && strcmp(dense_env, "global") == 0);
        unsigned grid;
        size_t words;
        if (in_smem) {
            grid = std::min<unsigned>(nr, (unsigned)ix.sms);
            words = (size_t)grid * 3 * S;
        } else {
            // one CTA per row at a time, each with 20 S bytes of accumulators: as many CTAs as 2 GB of them allow, at most
            // four per SM (at S = 4 M it is 32 CTAs)
            const uint64_t per_cta = (uint64_t)4 * sk::kDenseWordsPerGene * std::max<uint32_t>(S, 1);
            const unsigned fit = (unsigned)std::max<uint64_t>(32, std::min<uint64_t>((uint64_t)ix.sms * 4, (2ull << 30) / per_cta));
            grid = std::min<unsigned>(nr, fit);
            words = (size_t)grid * sk::kDenseWordsPerGene * S;
        }
        // the per-CTA layout follows S and the kernel: a context taken over from an index of another size (or used by the
        // other kernel) holds stale `touched` lists where this layout wants zeros
        const uint32_t layout = in_smem ? 2u : 1u;
        if (c.d_dense.n < words || c.dense_S != S || c.dense_layout != layout) {
            if (c.d_dense.n < words) c.d_dense.alloc(words);
            zero_words(c.d_dense.p, c.d_dense.n * sizeof(uint32_t), c.st);
            c.dense_S = S;
            c.dense_layout = layout;
        }
        sk::ScoreArgs b2 = a;
        b2.rows = c.d_ovf.p + n;
        b2.n_rows = nr;
        b2.cursor = c.d_cursors.p + kLevels;
        sk::DenseArgs d;
        d.S = S;
        d.acc = c.d_dense.p;
        rt::event_record(c.ev_k0, c.st);
        if (in_smem) {
            rt::allow_smem(sk::score_rows_dense_smem_kernel, smem_bytes);
            PD_LAUNCH(sk::score_rows_dense_smem_kernel, grid, sk::kDenseSmemThreads, smem_bytes, c.st, b2, d);
        } else {
            PD_LAUNCH(sk::score_rows_dense_kernel, grid, sk::kDenseThreads, 0, c.st, b2, d);
        }
        c.stats.launches++;
        c.stats.fallback_rows += nr;
        publish(c, 0, 8);
        rt::event_record(c.ev_k1, c.st);
        wait(c);
        tc.lap(trace::kRetryWait);
        c.stats.kernel_ms += rt::event_ms(c.ev_k0, c.ev_k1);
    }
    *pairs = c.h_counters.p[1];
    *lookups = c.h_counters.p[4];
    *fwd_entries = c.h_counters.p[5];
    return c.h_counters.p[0];
}

void Index::compute_scores(uint32_t genome, pd_scores* out) {
    if (genome >= info.G) throw Error(PD_ERR_INVALID, "unknown genome");
    if (shard) throw Error(PD_ERR_INVALID, "sharded build not finished (pd_shard_finish)");
    if (shard_world > 1 && (genome < own_g0 || genome >= own_g1)) throw Error(PD_ERR_INVALID, "genome belongs to another rank of the sharded index");
    rt::set_device(device);
    genome_lists();
    trace::Clock tc;
    ScoreContext* cp = acquire();
    ScoreContext& c = *cp;
    try {
        memset(&c.stats, 0, sizeof(c.stats));
        const uint32_t S = info.S, G = info.G;
        const uint32_t r0 = genome_ptr[genome], rows = genome_ptr[genome + 1] - r0;
        rt::event_record(c.ev_call0, c.st);
        c.d_bh.grow(std::max<size_t>((size_t)rows * G, 1));
        c.d_colmax.ensure(std::max<size_t>(S, 1));
        if (c.cap == 0) c.ensure_cells(opt.cell_capacity ? opt.cell_capacity : std::max<uint64_t>(1u << 16, (uint64_t)rows * 1536));
        uint64_t cells = 0, pairs = 0, lookups = 0, fwd_entries = 0;
        StageLock compute(compute_token, stage_tokens & 1);
        tc.lap(trace::kAcquire);
        for (int attempt = 0; attempt < 3; attempt++) {
            zero_words(c.d_bh.p, sizeof(uint32_t) * (size_t)rows * G, c.st);
            zero_words(c.d_colmax.p, sizeof(uint32_t) * S, c.st);
            cells = rows ? run_rows(c, rows, d_genome_rows.p + r0, 0, c.d_bh.p, c.d_colmax.p, &pairs, &lookups, &fwd_entries, true) : 0;
            if (cells <= c.cap) break;
            c.ensure_cells(cells + cells / 2);
        }
        if (cells > c.cap) throw Error(PD_ERR_CUDA, "cell count unstable between passes");
        if (cells > 0x7fffffffull) throw Error(PD_ERR_UNSUPPORTED, "more than 2^31 cells in one computeScores call");

        compute.unlock();
        tc = trace::Clock();
        StageLock copy(copy_token, stage_tokens & 2);
        tc.lap(trace::kTokenWait);
        c.ensure_host_cells();
        c.h_bh.grow(std::max<size_t>((size_t)rows * G, 1));
        c.h_colmax.ensure(std::max<size_t>(S, 1));
        if (c.h_map.n < std::max<size_t>(S, 1)) {  // flat_map (library.cpp:428-432): all INT32_MAX between calls
            c.h_map.ensure(std::max<size_t>(S, 1));
            for (uint32_t s = 0; s < S; s++) c.h_map.p[s] = INT32_MAX;
        }
        rt::d2h(c.h_score.p, c.d_score.p, sizeof(float) * cells, c.st);
        rt::d2h(c.h_perc.p, c.d_perc.p, sizeof(float) * cells, c.st);
        rt::d2h(c.h_trperc.p, c.d_trperc.p, sizeof(float) * cells, c.st);
        rt::d2h(c.h_row.p, c.d_row.p, sizeof(int32_t) * cells, c.st);
        rt::d2h(c.h_col.p, c.d_col.p, sizeof(int32_t) * cells, c.st);
        static const bool host_g1 = getenv("PD_HOST_G1") ? atoi(getenv("PD_HOST_G1")) != 0 : true;
        if (!host_g1) rt::d2h(c.h_g1.p, c.d_g1.p, sizeof(int32_t) * cells, c.st);
        rt::d2h(c.h_g2.p, c.d_g2.p, sizeof(int32_t) * cells, c.st);
        rt::d2h(c.h_bh.p, c.d_bh.p, sizeof(float) * (size_t)rows * G, c.st);
        rt::d2h(c.h_colmax.p, c.d_colmax.p, sizeof(float) * S, c.st);
        rt::event_record(c.ev_call1, c.st);
        publish(c, 0, 0);
        // while the copies run: first_seq_genome is the genome of the cell's row (library.cpp:571-572), i.e. the call's
        // genome in every cell: written here instead of crossing PCIe (one seventh of the cell bytes); and this
        // genome's rows go into flat_map (release() puts INT32_MAX back)
        if (host_g1) fill_i32_stream(c.h_g1.p, cells, (int32_t)genome);
        for (uint32_t i = 0; i < rows; i++) c.h_map.p[genome_rows[r0 + i]] = (int32_t)i;
        c.map_r0 = r0;
        c.map_rows = rows;
        tc.lap(trace::kCopyEnqueue);
        wait(c);
        copy.unlock();
        tc.lap(trace::kCopyWait);
        rt::collect();
        tc.lap(trace::kTail);
        if (trace::on()) trace::ns[trace::kCalls]++;
        c.stats.total_ms = rt::event_ms(c.ev_call0, c.ev_call1);
        c.stats.rows = rows;
        c.stats.lookups = lookups;
        c.stats.fwd_entries = fwd_entries;
        c.stats.pairs = pairs;
        c.stats.cells = cells;

        out->scoresCount = (int32_t)cells;
        out->S = (int32_t)S;
        out->rows = (int32_t)rows;
        out->G = (int32_t)G;
        out->scores = c.h_score.p; out->percs = c.h_perc.p; out->tr_percs = c.h_trperc.p;
        out->row = c.h_row.p; out->column = c.h_col.p;
        out->first_seq_genome = c.h_g1.p; out->second_seq_genome = c.h_g2.p;
        out->max_genome_score = c.h_bh.p;
        out->max_genome_score_col = c.h_colmax.p;
        out->scoresMaxMappings = c.h_map.p;
        out->owner = cp;
    } catch (...) {
        release(cp);
        throw;
    }
}

void Index::genome_edges(uint32_t genome, pd_edges* out) {
    if (genome >= info.G) throw Error(PD_ERR_INVALID, "unknown genome");
    if (shard) throw Error(PD_ERR_INVALID, "sharded build not finished (pd_shard_finish)");
    if (shard_world > 1 && (genome < own_g0 || genome >= own_g1)) throw Error(PD_ERR_INVALID, "genome belongs to another rank of the sharded index");
    rt::set_device(device);
    genome_lists();
    ScoreContext* cp = acquire();
    ScoreContext& c = *cp;
    try {
        memset(&c.stats, 0, sizeof(c.stats));
        const uint32_t S = info.S, G = info.G;
        const uint32_t r0 = genome_ptr[genome], rows = genome_ptr[genome + 1] - r0;
        rt::event_record(c.ev_call0, c.st);
        c.d_bh.grow(std::max<size_t>((size_t)rows * G, 1));
        c.d_colmax.ensure(std::max<size_t>(S, 1));
        if (c.cap == 0) c.ensure_cells(opt.cell_capacity ? opt.cell_capacity : std::max<uint64_t>(1u << 16, (uint64_t)rows * 1536));
        uint64_t cells = 0, pairs = 0, lookups = 0, fwd_entries = 0;
        StageLock compute(compute_token, stage_tokens & 1);
        c.g1_bhrow = true;  // the filter wants the row's index inside the genome next to every cell
        for (int attempt = 0; attempt < 3; attempt++) {
            zero_words(c.d_bh.p, sizeof(uint32_t) * (size_t)rows * G, c.st);
            zero_words(c.d_colmax.p, sizeof(uint32_t) * S, c.st);
            cells = rows ? run_rows(c, rows, d_genome_rows.p + r0, 0, c.d_bh.p, c.d_colmax.p, &pairs, &lookups, &fwd_entries, true) : 0;
            if (cells <= c.cap) break;
            c.ensure_cells(cells + cells / 2);
        }
        c.g1_bhrow = false;
        if (cells > c.cap) throw Error(PD_ERR_CUDA, "cell count unstable between passes");

        // ---- the Java host's filter over the cells, on the device (Pangenes.java:98-176)
        uint64_t edges = 0;
        if (cells) {
            c.d_flag.ensure(c.cap);
            c.d_imax.ensure(std::max<size_t>(G, 1));
            c.d_rowthr.grow(std::max<size_t>(rows, 1));
            if (c.edge_cap < c.cap) {  // edges <= cells <= cap: sized with the cell arrays, the retry below never runs
                c.edge_cap = c.cap;
                c.d_esrc.alloc(c.edge_cap); c.d_edst.alloc(c.edge_cap); c.d_escore.alloc(c.edge_cap);
            }
            fk::FilterArgs fa;
            memset(&fa, 0, sizeof(fa));
            fa.cells = cells;
            fa.score = c.d_score.p; fa.row = c.d_row.p; fa.col = c.d_col.p; fa.bhrow = c.d_g1.p; fa.g2 = c.d_g2.p;
            fa.genome = genome; fa.G = G;
            fa.bh = c.d_bh.p; fa.colmax = c.d_colmax.p; fa.local_of = d_local_of.p;
            fa.imax = c.d_imax.p; fa.rowthr = c.d_rowthr.p; fa.flag = c.d_flag.p;
            fa.n_edges = c.d_counters.p + 6;
            zero_words(c.d_imax.p, sizeof(uint32_t) * G, c.st);
            PD_LAUNCH(fill_u32_kernel, blocks_for(rows), 256, 0, c.st, c.d_rowthr.p, rows, 0x7F800000u);  // +inf
            PD_LAUNCH(fk::inter_mark_kernel, blocks_for(cells), 256, 0, c.st, fa);
            PD_LAUNCH(fk::row_threshold_kernel, blocks_for(cells), 256, 0, c.st, fa);
            for (int attempt = 0; attempt < 2; attempt++) {
                fa.e_src = c.d_esrc.p; fa.e_dst = c.d_edst.p; fa.e_score = c.d_escore.p;
                fa.edge_cap = c.edge_cap;
                zero_words(c.d_counters.p + 6, sizeof(unsigned long long), c.st);
                PD_LAUNCH(fk::edge_emit_kernel, blocks_for(cells), 256, 0, c.st, fa);
                publish(c, 6, 1);
                wait(c);
                edges = c.h_counters.p[6];
                if (edges <= c.edge_cap) break;
                c.edge_cap = edges + edges / 8;
                c.d_esrc.alloc(c.edge_cap); c.d_edst.alloc(c.edge_cap); c.d_escore.alloc(c.edge_cap);
            }
            c.stats.launches += 5;
        }
        compute.unlock();
        StageLock copy(copy_token, stage_tokens & 2);
        const size_t ne = std::max<uint64_t>(edges, 1);
        c.h_esrc.grow(ne); c.h_edst.grow(ne); c.h_escore.grow(ne);
        rt::d2h(c.h_esrc.p, c.d_esrc.p, sizeof(uint32_t) * edges, c.st);
        rt::d2h(c.h_edst.p, c.d_edst.p, sizeof(uint32_t) * edges, c.st);
        rt::d2h(c.h_escore.p, c.d_escore.p, sizeof(float) * edges, c.st);
        rt::event_record(c.ev_call1, c.st);
        publish(c, 0, 0);
        wait(c);
        copy.unlock();
        rt::collect();
        c.stats.total_ms = rt::event_ms(c.ev_call0, c.ev_call1);
        c.stats.rows = rows;
        c.stats.lookups = lookups;
        c.stats.fwd_entries = fwd_entries;
        c.stats.pairs = pairs;
        c.stats.cells = cells;
        out->count = edges;
        out->src = c.h_esrc.p;
        out->dst = c.h_edst.p;
        out->score = c.h_escore.p;
        out->cells = cells;
        out->owner = cp;
    } catch (...) {
        c.g1_bhrow = false;
        release(cp);
        throw;
    }
}

void Index::score_partition(uint32_t row_begin, uint32_t row_end, uint32_t rows_per_launch, float* d_best_hit, pd_score_stats* st_out) {
    if (row_begin > row_end || row_end > info.S) throw Error(PD_ERR_INVALID, "bad row range");
    if (shard) throw Error(PD_ERR_INVALID, "sharded build not finished (pd_shard_finish)");
    if (shard_world > 1 && row_begin < row_end && (row_begin < own_row0 || row_end > own_row1))
        throw Error(PD_ERR_INVALID, "rows belong to another rank of the sharded index");
    rt::set_device(device);
    trace::Clock ta;
    ScoreContext* cp = acquire();
    ScoreContext& c = *cp;
    ta.lap(trace::kAcquire);
    try {
        memset(&c.stats, 0, sizeof(c.stats));
        const uint32_t G = info.G;
        const uint32_t total_rows = row_end - row_begin;
        if (rows_per_launch == 0) rows_per_launch = 65536;
        uint32_t* bh = reinterpret_cast<uint32_t*>(d_best_hit);
        if (!bh) {
            c.d_bh.ensure(std::max<size_t>((size_t)total_rows * G, 1));
            bh = c.d_bh.p;
        }
        if (c.cap == 0) c.ensure_cells(opt.cell_capacity ? opt.cell_capacity : std::max<uint64_t>(1u << 20, (uint64_t)rows_per_launch * 1024));
        rt::event_record(c.ev_call0, c.st);
        zero_words(bh, sizeof(uint32_t) * (size_t)total_rows * G, c.st);
        for (uint32_t b0 = row_begin; b0 < row_end; b0 += rows_per_launch) {
            const uint32_t n = std::min(rows_per_launch, row_end - b0);
            uint64_t lookups = 0, pairs = 0, cells = 0, fwd_entries = 0;
            uint32_t* bh_blk = bh + (size_t)(b0 - row_begin) * G;
            for (int attempt = 0; attempt < 3; attempt++) {
                cells = run_rows(c, n, nullptr, b0, bh_blk, nullptr, &pairs, &lookups, &fwd_entries, false);
                if (cells <= c.cap) break;
                c.ensure_cells(cells + cells / 8);
            }
            if (cells > c.cap) throw Error(PD_ERR_CUDA, "cell count unstable between passes");
            c.stats.rows += n;
            c.stats.lookups += lookups;
            c.stats.fwd_entries += fwd_entries;
            c.stats.pairs += pairs;
            c.stats.cells += cells;
        }
        rt::event_record(c.ev_call1, c.st);
        trace::Clock tp;
        rt::sync(c.st);
        tp.lap(trace::kCopyWait);
        rt::collect();
        tp.lap(trace::kTail);
        c.stats.total_ms = rt::event_ms(c.ev_call0, c.ev_call1);
        if (st_out) *st_out = c.stats;
    } catch (...) {
        release(cp);
        throw;
    }
    release(cp);
}

void Index::partition_rows(uint32_t parts, bool snap, uint32_t* bounds) {
    if (parts == 0) throw Error(PD_ERR_INVALID, "parts must be > 0");
    host_mirrors();
    const uint32_t S = info.S;
    // cost of a row = postings it visits, +1 so that empty rows still spread
    std::vector<uint64_t> pre((size_t)S + 1, 0);
    for (uint32_t s = 0; s < S; s++) pre[s + 1] = pre[s] + visited[s] + 1;
    bool contiguous = true;  // genomes occupy contiguous gene ranges?
    for (uint32_t s = 1; s < S && contiguous; s++)
        if (genome_of[s] < genome_of[s - 1]) contiguous = false;
    bounds[0] = 0;
    for (uint32_t p = 1; p < parts; p++) {
        const uint64_t target = pre[S] / parts * p + (pre[S] % parts) * p / parts;
        uint32_t b = (uint32_t)(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin());
        b = std::min(b, S);
        if (snap && contiguous && b > 0 && b < S) {
            // move to the nearest genome boundary
            uint32_t lo = b, hi = b;
            while (lo > 0 && genome_of[lo - 1] == genome_of[lo]) lo--;
            while (hi < S && genome_of[hi - 1] == genome_of[hi]) hi++;
            b = (pre[b] - pre[lo] <= pre[hi] - pre[b]) ? lo : hi;
        }
        bounds[p] = std::max(b, bounds[p - 1]);
    }
    bounds[parts] = S;
}

void Index::context_stats(ScoreContext* c, pd_score_stats* out) { *out = c->stats; }

void Index::entries(uint64_t* rank, uint32_t* seq, uint32_t* count, uint32_t* gs, uint32_t* gl) {
    rt::set_device(device);
    const uint32_t U = (uint32_t)info.U;
    if (!U) return;
    if (shard_world > 1) throw Error(PD_ERR_INVALID, "pd_entries is not available on a sharded index");
    if ((rank || gs || gl) && !opt.keep_sorted) throw Error(PD_ERR_INVALID, "pd_entries: ranks and groups need pd_options.keep_sorted");
    rt::stream_t st = rt::stream_create();
    if (seq) {
        rt::d2h(seq, post.p, sizeof(uint32_t) * U, st);
        rt::sync(st);
        for (uint32_t e = 0; e < U; e++) seq[e] &= 0x7FFFFFFFu;
    }
    if (count) {
        rt::d2h(count, post_cnt.p, sizeof(uint32_t) * U, st);
        rt::sync(st);
    }
    if (rank) {
        rt::d2h(rank, ent_rank.p, sizeof(uint64_t) * U, st);
        rt::sync(st);
    }
    if (gs || gl) {
        rt::DevBuf<uint32_t> a(U), b(U);
        PD_LAUNCH(ik::entry_groups_kernel, blocks_for(U), 256, 0, st, (const uint32_t*)ent_gid.p, (const uint32_t*)grp_head.p, U, a.p, b.p);
        if (gs) rt::d2h(gs, a.p, sizeof(uint32_t) * U, st);
        if (gl) rt::d2h(gl, b.p, sizeof(uint32_t) * U, st);
        rt::sync(st);
    }
    rt::stream_destroy(st);
}

}  // namespace pd
