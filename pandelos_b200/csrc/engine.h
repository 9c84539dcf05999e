// Engine: owns the device-resident index (replaces `pair_info global_info`, reference ig/native/library.cpp:56-73)
// and the per-call scoring contexts.  Everything here throws pd::Error; c_api.cu maps that to status codes.
#pragma once

#include <condition_variable>
#include <mutex>
#include <vector>

#include "pandelos_b200.h"
#include "pd_rt.h"

namespace pd {

struct ScoreContext;

struct Index {
    int device = 0;
    pd_options opt;
    pd_index_info info;
    float thr = 0.f;            // 1/(2k) in float32 (library.cpp:499)
    int sms = 1;
    size_t smem_optin = 0;

    // device index
    rt::DevBuf<uint32_t> post;       // column gene | bit 31 (count > 1)
    rt::DevBuf<uint32_t> post_cnt;
    rt::DevBuf<uint2> fwd;
    rt::DevBuf<uint32_t> fwd_cnt;
    rt::DevBuf<uint32_t> fwd_ptr;              // S+1
    rt::DevBuf<unsigned long long> cls;        // S: forward entries per list class (3 x 21 bits)
    rt::DevBuf<unsigned long long> d_visited;  // S: total_visited
    rt::DevBuf<uint32_t> fam_key;              // S: row scheduling key (see gene_visited_kernel)
    rt::DevBuf<uint2> meta;                    // S: (kseq_len, genome)
    rt::DevBuf<uint32_t> ent_gid;    // kept only with opt.keep_sorted
    rt::DevBuf<uint32_t> grp_head;   // kept only with opt.keep_sorted
    rt::DevBuf<uint64_t> ent_rank;   // kept only with opt.keep_sorted
    rt::DevBuf<uint32_t> d_genome_rows;  // with genome_rows (below)
    rt::DevBuf<uint32_t> d_local_of;     // gene -> index inside its genome (flat_map of every genome, library.cpp:428-432)

    // host mirrors (O(S)), fetched from the device on first use: host_mirrors() / genome_lists()
    std::mutex mirror_mu;
    bool have_mirrors = false, have_genome_lists = false;
    std::vector<uint32_t> kseq;
    std::vector<uint32_t> genome_of;
    std::vector<uint64_t> visited;
    std::vector<uint32_t> genome_ptr;   // G+1
    rt::PinBuf<uint32_t> h_genome_rows;
    const uint32_t* genome_rows = nullptr;  // genes grouped by genome, input order inside (genome_sequences, library.cpp:245)
    void host_mirrors();
    void genome_lists();

    // contexts
    std::mutex mu;
    std::condition_variable cv;
    // Stage tokens of the per-genome calls.  Concurrent calls share the device-to-host copy engine by time slicing, so
    // calls that start copying together also finish together and the host gets every result late; the copy token makes
    // the copy stage first-come-first-served, and while one call sends its arrays the others compute.  The compute
    // token (off by default) does the same for the kernels; measured slower, because kernels of different calls fill
    // each other's tails.
    std::mutex compute_token, copy_token;
    int stage_tokens = 2;  // bit 0: compute stage, bit 1: copy stage (PD_STAGE_TOKENS)
    std::vector<ScoreContext*> free_ctx;
    std::vector<ScoreContext*> all_ctx;

    // ---- sharded build (several GPUs, one process each, build ONE index together: pd_build_shard / pd_shard_buffers /
    // pd_shard_finish).  Rank r sorts and groups the k-mers of the r-th slice of the rank space; the slices' postings are
    // all-gathered by the caller; forward lists are made for the rank's own query rows [own_row0, own_row1) only.
    uint32_t shard_rank = 0, shard_world = 1;
    uint32_t own_row0 = 0, own_row1 = 0, own_g0 = 0, own_g1 = 0;
    struct Shard {
        uint32_t U_r = 0, M_r = 0;                     // entries / repeated entries of this rank's slice
        uint64_t seg = 0, mseg = 0;                    // segment sizes of the gathered arrays
        rt::DevBuf<uint64_t> keys_a, keys_b;           // this rank's share grouped by destination slice (send); the keys of its slice (receive)
        uint64_t send_counts[32] = {0};                // keys of the share per destination rank
        uint64_t n_share = 0, n_recv = 0;
        int rank_bits = 0;
        rt::DevBuf<uint32_t> post_slice, heads_slice, multi_slice;
        rt::DevBuf<unsigned long long> gene_counts;    // [S] list-class counts, [S] total_visited: this slice's part, all-reduced in place by the caller
        rt::DevBuf<uint32_t> heads_all, multi_all, tile_heads;
        double ms = 0, fin_ms = 0;
        uint64_t launches = 0, fin_launches = 0;
        bool grouped = false, unusable = false;
    };
    Shard* shard = nullptr;
    uint64_t* shard_recv(uint64_t n_recv);
    void shard_sort(pd_shard_info* out);
    void shard_buffers(uint64_t max_entries, uint64_t max_multi, pd_shard_arrays* out);
    uint32_t* shard_multi(uint64_t max_multi);
    void shard_groups(const uint64_t* entries_of_rank, const uint64_t* multi_of_rank, uint32_t* bounds);
    void shard_finish();

    ~Index();
    void build(const uint8_t* residues, bool residues_on_device, const uint64_t* offsets, const uint32_t* genome_ids, uint32_t S,
               int32_t k, const pd_options* o);
    ScoreContext* acquire();
    void release(ScoreContext* c);
    static void release_context(ScoreContext* c);   // to the index the context belongs to
    void compute_scores(uint32_t genome, pd_scores* out);
    void genome_edges(uint32_t genome, pd_edges* out);
    void score_partition(uint32_t row_begin, uint32_t row_end, uint32_t rows_per_launch, float* d_best_hit, pd_score_stats* st);
    void partition_rows(uint32_t parts, bool snap, uint32_t* bounds);
    static void context_stats(ScoreContext* c, pd_score_stats* out);
    void entries(uint64_t* rank, uint32_t* seq, uint32_t* count, uint32_t* gs, uint32_t* gl);
};

void set_last_error(const std::string& s);
void trim_memory();

}  // namespace pd
