// The plain C ABI of include/pandelos_b200.h: argument checks, error-code mapping, nothing else.
#include <cstring>
#include <new>

#include "engine.h"

namespace pd {
const std::string& last_error();
}  // namespace pd

struct pd_index {
    pd::Index ix;
};

namespace {

template <class F>
int guarded(F f) {
    try {
        f();
        return PD_OK;
    } catch (const pd::Error& e) {
        pd::set_last_error(e.what());
        return e.code;
    } catch (const std::bad_alloc&) {
        pd::set_last_error("out of host memory");
        return PD_ERR_NOMEM;
    } catch (const std::exception& e) {
        pd::set_last_error(e.what());
        return PD_ERR_CUDA;
    }
}

int build_common(const uint8_t* residues, bool on_device, const uint64_t* offsets, const uint32_t* genome_of, uint32_t S, int32_t k,
                 const pd_options* opt, pd_index** out, uint32_t rank = 0, uint32_t world = 1) {
    if (!out) {
        pd::set_last_error("null output pointer");
        return PD_ERR_INVALID;
    }
    *out = nullptr;
    pd_index* h = nullptr;
    int rc = guarded([&] {
        h = new pd_index;
        h->ix.shard_rank = rank;
        h->ix.shard_world = world;
        h->ix.build(residues, on_device, offsets, genome_of, S, k, opt);
    });
    if (rc != PD_OK) {
        delete h;
        return rc;
    }
    *out = h;
    return PD_OK;
}

}  // namespace

extern "C" {

const char* pd_last_error(void) { return pd::last_error().c_str(); }

int pd_device_count(void) { return pd::rt::device_count(); }

int pd_build(const uint8_t* residues, const uint64_t* offsets, const uint32_t* genome_of, uint32_t S, int32_t k, const pd_options* opt,
             pd_index** out) {
    return build_common(residues, false, offsets, genome_of, S, k, opt, out);
}

int pd_build_device(const uint8_t* d_residues, const uint64_t* offsets, const uint32_t* genome_of, uint32_t S, int32_t k,
                    const pd_options* opt, pd_index** out) {
    return build_common(d_residues, true, offsets, genome_of, S, k, opt, out);
}

int pd_build_shard(const uint8_t* residues, int32_t residues_on_device, const uint64_t* offsets, const uint32_t* genome_of, uint32_t S, int32_t k,
                   const pd_options* opt, uint32_t rank, uint32_t world, pd_index** out, pd_shard_keys* keys) {
    if (!keys || world < 2 || rank >= world) {
        pd::set_last_error("pd_build_shard: needs world >= 2, rank < world, keys");
        return PD_ERR_INVALID;
    }
    const int rc = build_common(residues, residues_on_device != 0, offsets, genome_of, S, k, opt, out, rank, world);
    if (rc != PD_OK) return rc;
    pd::Index& x = (*out)->ix;
    memset(keys, 0, sizeof(*keys));
    keys->d_send = x.shard->keys_a.p;
    for (uint32_t r = 0; r < world; r++) keys->send_counts[r] = x.shard->send_counts[r];
    return PD_OK;
}

int pd_shard_recv(pd_index* ix, uint64_t n_recv, uint64_t** d_recv) {
    if (!ix || !d_recv) return PD_ERR_INVALID;
    return guarded([&] { *d_recv = ix->ix.shard_recv(n_recv); });
}

int pd_shard_sort(pd_index* ix, pd_shard_info* info) {
    if (!ix || !info) return PD_ERR_INVALID;
    return guarded([&] { ix->ix.shard_sort(info); });
}

int pd_shard_buffers(pd_index* ix, uint64_t max_entries, uint64_t max_multi, pd_shard_arrays* out) {
    if (!ix || !out) return PD_ERR_INVALID;
    return guarded([&] { ix->ix.shard_buffers(max_entries, max_multi, out); });
}

int pd_shard_groups(pd_index* ix, const uint64_t* entries_of_rank, const uint64_t* multi_of_rank, uint32_t* bounds) {
    if (!ix || !entries_of_rank || !multi_of_rank || !bounds) return PD_ERR_INVALID;
    return guarded([&] { ix->ix.shard_groups(entries_of_rank, multi_of_rank, bounds); });
}

int pd_shard_finish(pd_index* ix) {
    if (!ix) return PD_ERR_INVALID;
    return guarded([&] { ix->ix.shard_finish(); });
}

void pd_free(pd_index* ix) { delete ix; }

int pd_info(const pd_index* ix, pd_index_info* out) {
    if (!ix || !out) return PD_ERR_INVALID;
    *out = ix->ix.info;
    return PD_OK;
}

int pd_gene_stats(const pd_index* ix, uint32_t* kseq_len, uint64_t* total_visited) {
    if (!ix) return PD_ERR_INVALID;
    return guarded([&] {
        pd::Index& x = const_cast<pd_index*>(ix)->ix;
        x.host_mirrors();
        const uint32_t S = x.info.S;
        if (kseq_len && S) memcpy(kseq_len, x.kseq.data(), sizeof(uint32_t) * S);
        if (total_visited && S) memcpy(total_visited, x.visited.data(), sizeof(uint64_t) * S);
    });
}

int pd_entries(const pd_index* ix, uint64_t* rank, uint32_t* seq, uint32_t* count, uint32_t* group_start, uint32_t* group_len) {
    if (!ix) return PD_ERR_INVALID;
    return guarded([&] { const_cast<pd_index*>(ix)->ix.entries(rank, seq, count, group_start, group_len); });
}

int pd_compute_scores(pd_index* ix, uint32_t genome, pd_scores* out) {
    if (!ix || !out) {
        pd::set_last_error("null argument");
        return PD_ERR_INVALID;
    }
    memset(out, 0, sizeof(*out));
    return guarded([&] { ix->ix.compute_scores(genome, out); });
}

void pd_scores_release(pd_index* ix, pd_scores* s) {
    if (!ix || !s || !s->owner) return;
    ix->ix.release(static_cast<pd::ScoreContext*>(s->owner));
    memset(s, 0, sizeof(*s));
}

int pd_genome_edges(pd_index* ix, uint32_t genome, pd_edges* out) {
    if (!ix || !out) {
        pd::set_last_error("null argument");
        return PD_ERR_INVALID;
    }
    memset(out, 0, sizeof(*out));
    return guarded([&] { ix->ix.genome_edges(genome, out); });
}

void pd_edges_release(pd_index* ix, pd_edges* e) {
    if (!ix || !e || !e->owner) return;
    ix->ix.release(static_cast<pd::ScoreContext*>(e->owner));
    memset(e, 0, sizeof(*e));
}

int pd_last_score_stats(const pd_scores* s, pd_score_stats* out) {
    if (!s || !s->owner || !out) return PD_ERR_INVALID;
    pd::Index::context_stats(static_cast<pd::ScoreContext*>(s->owner), out);
    return PD_OK;
}

int pd_score_partition_device(pd_index* ix, uint32_t row_begin, uint32_t row_end, uint32_t rows_per_launch, float* d_best_hit,
                              pd_score_stats* stats) {
    if (!ix) return PD_ERR_INVALID;
    return guarded([&] { ix->ix.score_partition(row_begin, row_end, rows_per_launch, d_best_hit, stats); });
}

int pd_partition_rows(const pd_index* ix, uint32_t parts, int32_t snap_to_genomes, uint32_t* bounds) {
    if (!ix || !bounds) return PD_ERR_INVALID;
    return guarded([&] { const_cast<pd_index*>(ix)->ix.partition_rows(parts, snap_to_genomes != 0, bounds); });
}

}  // extern "C"
