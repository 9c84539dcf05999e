// The plain C ABI of include/pandelos_b200.h: argument checks, error-code mapping, nothing else.
#include <cstring>
#include <new>
#include <thread>
#include <algorithm>
#include <string>
#include <vector>

#include "engine.h"

namespace pd {
const std::string& last_error();
}  // namespace pd

// The handle: one index, or (pd_options.devices > 1) one replica of it per device with the genomes dealt out to the
// devices in contiguous, posting-list-volume-balanced blocks.  Every per-genome call goes to the replica that owns the genome;
// calls for genomes of different devices run in parallel (the Java host calls computeScores from a thread pool,
// Pangenes.java:54-66).  The per-index calls (pd_info, pd_gene_stats, ...) are answered by the first replica.
struct pd_index {
    pd::Index ix;
    std::vector<pd::Index*> more;           // replicas on further devices
    std::vector<uint32_t> genome_owner;     // genome -> 0 (ix) or 1 + position in `more`
    std::vector<uint32_t> bounds;           // row bounds of the devices' blocks (devices + 1)
    ~pd_index() {
        for (pd::Index* x : more) delete x;
    }
    pd::Index& of_genome(uint32_t g) {
        const uint32_t o = g < genome_owner.size() ? genome_owner[g] : 0u;
        return o == 0 ? ix : *more[o - 1];
    }
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
    struct DeviceRestore {   // pd_options.device selects the index's device; the caller's current device is left as it was
        int dev = -1;
        DeviceRestore() { if (pd::rt::device_count() > 0) { try { dev = pd::rt::current_device(); } catch (...) { dev = -1; } } }
        ~DeviceRestore() { if (dev >= 0) { try { pd::rt::set_device(dev); } catch (...) {} } }
    } restore;
    int rc = guarded([&] {
        h = new pd_index;
        h->ix.shard_rank = rank;
        h->ix.shard_world = world;
        int devices = opt ? opt->devices : 0;
        const int have = pd::rt::device_count();
        if (devices > have) devices = have;
        if (devices <= 1 || world > 1) {
            h->ix.build(residues, on_device, offsets, genome_of, S, k, opt);
            return;
        }
        if (on_device) throw pd::Error(PD_ERR_INVALID, "pd_options.devices > 1 needs host residues (every device gets its own copy)");
        // one replica per device, built at the same time by one host thread each
        const int first = opt->device >= 0 ? opt->device : pd::rt::current_device();
        std::vector<std::string> errs((size_t)devices);
        std::vector<int> codes((size_t)devices, PD_OK);
        std::vector<std::thread> th;
        h->more.assign((size_t)devices - 1, nullptr);
        for (int d = 0; d < devices; d++) {
            th.emplace_back([&, d]() {
                try {
                    pd_options o = *opt;
                    o.device = (first + d) % have;
                    o.devices = 1;
                    if (d) o.verbose = 0;   // one cost report
                    pd::Index* x = d ? new pd::Index : &h->ix;
                    if (d) h->more[(size_t)d - 1] = x;
                    x->build(residues, false, offsets, genome_of, S, k, &o);
                } catch (const pd::Error& e) {
                    errs[(size_t)d] = e.what();
                    codes[(size_t)d] = e.code;
                } catch (const std::exception& e) {
                    errs[(size_t)d] = e.what();
                    codes[(size_t)d] = PD_ERR_CUDA;
                }
            });
        }
        for (std::thread& t : th) t.join();
        for (int d = 0; d < devices; d++)
            if (codes[(size_t)d] != PD_OK) throw pd::Error(codes[(size_t)d], errs[(size_t)d]);
        // genomes to devices: contiguous row blocks of equal posting-list volume, cut at genome boundaries
        h->bounds.assign((size_t)devices + 1, 0);
        h->ix.partition_rows((uint32_t)devices, true, h->bounds.data());
        h->ix.host_mirrors();
        h->genome_owner.assign(h->ix.info.G, 0);
        for (uint32_t s = 0; s < S; s++) {
            uint32_t d = 0;
            while (d + 1 < (uint32_t)devices && s >= h->bounds[d + 1]) d++;
            h->genome_owner[h->ix.genome_of[s]] = d;   // (genomes that are not contiguous gene ranges: the device of their last gene)
        }
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

int pd_shard_multi(pd_index* ix, uint64_t max_multi, uint32_t** d_multi) {
    if (!ix || !d_multi) return PD_ERR_INVALID;
    return guarded([&] { *d_multi = ix->ix.shard_multi(max_multi); });
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

void pd_trim(void) {
    try {
        pd::trim_memory();
    } catch (...) {
    }
}

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
    return guarded([&] { ix->of_genome(genome).compute_scores(genome, out); });
}

void pd_scores_release(pd_index* ix, pd_scores* s) {
    if (!ix || !s || !s->owner) return;
    pd::Index::release_context(static_cast<pd::ScoreContext*>(s->owner));
    memset(s, 0, sizeof(*s));
}

int pd_genome_edges(pd_index* ix, uint32_t genome, pd_edges* out) {
    if (!ix || !out) {
        pd::set_last_error("null argument");
        return PD_ERR_INVALID;
    }
    memset(out, 0, sizeof(*out));
    return guarded([&] { ix->of_genome(genome).genome_edges(genome, out); });
}

void pd_edges_release(pd_index* ix, pd_edges* e) {
    if (!ix || !e || !e->owner) return;
    pd::Index::release_context(static_cast<pd::ScoreContext*>(e->owner));
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
    if (ix->more.empty()) return guarded([&] { ix->ix.score_partition(row_begin, row_end, rows_per_launch, d_best_hit, stats); });
    // several devices: every device scores the part of the range that lies in its block, all at the same time
    return guarded([&] {
        if (d_best_hit) throw pd::Error(PD_ERR_INVALID, "pd_score_partition_device: no caller-side best-hit table with pd_options.devices > 1");
        const size_t D = ix->more.size() + 1;
        std::vector<pd_score_stats> st(D);
        std::vector<std::string> errs(D);
        std::vector<int> codes(D, PD_OK);
        std::vector<std::thread> th;
        for (size_t d = 0; d < D; d++) {
            memset(&st[d], 0, sizeof(pd_score_stats));
            const uint32_t b = std::max(row_begin, ix->bounds[d]), e = std::min(row_end, ix->bounds[d + 1]);
            if (b >= e) continue;
            th.emplace_back([&, d, b, e]() {
                try {
                    (d ? *ix->more[d - 1] : ix->ix).score_partition(b, e, rows_per_launch, nullptr, &st[d]);
                } catch (const pd::Error& x) {
                    errs[d] = x.what();
                    codes[d] = x.code;
                } catch (const std::exception& x) {
                    errs[d] = x.what();
                    codes[d] = PD_ERR_CUDA;
                }
            });
        }
        for (std::thread& t : th) t.join();
        for (size_t d = 0; d < D; d++)
            if (codes[d] != PD_OK) throw pd::Error(codes[d], errs[d]);
        if (stats) {
            memset(stats, 0, sizeof(*stats));
            for (size_t d = 0; d < D; d++) {
                stats->rows += st[d].rows; stats->lookups += st[d].lookups; stats->pairs += st[d].pairs; stats->cells += st[d].cells;
                stats->fallback_rows += st[d].fallback_rows; stats->launches += st[d].launches; stats->fwd_entries += st[d].fwd_entries;
                stats->retry_rows += st[d].retry_rows;
                stats->kernel_ms = std::max(stats->kernel_ms, st[d].kernel_ms);   // the devices run side by side
                stats->total_ms = std::max(stats->total_ms, st[d].total_ms);
            }
        }
    });
}

int pd_devices(const pd_index* ix) { return ix ? (int)ix->more.size() + 1 : 0; }

int pd_genome_device(const pd_index* ix, uint32_t genome) {
    if (!ix || genome >= ix->ix.info.G) return -1;
    return genome < ix->genome_owner.size() ? (int)ix->genome_owner[genome] : 0;
}

int pd_partition_rows(const pd_index* ix, uint32_t parts, int32_t snap_to_genomes, uint32_t* bounds) {
    if (!ix || !bounds) return PD_ERR_INVALID;
    return guarded([&] { const_cast<pd_index*>(ix)->ix.partition_rows(parts, snap_to_genomes != 0, bounds); });
}

}  // extern "C"
