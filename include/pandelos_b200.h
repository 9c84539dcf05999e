/* pandelos_b200 — C ABI of the B200-native PanDelos `Pangenes` similarity engine.
 *
 * This is the drop-in boundary for ONE path of Guilucand/PanDelos: everything the reference's JNI library
 * `ig/native/library.cpp` does behind its two entry points (citations relative to the reference tree):
 *
 *   Java_infoasys_cli_pangenes_PangeneNative_preprocessSequences   ig/native/pangene_native.h:16-17, library.cpp:189-371
 *   Java_infoasys_cli_pangenes_PangeneNative_computeScores         ig/native/pangene_native.h:24-25, library.cpp:529-604
 *
 * `libnative.so` (csrc/jni_shim.cpp) exports exactly those two JNI symbols and forwards to the functions below;
 * the native `pangenes` CLI, the Python/ctypes binding and bench.py call them directly.  Plain pointers and sizes
 * only.  All work runs in hand-written sm_100a CUDA kernels; there is no CPU implementation behind this ABI —
 * without a CUDA device every entry point fails with PD_ERR_NO_DEVICE.
 *
 * Results are bit-identical to the reference: same k-mer ranks and counts, same candidate cells, same float32
 * score/perc/tr_perc bit patterns, same best-hit maxima.  The ORDER of the returned cells is unspecified (the
 * reference's consumers, Pangenes.java:98-176, only take max/min/set-insert over them).
 */
#ifndef PANDELOS_B200_H
#define PANDELOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PD_OK 0
#define PD_ERR_INVALID (-1)     /* bad argument (k <= 0, null pointer, unknown genome, offsets not ascending) */
#define PD_ERR_UNSUPPORTED (-2) /* base^k >= 2^63 (the reference's Rabin-hash fallback, library.cpp:81-86,103-119),
                                   >= 2^31 k-mers (the reference's own int limit, library.cpp:174,281,300),
                                   or a gene of 2^20 or more residues */
#define PD_ERR_CUDA (-3)        /* a CUDA call or kernel failed */
#define PD_ERR_NO_DEVICE (-4)   /* no usable CUDA device */
#define PD_ERR_NOMEM (-5)

typedef struct pd_index pd_index; /* the device-resident index: replaces `pair_info global_info` (library.cpp:56-73) */

typedef struct pd_options {
    int32_t device;          /* CUDA ordinal; -1 = the calling thread's current device */
    int32_t verbose;         /* 1 = print the reference's cost report (library.cpp:347-370) to stdout */
    int32_t contexts;        /* concurrent pd_compute_scores calls served without blocking (0 = default 4) */
    int32_t hash_log2;       /* log2 slots of the per-row shared-memory accumulator (0 = default 12) */
    uint64_t cell_capacity;  /* initial per-call cell buffer, in cells (0 = automatic) */
    int32_t keep_sorted;     /* 1 = keep the sorted k-mer keys so pd_entries can return ranks (tests) */
    int32_t devices;         /* > 1: one replica of the index on each of that many devices (starting at `device`), genomes dealt
                                out in posting-list-volume-balanced blocks; pd_compute_scores / pd_genome_edges go to the owner's
                                device, calls for different devices run side by side.  0 / 1 = one device.  Host residues only. */
    int32_t table_on_device; /* pd_build_device only: `offsets` and `genome_of` are device pointers as well */
    int32_t reserved;
} pd_options;

typedef struct pd_index_info {
    uint32_t S;            /* genes */
    uint32_t G;            /* genomes = max genome id + 1 (library.cpp:242) */
    int32_t k;
    uint32_t base;         /* alphabet size = distinct bytes present (library.cpp:96-100) */
    uint32_t rank_bits;    /* bits of a k-mer rank */
    uint32_t seq_bits;     /* bits of a gene id in the 64-bit sort key */
    uint64_t N;            /* k-mer occurrences */
    uint64_t U;            /* unique (k-mer, gene) entries = posting entries */
    uint64_t groups;       /* rank groups after the tail merge (library.cpp:300-306) */
    uint64_t R;            /* forward entries: (gene, shared k-mer) pairs */
    uint64_t lookups;      /* the reference's "Total cost: N lookups" (library.cpp:349) */
    uint32_t max_kseq;     /* longest gene, in k-mers */
    uint32_t reserved;
    double build_ms[8];    /* device time: histogram, encode, sort, groups, forward index, total, h2d, unused */
} pd_index_info;

/* Flat image of infoasys.cli.pangenes.Scores (ig/infoasys/cli/pangenes/Scores.java:3-35), one computeScores call.
 * All arrays are pinned host memory owned by the library until pd_scores_release. */
typedef struct pd_scores {
    int32_t scoresCount;
    int32_t S;                   /* length of scoresMaxMappings / max_genome_score_col */
    int32_t rows;                /* genes of this genome = rows of max_genome_score */
    int32_t G;
    float* scores;               /* Jaccard = sum(min)/sum(max), float32 (library.cpp:501) */
    float* percs;                /* library.cpp:497 */
    float* tr_percs;             /* library.cpp:498 */
    int32_t* row;
    int32_t* column;
    int32_t* first_seq_genome;
    int32_t* second_seq_genome;
    float* max_genome_score;     /* rows x G row-major: best score of the row against each genome (library.cpp:513-514) */
    float* max_genome_score_col; /* S: best score of each column against this genome's rows (library.cpp:515) */
    int32_t* scoresMaxMappings;  /* S: gene -> row index within this genome, INT32_MAX elsewhere (library.cpp:428-432) */
    void* owner;                 /* private */
} pd_scores;

/* Device-side timing/volume of the last scoring call on a context or of pd_score_partition_device. */
typedef struct pd_score_stats {
    uint64_t rows;
    uint64_t lookups;        /* posting entries visited */
    uint64_t pairs;          /* candidate (row, col != row) cells evaluated (finalize evaluations, library.cpp:493) */
    uint64_t cells;          /* non-zero cells emitted */
    uint64_t fallback_rows;  /* rows that overflowed the shared-memory accumulator and took the dense global path */
    uint64_t launches;       /* kernels launched */
    uint64_t fwd_entries;    /* (row, shared k-mer) forward entries read */
    uint64_t retry_rows;     /* rows that overflowed their first table and were re-run with the largest one */
    double kernel_ms;        /* CUDA-event time of the scoring kernels */
    double total_ms;         /* CUDA-event time of the whole call on its stream (memsets, kernels, copies) */
} pd_score_stats;

const char* pd_last_error(void);  /* message of the calling thread's last failure */
int pd_device_count(void);

/* preprocessSequences.  residues: concatenated sequence bytes (Latin-1 / ASCII, as `jchar < 256` in library.cpp:223);
 * offsets[S+1] ascending with offsets[0] == 0; genome_of[S].  Host pointers. */
int pd_build(const uint8_t* residues, const uint64_t* offsets, const uint32_t* genome_of, uint32_t S, int32_t k,
             const pd_options* opt, pd_index** out);
/* Same, with `residues` already in device memory (offsets / genome_of stay host pointers — O(S) metadata — unless
 * pd_options.table_on_device says they are resident too). */
int pd_build_device(const uint8_t* d_residues, const uint64_t* offsets, const uint32_t* genome_of, uint32_t S, int32_t k,
                    const pd_options* opt, pd_index** out);
void pd_free(pd_index* ix);
/* pd_free keeps the index's score contexts (streams, device and pinned result buffers) and its device blocks in per-device
 * caches for the next index (a new index per request is the expected use; PD_CACHE_KEEP_MB caps the idle device bytes).
 * pd_trim gives all of it back to the driver; call it when no index will be built for a while. */
void pd_trim(void);

int pd_info(const pd_index* ix, pd_index_info* out);
/* per gene: kseq_lengths (library.cpp:250-262) and computation_costs[].total_visited (library.cpp:327); either may be NULL */
int pd_gene_stats(const pd_index* ix, uint32_t* kseq_len, uint64_t* total_visited);
/* the sorted, count-deduplicated entry list (library.cpp:280-287) and each entry's group (start, len); any may be NULL.
 * `rank` needs pd_options.keep_sorted. */
int pd_entries(const pd_index* ix, uint64_t* rank, uint32_t* seq, uint32_t* count, uint32_t* group_start, uint32_t* group_len);

/* computeScores(genome).  Thread-safe; blocks while all contexts are in use. */
int pd_compute_scores(pd_index* ix, uint32_t genome, pd_scores* out);
void pd_scores_release(pd_index* ix, pd_scores* s);
int pd_last_score_stats(const pd_scores* s, pd_score_stats* out);

/* The whole per-genome task of the reference's Java host on the device (Pangenes.java:60-183): computeScores(genome),
 * the inter-genome bidirectional-best-hit test (:98-128), the per-row threshold (:146-155) and the intra-genome
 * (paralog) test (:164-176).  Returns the arguments of the addConnection calls the Java code would make: one entry
 * (src, dst, score) per inter-genome BBH cell (src = the row gene of `genome`; Java also adds the mirrored (dst, src),
 * which the other genome's call returns too) and per intra-genome edge (src < dst).  The cells themselves stay in HBM.
 * Arrays are pinned host memory owned by the library until pd_edges_release.  Thread-safe like pd_compute_scores. */
typedef struct pd_edges {
    uint64_t count;
    const uint32_t* src;
    const uint32_t* dst;
    const float* score;
    uint64_t cells;   /* non-zero cells the filter looked at (= scoresCount of computeScores) */
    void* owner;      /* private */
} pd_edges;
int pd_genome_edges(pd_index* ix, uint32_t genome, pd_edges* out);
void pd_edges_release(pd_index* ix, pd_edges* e);

/* Device-resident scoring of the gene range [row_begin, row_end) (multi-GPU partitions, bench `value`):
 * cells stay in the context's HBM buffers; d_best_hit (device, (row_end-row_begin) x G floats, may be NULL)
 * receives BH[r][h] = best score of gene r against genome h.  Rows are processed in blocks of `rows_per_launch`
 * genes (0 = automatic), each block's cells overwriting the previous block's. */
int pd_score_partition_device(pd_index* ix, uint32_t row_begin, uint32_t row_end, uint32_t rows_per_launch,
                              float* d_best_hit, pd_score_stats* stats);

/* ---- One index built by several GPUs (one process per GPU; the reference has no counterpart: its preprocessSequences is
 * single-threaded, library.cpp:189-371).  The rank space of the k-mers is cut into `world` slices.  Every rank
 *   1. pd_build_shard: makes the k-mers of ITS SHARE of the genes (about 1 / world of the residues) and groups them by the
 *      slice they fall in (one stable pass).  The caller all-gathers pd_shard_keys.send_counts, gets room for its slice with
 *      pd_shard_recv(sum of what the others hold for it) and moves the keys with one all-to-all (8 B per k-mer; blocks
 *      arrive in source-rank order = gene order),
 *   2. pd_shard_sort: sorts, count-dedups and groups the k-mers of its slice.  The caller all-gathers (entries, multi) of
 *      pd_shard_info over the ranks and all-reduces (sum, uint64) pd_shard_info.d_gene_counts [2 x S] in place,
 *   3. calls pd_shard_buffers(max entries, max multi) and all-gathers each of the three arrays IN PLACE (segment r of each
 *      array is rank r's, already filled on that rank); the postings — by far the largest — may still be in flight during 4a,
 *   4a. calls pd_shard_groups: multiplicities, group structure of the whole entry list from the head bits, genome-aligned
 *      query partition by posting-list volume (bounds[world + 1], identical on every rank),
 *   4b. once the postings have arrived, calls pd_shard_finish: forward lists for THIS rank's rows [bounds[rank],
 *      bounds[rank + 1]).
 * Afterwards pd_compute_scores / pd_genome_edges / pd_score_partition_device serve this rank's genomes / rows only
 * (PD_ERR_INVALID for others), with results bit-identical to a single-GPU index.  Needs the genes of a genome to be
 * contiguous and genomes in ascending order (as in every .faa PanDelos reads), at most 32 ranks; pd_entries is not
 * available.  pandelos_b200/multigpu.py (build_sharded) drives these calls over torch.distributed. */
typedef struct pd_shard_keys {
    uint64_t* d_send;           /* device: this rank's keys grouped by destination rank */
    uint64_t send_counts[32];   /* keys per destination rank */
} pd_shard_keys;
typedef struct pd_shard_info {
    uint64_t entries;         /* entries of this rank's slice */
    uint64_t multi;           /* ... of which held more than once by their gene */
    uint64_t kmers;           /* k-mer occurrences in this rank's slice */
    uint64_t* d_gene_counts;  /* device, 2 x S: per gene forward entries per list class (packed) and total_visited — partial */
} pd_shard_info;
typedef struct pd_shard_arrays {
    uint32_t* d_post;   /* world x seg  : postings */
    uint32_t* d_heads;  /* world x seg / 32 : one bit per entry, set where a rank group starts */
    uint32_t* d_multi;  /* world x mseg x 2 : (entry inside its slice, multiplicity) */
    uint64_t seg, mseg;
} pd_shard_arrays;
int pd_build_shard(const uint8_t* residues, int32_t residues_on_device, const uint64_t* offsets, const uint32_t* genome_of, uint32_t S,
                   int32_t k, const pd_options* opt, uint32_t rank, uint32_t world, pd_index** out, pd_shard_keys* keys);
int pd_shard_recv(pd_index* ix, uint64_t n_recv, uint64_t** d_recv);
int pd_shard_sort(pd_index* ix, pd_shard_info* info);
int pd_shard_buffers(pd_index* ix, uint64_t max_entries, uint64_t max_multi, pd_shard_arrays* out);
/* pd_shard_buffers may be called with max_multi = 0 as soon as an upper bound of the slices' entries is known (a slice has at
 * most as many entries as it received k-mers), so that the all-gather of the postings starts before the exact counts have been
 * exchanged; pd_shard_multi then makes the third array (world x max_multi x 2) once the list sizes are known. */
int pd_shard_multi(pd_index* ix, uint64_t max_multi, uint32_t** d_multi);
int pd_shard_groups(pd_index* ix, const uint64_t* entries_of_rank, const uint64_t* multi_of_rank, uint32_t* bounds);
int pd_shard_finish(pd_index* ix);

/* pd_options.devices > 1: number of replicas, and which of them (0 .. pd_devices - 1) serves a genome; a host that issues its
 * per-genome calls device by device keeps all devices busy. */
int pd_devices(const pd_index* ix);
int pd_genome_device(const pd_index* ix, uint32_t genome);

/* Splits [0, S) into `parts` contiguous gene ranges of near-equal total_visited (query partitioning by
 * posting-list volume); bounds[parts+1].  snap_to_genomes != 0 moves boundaries to genome boundaries. */
int pd_partition_rows(const pd_index* ix, uint32_t parts, int32_t snap_to_genomes, uint32_t* bounds);

#ifdef __cplusplus
}
#endif
#endif
