/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  CPU restatement ("port") of the PanDelos Pangenes similarity
 * hot path, used only as the parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * See pangenes_oracle.c for the reference file:line each function follows.
 *
 * Parity pin: the reference ships no golden vectors or tests for this path, so this
 * restatement is pinned against the UNMODIFIED reference library (oracle/_ref/libnative_ref.so driven through
 * oracle/fakejni.cpp) on every fixture in tests/golden/ and on seeded random inputs (tests/test_oracle.py). */
#ifndef PANGENES_ORACLE_H
#define PANGENES_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct po_index po_index;

/* Same field layout as fj_scores in fakejni.cpp and as infoasys.cli.pangenes.Scores (Scores.java:3-35). */
typedef struct po_scores {
    int32_t scoresCount;
    int32_t S;
    int32_t rows;
    int32_t G;
    float* scores;
    float* percs;
    float* tr_percs;
    int32_t* row;
    int32_t* column;
    int32_t* first_seq_genome;
    int32_t* second_seq_genome;
    float* max_genome_score;     /* rows x G */
    float* max_genome_score_col; /* S */
    int32_t* scoresMaxMappings;  /* S */
} po_scores;

/* returns NULL (and prints why) for k <= 0 or base^k not representable (the reference's Rabin-hash
 * fallback, library.cpp:81-86,112-118, is out of scope) */
po_index* po_build(const uint8_t* residues, const uint64_t* offsets, const uint32_t* genome_of, uint32_t S, int32_t k);
void po_free(po_index* ix);

uint32_t po_genomes(const po_index* ix);
uint32_t po_alphabet_base(const po_index* ix);
uint64_t po_num_kmers(const po_index* ix);      /* N: k-mer occurrences */
uint64_t po_num_entries(const po_index* ix);    /* U: unique (k-mer, gene) entries */
uint64_t po_total_lookups(const po_index* ix);  /* the reference's "Total cost: N lookups" (library.cpp:349) */
/* copies out the sorted, count-deduplicated entry list (any pointer may be NULL) */
void po_entries(const po_index* ix, uint64_t* rank, uint32_t* seq, uint32_t* count);
/* per entry: first entry index and length of its shared-k-mer group after the tail-merge quirk (len 1 = unshared) */
void po_groups(const po_index* ix, uint32_t* group_start, uint32_t* group_len);
void po_gene_stats(const po_index* ix, uint32_t* kseq_len, uint64_t* total_visited);

po_scores* po_compute_scores(const po_index* ix, uint32_t genome);
void po_scores_free(po_scores* s);
/* number of (row, col != row) cells touched while scoring this genome (finalize evaluations, library.cpp:493) */
uint64_t po_candidate_pairs(const po_index* ix, uint32_t genome);

#ifdef __cplusplus
}
#endif
#endif
