/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Plain-C restatement of the PanDelos `Pangenes` similarity hot path as implemented by the reference's JNI
 * library (all citations relative to /root/reference):
 *
 *   ig/native/library.cpp:189-371   preprocessSequences : alphabet ranks, k-mer ranks, sort, count-dedup, groups
 *   ig/native/library.cpp:409-527   computeScores       : accumulate along posting lists, Jaccard, best hits
 *   ig/native/library.cpp:529-604   JNI marshalling     : the ten Scores fields
 *
 * It is written for readability, not speed: comparison sort instead of LSD radix, direct polynomial k-mer
 * value instead of the rolling update, dense per-row accumulators with a touched list.  It reproduces every
 * result-affecting behaviour of the reference, including the tail-group merge (library.cpp:300-306) and the
 * float32 arithmetic of the finalize step (library.cpp:494-502).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Parity pin: checked against the unmodified reference library on every fixture (tests/test_oracle.py);
 * the reference itself ships no golden vectors for this path.
 */
#include "pangenes_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t rank;
    uint32_t seq;
    uint32_t count;
} entry_t;

struct po_index {
    uint32_t S, G;
    int32_t k;
    uint32_t base;
    uint8_t val[256];
    uint64_t N, U;
    entry_t* e;            /* U entries sorted by (rank, seq) */
    uint32_t* grp_start;   /* per entry */
    uint32_t* grp_len;     /* per entry */
    uint32_t* kseq;        /* S : k-mer count per gene incl. repeats (kseq_lengths, library.cpp:250-262) */
    uint32_t* genome_of;   /* S */
    uint64_t* visited;     /* S : total_visited (library.cpp:327) */
    uint64_t* own_ptr;     /* S+1 : CSR of each gene's entries that sit in shared groups */
    uint32_t* own_idx;     /* entry indices, ascending */
    uint64_t lookups;
};

static int cmp_entry(const void* a, const void* b) {
    const entry_t* x = (const entry_t*)a;
    const entry_t* y = (const entry_t*)b;
    if (x->rank != y->rank) return x->rank < y->rank ? -1 : 1;
    if (x->seq != y->seq) return x->seq < y->seq ? -1 : 1;
    return 0;
}

po_index* po_build(const uint8_t* residues, const uint64_t* offsets, const uint32_t* genome_of, uint32_t S, int32_t k) {
    /* library.cpp:90-93 */
    if (k <= 0) {
        fprintf(stderr, "oracle: K value must be greater than 0.\n");
        return NULL;
    }
    po_index* ix = (po_index*)calloc(1, sizeof(po_index));
    ix->S = S;
    ix->k = k;

    /* alphabet: dense rank, in ascending byte order, of every byte present anywhere (library.cpp:96-100,216-228) */
    uint64_t hist[256] = {0};
    for (uint64_t i = offsets[0]; i < offsets[S]; i++) hist[residues[i]]++;
    uint32_t base = 0;
    for (int b = 0; b < 256; b++)
        if (hist[b]) ix->val[b] = (uint8_t)base++;
    ix->base = base;

    /* base^k must fit: the reference switches to a Rabin hash otherwise (library.cpp:103-119) — out of scope */
    {
        unsigned __int128 p = 1;
        for (int i = 0; i < k; i++) {
            p *= (base ? base : 1);
            if (p >> 63) {
                fprintf(stderr, "oracle: base^k overflows (base=%u k=%d); hash fallback unsupported\n", base, k);
                free(ix);
                return NULL;
            }
        }
    }

    ix->kseq = (uint32_t*)calloc(S + 1, sizeof(uint32_t));
    ix->genome_of = (uint32_t*)calloc(S + 1, sizeof(uint32_t));
    uint64_t N = 0;
    uint32_t G = 0;
    for (uint32_t s = 0; s < S; s++) {
        int64_t len = (int64_t)(offsets[s + 1] - offsets[s]);
        int64_t kl = len - k + 1; /* library.cpp:250 */
        ix->kseq[s] = kl > 0 ? (uint32_t)kl : 0;
        N += ix->kseq[s];
        ix->genome_of[s] = genome_of[s];
        if (genome_of[s] + 1 > G) G = genome_of[s] + 1; /* library.cpp:242 */
    }
    ix->G = G;
    ix->N = N;

    /* one (rank, seq, 1) per position; rank = sum val[c_j] * base^(k-1-j) (library.cpp:75-79,134-150) */
    entry_t* e = (entry_t*)malloc(sizeof(entry_t) * (N + 1));
    uint64_t n = 0;
    for (uint32_t s = 0; s < S; s++) {
        const uint8_t* p = residues + offsets[s];
        for (uint32_t i = 0; i < ix->kseq[s]; i++) {
            uint64_t r = 0;
            for (int j = 0; j < k; j++) r = r * base + ix->val[p[i + j]];
            e[n].rank = r;
            e[n].seq = s;
            e[n].count = 1;
            n++;
        }
    }
    /* total order by (rank, seq) (library.cpp:270-278), then merge duplicates into counts (library.cpp:280-287) */
    qsort(e, N, sizeof(entry_t), cmp_entry);
    uint64_t U = 0;
    for (uint64_t i = 0; i < N; i++) {
        if (U && e[U - 1].rank == e[i].rank && e[U - 1].seq == e[i].seq)
            e[U - 1].count++;
        else
            e[U++] = e[i];
    }
    ix->U = U;
    ix->e = e;

    /* groups = runs of equal rank; the final entry, when alone in its rank, joins the run before it
     * (library.cpp:300-306: the `|| last` branch closes [current_rank_start, i+1) whatever the last rank is) */
    ix->grp_start = (uint32_t*)calloc(U + 1, sizeof(uint32_t));
    ix->grp_len = (uint32_t*)calloc(U + 1, sizeof(uint32_t));
    {
        uint64_t start = 0;
        for (uint64_t i = 1; i <= U; i++) {
            int boundary = (i == U) || (e[i].rank != e[i - 1].rank);
            if (boundary && i == U - 1) boundary = 0; /* tail merge */
            if (boundary) {
                for (uint64_t j = start; j < i; j++) {
                    ix->grp_start[j] = (uint32_t)start;
                    ix->grp_len[j] = (uint32_t)(i - start);
                }
                start = i;
            }
        }
    }

    /* per gene: its entries that are in groups of >= 2, and the cost model (library.cpp:308-330) */
    ix->visited = (uint64_t*)calloc(S + 1, sizeof(uint64_t));
    ix->own_ptr = (uint64_t*)calloc((size_t)S + 2, sizeof(uint64_t));
    for (uint64_t i = 0; i < U; i++)
        if (ix->grp_len[i] > 1) {
            ix->own_ptr[e[i].seq + 1]++;
            ix->visited[e[i].seq] += ix->grp_len[i];
            ix->lookups += ix->grp_len[i];
        }
    for (uint32_t s = 0; s < S; s++) ix->own_ptr[s + 1] += ix->own_ptr[s];
    ix->own_idx = (uint32_t*)malloc(sizeof(uint32_t) * (ix->own_ptr[S] + 1));
    {
        uint64_t* cur = (uint64_t*)malloc(sizeof(uint64_t) * (S + 1));
        memcpy(cur, ix->own_ptr, sizeof(uint64_t) * (S + 1));
        for (uint64_t i = 0; i < U; i++)
            if (ix->grp_len[i] > 1) ix->own_idx[cur[e[i].seq]++] = (uint32_t)i;
        free(cur);
    }
    return ix;
}

void po_free(po_index* ix) {
    if (!ix) return;
    free(ix->e); free(ix->grp_start); free(ix->grp_len); free(ix->kseq); free(ix->genome_of);
    free(ix->visited); free(ix->own_ptr); free(ix->own_idx);
    free(ix);
}

uint32_t po_genomes(const po_index* ix) { return ix->G; }
uint32_t po_alphabet_base(const po_index* ix) { return ix->base; }
uint64_t po_num_kmers(const po_index* ix) { return ix->N; }
uint64_t po_num_entries(const po_index* ix) { return ix->U; }
uint64_t po_total_lookups(const po_index* ix) { return ix->lookups; }

void po_entries(const po_index* ix, uint64_t* rank, uint32_t* seq, uint32_t* count) {
    for (uint64_t i = 0; i < ix->U; i++) {
        if (rank) rank[i] = ix->e[i].rank;
        if (seq) seq[i] = ix->e[i].seq;
        if (count) count[i] = ix->e[i].count;
    }
}

void po_groups(const po_index* ix, uint32_t* group_start, uint32_t* group_len) {
    if (group_start) memcpy(group_start, ix->grp_start, sizeof(uint32_t) * ix->U);
    if (group_len) memcpy(group_len, ix->grp_len, sizeof(uint32_t) * ix->U);
}

void po_gene_stats(const po_index* ix, uint32_t* kseq_len, uint64_t* total_visited) {
    if (kseq_len) memcpy(kseq_len, ix->kseq, sizeof(uint32_t) * ix->S);
    if (total_visited) memcpy(total_visited, ix->visited, sizeof(uint64_t) * ix->S);
}

typedef struct {
    float score, perc, tr_perc;
    uint32_t x, y;
} cell_t;

static int cmp_u32(const void* a, const void* b) {
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x < y ? -1 : x > y;
}

static po_scores* score_genome(const po_index* ix, uint32_t genome, uint64_t* pairs_out) {
    const uint32_t S = ix->S, G = ix->G;
    po_scores* out = (po_scores*)calloc(1, sizeof(po_scores));
    out->S = (int32_t)S;
    out->G = (int32_t)G;

    /* rows of this genome in input order, flat_map (library.cpp:428-432; INT32_MAX elsewhere) */
    out->scoresMaxMappings = (int32_t*)malloc(sizeof(int32_t) * (S + 1));
    uint32_t rows = 0;
    for (uint32_t s = 0; s < S; s++) out->scoresMaxMappings[s] = ix->genome_of[s] == genome ? (int32_t)rows++ : INT32_MAX;
    out->rows = (int32_t)rows;
    out->max_genome_score = (float*)calloc((size_t)rows * G + 1, sizeof(float));
    out->max_genome_score_col = (float*)calloc(S + 1, sizeof(float));

    int32_t* inter = (int32_t*)calloc(S + 1, sizeof(int32_t));
    int32_t* pc = (int32_t*)calloc(S + 1, sizeof(int32_t));
    int32_t* tc = (int32_t*)calloc(S + 1, sizeof(int32_t));
    uint8_t* seen = (uint8_t*)calloc(S + 1, 1);
    uint32_t* touched = (uint32_t*)malloc(sizeof(uint32_t) * (S + 1));

    size_t cap = 1024, nc = 0;
    cell_t* cells = (cell_t*)malloc(sizeof(cell_t) * cap);
    uint64_t pairs = 0;

    for (uint32_t r = 0; r < S; r++) {
        if (ix->genome_of[r] != genome) continue;
        uint32_t nt = 0;
        /* every entry of r in a shared group, against every entry of that group (library.cpp:461-479) */
        for (uint64_t o = ix->own_ptr[r]; o < ix->own_ptr[r + 1]; o++) {
            uint32_t mine = ix->own_idx[o];
            uint32_t my_cnt = ix->e[mine].count;
            uint32_t gs = ix->grp_start[mine], gl = ix->grp_len[mine];
            for (uint32_t j = gs; j < gs + gl; j++) {
                uint32_t c = ix->e[j].seq, cnt = ix->e[j].count;
                if (!seen[c]) {
                    seen[c] = 1;
                    touched[nt++] = c;
                }
                inter[c] += (int32_t)(cnt < my_cnt ? cnt : my_cnt);
                pc[c] += (int32_t)my_cnt;
                tc[c] += (int32_t)cnt;
            }
        }
        /* identity cell cleared (library.cpp:485-487) */
        inter[r] = 0; pc[r] = 0; tc[r] = 0;
        qsort(touched, nt, sizeof(uint32_t), cmp_u32); /* cell order is not part of the contract; ascending here */
        for (uint32_t t = 0; t < nt; t++) {
            uint32_t c = touched[t];
            if (c != r) pairs++;
            /* library.cpp:494-502, float32 throughout */
            int32_t my_k = (int32_t)ix->kseq[r], other_k = (int32_t)ix->kseq[c];
            int32_t uni = my_k + other_k - inter[c];
            float perc = (float)pc[c] / (float)my_k;
            float tr_perc = (float)tc[c] / (float)other_k;
            float thr = 1.0f / (2.0f * (float)ix->k);
            int valid = perc >= thr || tr_perc >= thr;
            float score = (float)inter[c] / (float)uni * (valid ? 1.0f : 0.0f);
            if (score > 0.0f) {
                if (nc == cap) {
                    cap *= 2;
                    cells = (cell_t*)realloc(cells, sizeof(cell_t) * cap);
                }
                cells[nc].score = score; cells[nc].perc = perc; cells[nc].tr_perc = tr_perc;
                cells[nc].x = r; cells[nc].y = c;
                nc++;
                /* library.cpp:513-515 */
                float* m = &out->max_genome_score[(size_t)out->scoresMaxMappings[r] * G + ix->genome_of[c]];
                if (score > *m) *m = score;
                if (score > out->max_genome_score_col[c]) out->max_genome_score_col[c] = score;
            }
            inter[c] = 0; pc[c] = 0; tc[c] = 0; seen[c] = 0;
        }
    }
    if (pairs_out) *pairs_out = pairs;

    out->scoresCount = (int32_t)nc;
    out->scores = (float*)malloc(sizeof(float) * (nc + 1));
    out->percs = (float*)malloc(sizeof(float) * (nc + 1));
    out->tr_percs = (float*)malloc(sizeof(float) * (nc + 1));
    out->row = (int32_t*)malloc(sizeof(int32_t) * (nc + 1));
    out->column = (int32_t*)malloc(sizeof(int32_t) * (nc + 1));
    out->first_seq_genome = (int32_t*)malloc(sizeof(int32_t) * (nc + 1));
    out->second_seq_genome = (int32_t*)malloc(sizeof(int32_t) * (nc + 1));
    for (size_t i = 0; i < nc; i++) {
        out->scores[i] = cells[i].score; out->percs[i] = cells[i].perc; out->tr_percs[i] = cells[i].tr_perc;
        out->row[i] = (int32_t)cells[i].x; out->column[i] = (int32_t)cells[i].y;
        out->first_seq_genome[i] = (int32_t)ix->genome_of[cells[i].x];   /* library.cpp:571-575 */
        out->second_seq_genome[i] = (int32_t)ix->genome_of[cells[i].y];
    }
    free(cells); free(inter); free(pc); free(tc); free(seen); free(touched);
    return out;
}

po_scores* po_compute_scores(const po_index* ix, uint32_t genome) { return score_genome(ix, genome, NULL); }

uint64_t po_candidate_pairs(const po_index* ix, uint32_t genome) {
    uint64_t p = 0;
    po_scores_free(score_genome(ix, genome, &p));
    return p;
}

void po_scores_free(po_scores* s) {
    if (!s) return;
    free(s->scores); free(s->percs); free(s->tr_percs);
    free(s->row); free(s->column); free(s->first_seq_genome); free(s->second_seq_genome);
    free(s->max_genome_score); free(s->max_genome_score_col); free(s->scoresMaxMappings);
    free(s);
}
