"""Hand-written and seeded inputs shared by the CPU and GPU parity tests.

LITERAL: name -> (sequences, genome ids, k).  They cover the edge cases of SURVEY.md §7.2: the tail-group merge
(library.cpp:300-306), genes shorter than k, repeated k-mers inside a gene (counts > 1), duplicate genes,
non-standard residue letters (they enlarge the alphabet base, library.cpp:96-100), a single genome, genomes whose
genes are not contiguous, empty sequences.
"""
import numpy as np

from pandelos_b200 import synth

LITERAL = {
    # SURVEY.md §8c known answers: (0,1) 1/3, (0,2) 1/2 through the tail merge
    "tail_merge_kat": (["AAC", "AAA", "CC"], [0, 1, 2], 2),
    # last entry shares its rank with the run before it: no merge
    "tail_no_merge": (["AAC", "AAA", "AC"], [0, 1, 2], 2),
    "identical_pair": (["MKTAYIAKQRQISFVKSHFSRQ", "MKTAYIAKQRQISFVKSHFSRQ"], [0, 1], 4),
    "shorter_than_k": (["ACDEFGHIKL", "ACDEFGHIKL", "AC", "", "ACDEFAAAAAAA", "AAAAAAAAAA"], [0, 1, 1, 2, 2, 0], 3),
    "repeats": (["AAAAAAAAAAAAAAAA", "AAAAAAAACAAAAAAA", "ACACACACACACACAC", "CACACACACACACA", "AAAAACACACAC"], [0, 1, 2, 0, 1], 3),
    "nonstandard_letters": (["MKXUBZ*ACDEF", "MKXUBZ*ACDEG", "ACDEFGHIKLMNPQRSTVWY", "mkxubz*acdef"], [0, 1, 2, 3], 3),
    "single_genome": (["ACDEFGHIKL", "CDEFGHIKLM", "DEFGHIKLMN", "ACDEFGHIKL"], [0, 0, 0, 0], 4),
    "interleaved_genomes": (["ACDEFGHIKL", "CDEFGHIKLM", "ACDEFGHIKL", "DEFGHIKLMN", "CDEFGHIKLM", "ACDEFGHIKA"], [0, 1, 2, 0, 1, 2], 4),
    # genome 1 has no genes (ids stay below the gene count: the reference sizes genome_sequences by it, library.cpp:212)
    "sparse_genome_ids": (["ACDEFGHIKL", "CDEFGHIKLM", "ACDEFGHIKL", "CDEFGHIKLA"], [0, 2, 3, 2], 4),
    "one_gene": (["ACDEFGHIKL"], [0], 3),
    "all_short": (["AC", "A", ""], [0, 1, 2], 5),
    "k1": (["ACCA", "CAAC", "GG"], [0, 1, 2], 1),
}


def literal(name):
    seqs, genomes, k = LITERAL[name]
    return synth.from_sequences(seqs, genomes, name=name), k


def random_workload(seed, genes=60, genomes=4, max_len=40, alphabet="ACDE", k=3, shuffle_genomes=True):
    """Small random proteins over a tiny alphabet: dense k-mer sharing, many counts > 1."""
    rng = np.random.default_rng(seed)
    seqs, gids = [], []
    for i in range(genes):
        n = int(rng.integers(0, max_len))
        seqs.append("".join(rng.choice(list(alphabet), size=n)))
        gids.append(int(rng.integers(0, genomes)) if shuffle_genomes else i * genomes // genes)
    # make sure every genome id below `genomes` exists so G matches across implementations
    gids[0] = genomes - 1
    return synth.from_sequences(seqs, gids, name="random%d" % seed), k
