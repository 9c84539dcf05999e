"""Shared parity checker: an engine `PangeneNative` against the CPU oracle (oracle/pangenes_oracle.c), bit for bit."""
import numpy as np

from oracle import cport
from pandelos_b200 import native


def assert_bits_equal(a, b, what):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype.itemsize == 4:
        a, b = a.view(np.uint32), b.view(np.uint32)
    bad = np.nonzero(a != b)[0] if a.ndim == 1 else np.argwhere(a != b)
    assert len(bad) == 0, "%s: %d mismatches, first at %s" % (what, len(bad), bad[0])


def check_index(pn, o):
    """k-mer counts, posting entries, groups (incl. the tail merge), per-gene cost model."""
    i = pn.info
    assert i.N == o.num_kmers
    assert i.U == o.num_entries
    assert i.base == o.base
    assert i.G == o.genomes
    rank, seq, cnt, gs, gl = pn.entries()
    orank, oseq, ocnt = o.entries()
    assert_bits_equal(rank, orank, "entry rank")
    assert_bits_equal(seq, oseq, "entry seq")
    assert_bits_equal(cnt, ocnt, "entry count")
    ogs, ogl = o.groups()
    assert_bits_equal(gs, ogs, "group start")
    assert_bits_equal(gl, ogl, "group length")
    kl, tv = pn.gene_stats()
    okl, otv = o.gene_stats()
    assert_bits_equal(kl, okl, "kseq_lengths")
    assert (tv == otv).all(), "total_visited"
    assert i.lookups == o.total_lookups


def check_scores(a, b, what=""):
    """Two Scores objects (engine vs oracle/reference): identical cell sets with identical float bits, identical
    best-hit tables.  Cell order is free (Pangenes.java:98-176 only takes max/min/set-insert)."""
    assert a.scoresCount == b.scoresCount, (what, a.scoresCount, b.scoresCount)
    ca, cb = a.canonical(), b.canonical()
    for f in native.Scores.FIELDS:
        assert_bits_equal(ca[f], cb[f], "%s %s" % (what, f))
    assert_bits_equal(a.max_genome_score, b.max_genome_score, what + " max_genome_score")
    assert_bits_equal(a.max_genome_score_col, b.max_genome_score_col, what + " max_genome_score_col")
    assert_bits_equal(a.scoresMaxMappings, b.scoresMaxMappings, what + " scoresMaxMappings")


def check_workload(w, k, genomes=None, index=True, **engine_kw):
    """Builds engine + oracle on the same packed input and compares everything; returns summary stats."""
    data = native.PangeneIData(w.residues, w.offsets, w.genome_of)
    pn = native.PangeneNative(k, data, keep_sorted=index, **engine_kw)
    o = cport.OracleIndex(w.residues, w.offsets, w.genome_of, k)
    try:
        if index:
            check_index(pn, o)
        else:
            assert pn.info.lookups == o.total_lookups
        cells = pairs = fallback = 0
        for g in (range(pn.info.G) if genomes is None else genomes):
            a = pn.generateScoresPart(g)
            b = o.compute_scores(g)
            check_scores(a, b, "genome %d" % g)
            assert pn.last_stats.pairs == o.candidate_pairs(g), "candidate pairs of genome %d" % g
            cells += a.scoresCount
            pairs += pn.last_stats.pairs
            fallback += pn.last_stats.fallback_rows
        return {"S": pn.info.S, "G": pn.info.G, "N": pn.info.N, "U": pn.info.U, "R": pn.info.R, "lookups": pn.info.lookups,
                "cells": cells, "pairs": pairs, "fallback_rows": fallback}
    finally:
        pn.close()
        o.close()
