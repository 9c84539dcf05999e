"""Order-independent digests of `Scores` objects (reference ig/infoasys/cli/pangenes/Scores.java:3-35).

The order of the cells a computeScores call returns is not part of the contract (every consumer in
Pangenes.java:98-176 takes max / min / set-insert over them), so two implementations are compared on the multiset of
cells: every cell (row, column, score bits, perc bits, tr_perc bits, first_seq_genome, second_seq_genome) is hashed to
64 bits and the hashes are summed and xor-ed.  The three tables (max_genome_score, max_genome_score_col,
scoresMaxMappings) are position-dependent and go through sha256.

Pure numpy; used by tests/, bench.py's parity legs and the golden-vector scripts on Scores objects of the engine
(native.Scores), of the oracle port and of the unmodified reference library alike.
"""
import hashlib

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix(x):
    x = x ^ (x >> np.uint64(30))
    x = x * _M1
    x = x ^ (x >> np.uint64(27))
    x = x * _M2
    return x ^ (x >> np.uint64(31))


def _u64(a, n):
    a = np.ascontiguousarray(a[:n])
    if a.dtype == np.float32:
        a = a.view(np.uint32)
    return a.astype(np.int64).astype(np.uint64)


def cells_digest(s):
    """(count, sum of the cell hashes mod 2^64, xor of the cell hashes)."""
    n = int(s.scoresCount)
    if n == 0:
        return 0, 0, 0
    with np.errstate(over="ignore"):
        h = _mix(_u64(s.row, n) * _G + _u64(s.column, n))
        h = _mix(h + (_u64(s.scores, n) << np.uint64(32)) + _u64(s.percs, n))
        h = _mix(h + (_u64(s.tr_percs, n) << np.uint64(32)) + (_u64(s.first_seq_genome, n) << np.uint64(16)) + _u64(s.second_seq_genome, n))
        return n, int(h.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(h))


def tables_digest(s):
    """sha256 over the bit patterns of max_genome_score, max_genome_score_col and scoresMaxMappings."""
    h = hashlib.sha256()
    for a in (s.max_genome_score, s.max_genome_score_col, s.scoresMaxMappings):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def scores_digest(s):
    """JSON-ready digest of every field of one computeScores result."""
    n, hs, hx = cells_digest(s)
    return {"cells": n, "sum": "%016x" % hs, "xor": "%016x" % hx, "tables": tables_digest(s)}
