"""Host logic of the multi-GPU path (one process per GPU, `torch.distributed`): the index is built on every rank
("built once and replicated": the build is O(N) and cheap next to O(L) scoring), the query genes are block-partitioned
by posting-list volume, and the best-hit table slices are all-gathered (NCCL over NVLink on GPUs; gloo in the CPU tests).

The reference has no counterpart: it parallelises `computeScores` over genomes on one JVM thread pool
(reference ig/infoasys/cli/pangenes/Pangenes.java:54-66).  The unit handed to a rank is the same — whole genomes —
so every `computeScores(g)` of the reference maps to exactly one rank and no cell is ever exchanged; only
BH[r][h] (library.cpp:513-514) is gathered, because colmax_g[c] == BH[c][g] (score symmetry).
"""
import numpy as np


def genome_bounds(genome_of, G):
    """First gene of every genome (genes of a genome contiguous, as in every .faa PanDelos reads); length G + 1."""
    genome_of = np.asarray(genome_of)
    if len(genome_of) > 1 and (np.diff(genome_of.astype(np.int64)) < 0).any():
        raise ValueError("genes of a genome must be contiguous for row partitioning")
    return np.searchsorted(genome_of, np.arange(G + 1), side="left").astype(np.int64)


def balanced_bounds(visited, row_begin, row_end, parts, snap=None):
    """Splits genes [row_begin, row_end) into `parts` contiguous ranges of near-equal posting-list volume
    (sum of total_visited + 1 per gene, the reference's own cost model, library.cpp:327).  `snap`: sorted array of
    allowed boundaries (genome starts); each cut moves to the nearest one.  Returns int64[parts + 1]."""
    cost = np.asarray(visited[row_begin:row_end], dtype=np.float64) + 1.0
    pre = np.concatenate([[0.0], np.cumsum(cost)])
    out = np.empty(parts + 1, np.int64)
    out[0], out[parts] = row_begin, row_end
    for p in range(1, parts):
        b = row_begin + int(np.searchsorted(pre, pre[-1] * p / parts, side="left"))
        b = min(max(b, row_begin), row_end)
        if snap is not None:
            cand = snap[(snap >= row_begin) & (snap <= row_end)]
            j = int(np.searchsorted(cand, b))
            lo = cand[max(j - 1, 0)]
            hi = cand[min(j, len(cand) - 1)]
            b = int(lo if (b - lo) <= (hi - b) else hi)
        out[p] = max(b, out[p - 1])
    return out


def allgather_best_hits(dist, bh_local, rows_of_rank, G, device):
    """All-gathers the ranks' best-hit slices (rows_of_rank[r] x G float32, row-partitioned) into one
    (sum(rows) x G) tensor on every rank.  Slices are padded to the largest one (all_gather_into_tensor needs equal
    shapes); `bh_local` must already have max(rows_of_rank) rows.  Returns (bh_all, padded buffer)."""
    import torch
    world = len(rows_of_rank)
    max_rows = int(max(rows_of_rank)) if world else 0
    assert bh_local.shape == (max_rows, G) and bh_local.dtype == torch.float32
    padded = torch.empty((world * max_rows, G), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(padded, bh_local)
    if all(int(r) == max_rows for r in rows_of_rank):
        return padded, padded
    parts = [padded[r * max_rows:r * max_rows + int(rows_of_rank[r])] for r in range(world)]
    return torch.cat(parts, 0), padded
