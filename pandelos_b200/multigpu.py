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


class ChunkedBestHitGather:
    """Scoring and the all-gather of the best-hit slices, overlapped chunk by chunk.

    Every rank's slice (padded to the same number of rows) is cut into `chunks` row chunks of `chunk_rows` rows; as soon
    as the engine has scored a chunk its rows are all-gathered asynchronously (NCCL on its own stream, NVLink) while the
    next chunk is being scored, so only the last chunk's exchange is exposed.  One all-gather after all scoring — the
    simple `allgather_best_hits` above — cost 27 ms on top of a 205 ms step on 8 B200s (8 x 1.9 GB slices).

    The gathered table is chunk-major: `table[c, r, i]` is local row `c * chunk_rows + i` of rank `r`; `offset_of(rank,
    local_row)` gives the row index into `table.view(-1, G)`, `assemble()` the plain rank-major (sum(rows) x G) table.
    """

    def __init__(self, dist, rows_of_rank, G, device, chunks):
        import torch
        self.dist, self.G, self.device = dist, int(G), device
        self.rows_of_rank = [int(r) for r in rows_of_rank]
        self.world = len(self.rows_of_rank)
        max_rows = max(self.rows_of_rank) if self.world else 0
        self.chunks = max(1, min(int(chunks), max(1, max_rows)))
        self.chunk_rows = max(1, -(-max_rows // self.chunks))
        self.table = torch.zeros((self.chunks, self.world, self.chunk_rows, self.G), dtype=torch.float32, device=device)
        # this rank's slice, padded to chunks x chunk_rows rows: the engine writes best hits straight into it
        self.local = torch.zeros((self.chunks * self.chunk_rows, self.G), dtype=torch.float32, device=device)
        self.works = []

    def local_range(self, rank, c):
        """Local rows [lo, hi) of `rank` inside chunk c (hi <= that rank's row count; may be empty)."""
        lo = min(c * self.chunk_rows, self.rows_of_rank[rank])
        hi = min((c + 1) * self.chunk_rows, self.rows_of_rank[rank])
        return lo, hi

    def gather_chunk(self, c):
        """Queues the all-gather of chunk c of every rank (asynchronous; call after the chunk's rows are final)."""
        src = self.local[c * self.chunk_rows:(c + 1) * self.chunk_rows]
        dst = self.table[c].view(self.world * self.chunk_rows, self.G)
        self.works.append(self.dist.all_gather_into_tensor(dst, src, async_op=True))

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []

    def offset_of(self, rank, local_row):
        c, i = divmod(int(local_row), self.chunk_rows)
        return (c * self.world + int(rank)) * self.chunk_rows + i

    def assemble(self):
        import torch
        parts = []
        for r in range(self.world):
            rows = self.table[:, r].reshape(self.chunks * self.chunk_rows, self.G)
            parts.append(rows[:self.rows_of_rank[r]])
        return torch.cat(parts, 0)


def score_and_gather(pn, gather, rank, row_begin):
    """Scores this rank's genes [row_begin, row_begin + rows) chunk by chunk with `pn.score_partition_device`, queueing
    each chunk's all-gather behind it; returns the summed pd_score_stats fields as a dict.  `gather.local` must be zero
    where rows are padding (it is after construction; the engine zeroes the rows it scores)."""
    try:   # gather.local was made (zero-filled) on torch's stream; the engine writes it from its own streams
        import torch
        if gather.local.is_cuda:
            torch.cuda.current_stream(gather.local.device).synchronize()
    except ImportError:
        pass
    total = None
    for c in range(gather.chunks):
        lo, hi = gather.local_range(rank, c)
        if hi > lo:
            st = pn.score_partition_device(row_begin + lo, row_begin + hi, best_hit_ptr=gather.local[lo].data_ptr())
            d = st.as_dict()
            total = d if total is None else {k: total[k] + v for k, v in d.items()}
        gather.gather_chunk(c)
    gather.wait()
    return total


class _DeviceArray:
    """A raw device pointer as something torch.as_tensor understands (__cuda_array_interface__)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


def _as_tensor(ptr, count, bits, device):
    """int32 / int64 tensor over `count` elements at `ptr` (device memory of `device`, or host memory for the CPU tests)."""
    import torch
    if device is not None and getattr(device, "type", "cpu") == "cuda":
        return torch.as_tensor(_DeviceArray(ptr, count, "<i4" if bits == 32 else "<i8"), device=device)
    import ctypes as C
    ct = C.c_int32 if bits == 32 else C.c_int64
    return torch.from_numpy(np.ctypeslib.as_array((ct * int(count)).from_address(int(ptr))))


_bulk_groups = {}


def _bulk_group(dist):
    """A second communicator for the one large transfer (the all-gather of the postings): collectives of one communicator
    run in order, so on the default group the small exchanges that follow would queue behind it."""
    key = (dist.get_backend(), dist.get_world_size())
    if key not in _bulk_groups:
        _bulk_groups[key] = dist.new_group(ranks=list(range(dist.get_world_size())))
    return _bulk_groups[key]


def build_sharded(dist, native, k, data, device=None, device_index=-1, residues_device_ptr=None, **engine_kw):
    """ONE index built by all ranks of `dist` together (include/pandelos_b200.h, pd_build_shard ...): every rank makes the
    k-mers of its share of the genes, an all-to-all moves every k-mer to the rank that owns its slice of the rank space,
    every rank sorts / dedups / groups its slice, the slices' postings / group-head bits / repeated-entry lists are
    all-gathered in place, the per-gene partial counts all-reduced, and every rank makes the forward lists of its own,
    genome-aligned, posting-list-volume-balanced share of the query rows.  Returns (PangeneNative, bounds[world + 1]):
    rank r serves genes [bounds[r], bounds[r + 1]).  The collectives of the multi-GPU path (NCCL over NVLink on GPUs;
    gloo in the CPU tests): 2 tiny all-gathers, 1 all-to-all (8 B per k-mer), 1 all-reduce (16 B per gene), 3 all-gathers
    (4 B per posting in all)."""
    import os
    import time
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    cpu = device is None or getattr(device, "type", "cpu") != "cuda"
    tdev = None if cpu else device
    bulk = _bulk_group(dist)   # (made once; every rank must make it at the same point)
    trace = os.environ.get("PD_SHARD_TRACE")   # host wall clock per stage, printed for calls slower than this many ms
    marks = [("start", time.perf_counter())]

    def mark(name):
        if trace:
            marks.append((name, time.perf_counter()))

    pn = native.PangeneNative(k, data, device=device_index, residues_device_ptr=residues_device_ptr, shard=(rank, world), **engine_kw)
    mark("build_shard")
    # ---- k-mers to the rank that sorts them
    send_counts = torch.tensor([int(pn.shard_keys.send_counts[r]) for r in range(world)], dtype=torch.int64, device=tdev)
    all_counts = torch.zeros(world * world, dtype=torch.int64, device=tdev)
    dist.all_gather_into_tensor(all_counts, send_counts)
    all_counts = all_counts.cpu().numpy().reshape(world, world)      # [source, destination]
    mark("counts")
    if int(all_counts.sum(axis=0).min()) < 2:   # known on every rank at the same point: all refuse together, nobody waits in a collective
        pn.close()
        raise native.PdError(native.PD_ERR_UNSUPPORTED, "sharded build: the input is too small for %d ranks (rank %d would get %d k-mers to sort)" % (
            world, int(all_counts.sum(axis=0).argmin()), int(all_counts.sum(axis=0).min())))
    in_splits = [int(v) for v in all_counts[:, rank]]
    out_splits = [int(v) for v in all_counts[rank, :]]
    n_recv, n_send = sum(in_splits), sum(out_splits)
    recv = _as_tensor(pn.shard_recv(n_recv), max(n_recv, 1), 64, device)[:n_recv]
    send = _as_tensor(pn.shard_keys.d_send, max(n_send, 1), 64, device)[:n_send]
    dist.all_to_all_single(recv, send, in_splits, out_splits)
    if not cpu:
        torch.cuda.current_stream().synchronize()   # the engine works on its own streams
    mark("all_to_all")
    si = pn.shard_sort()
    mark("shard_sort")
    # ---- the slices' results to every rank.  A slice has at most as many entries as it received k-mers, so the segment size
    # of the postings is known already: their all-gather (4 B per posting, the bulk) starts now and runs behind everything else
    seg_bound = int(all_counts.sum(axis=0).max())
    arr = pn.shard_buffers(seg_bound, 0)
    seg = int(arr.seg)
    post_all = _as_tensor(arr.d_post, seg * world, 32, device)
    mark("buffers")
    post_work = dist.all_gather_into_tensor(post_all, post_all[rank * seg:(rank + 1) * seg], group=bulk, async_op=True)   # in place: segment r is rank r's
    heads_all = _as_tensor(arr.d_heads, seg // 32 * world, 32, device)
    dist.all_gather_into_tensor(heads_all, heads_all[rank * (seg // 32):(rank + 1) * (seg // 32)])
    mine = torch.tensor([int(si.entries), int(si.multi)], dtype=torch.int64, device=tdev)
    counts = torch.zeros(2 * world, dtype=torch.int64, device=tdev)
    dist.all_gather_into_tensor(counts, mine)
    S = data.sequences_count
    dist.all_reduce(_as_tensor(si.d_gene_counts, 2 * S, 64, device))
    counts = counts.cpu().numpy().reshape(world, 2)
    if int(counts[world - 1, 0]) < 2:   # the reference's tail merge (library.cpp:300-306) joins the LAST entry to the group before it:
        post_work.wait()                # both must be in the last slice.  Known on every rank here: all refuse together.
        pn.close()
        raise native.PdError(native.PD_ERR_UNSUPPORTED, "sharded build: the last rank's slice holds fewer than two entries — use fewer ranks for this input")
    mseg = max(int(counts[:, 1].max()), 1)
    mark("small_collectives")
    multi_all = _as_tensor(pn.shard_multi(mseg), 2 * mseg * world, 32, device)
    dist.all_gather_into_tensor(multi_all, multi_all[rank * 2 * mseg:(rank + 1) * 2 * mseg])
    if not cpu:
        torch.cuda.current_stream().synchronize()   # the engine works on its own streams
    mark("multi")
    bounds = pn.shard_groups(counts[:, 0], counts[:, 1])   # while the postings are still travelling
    mark("groups")
    post_work.wait()
    if not cpu:
        torch.cuda.current_stream().synchronize()
    mark("post_wait")
    pn.shard_finish()
    mark("finish")
    if trace and (marks[-1][1] - marks[0][1]) * 1e3 > float(trace):
        import sys
        print("[shard trace] rank %d: " % rank + ", ".join("%s %.1f" % (marks[i][0], (marks[i][1] - marks[i - 1][1]) * 1e3) for i in range(1, len(marks))), file=sys.stderr, flush=True)
    return pn, bounds.astype(np.int64)
