"""ctypes binding of the C ABI in include/pandelos_b200.h (libpandelos_b200.so: sm_100a kernels, no CPU path).

The class and method names mirror the reference's Java-side operator for this path:

* ``PangeneNative(k, data)``             reference ig/infoasys/cli/pangenes/PangeneNative.java:5-7  (preprocessSequences)
* ``PangeneNative.generateScoresPart(g)`` reference PangeneNative.java:17-21                          (computeScores)
* ``Scores``                              reference ig/infoasys/cli/pangenes/Scores.java:3-35

so host code and tests written against them read like ``Pangenes.main`` (Pangenes.java:39,66).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ENGINE_LIB = os.path.join(_HERE, "libpandelos_b200.so")

PD_OK, PD_ERR_INVALID, PD_ERR_UNSUPPORTED, PD_ERR_CUDA, PD_ERR_NO_DEVICE, PD_ERR_NOMEM = 0, -1, -2, -3, -4, -5


class PdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pandelos_b200 error %d: %s" % (code, msg))
        self.code = code


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("verbose", C.c_int32), ("contexts", C.c_int32), ("hash_log2", C.c_int32),
                ("cell_capacity", C.c_uint64), ("keep_sorted", C.c_int32), ("devices", C.c_int32),
                ("table_on_device", C.c_int32), ("reserved", C.c_int32)]


class IndexInfo(C.Structure):
    _fields_ = [("S", C.c_uint32), ("G", C.c_uint32), ("k", C.c_int32), ("base", C.c_uint32), ("rank_bits", C.c_uint32),
                ("seq_bits", C.c_uint32), ("N", C.c_uint64), ("U", C.c_uint64), ("groups", C.c_uint64), ("R", C.c_uint64),
                ("lookups", C.c_uint64), ("max_kseq", C.c_uint32), ("reserved", C.c_uint32), ("build_ms", C.c_double * 8)]


class ScoresStruct(C.Structure):
    _fields_ = [("scoresCount", C.c_int32), ("S", C.c_int32), ("rows", C.c_int32), ("G", C.c_int32),
                ("scores", C.POINTER(C.c_float)), ("percs", C.POINTER(C.c_float)), ("tr_percs", C.POINTER(C.c_float)),
                ("row", C.POINTER(C.c_int32)), ("column", C.POINTER(C.c_int32)),
                ("first_seq_genome", C.POINTER(C.c_int32)), ("second_seq_genome", C.POINTER(C.c_int32)),
                ("max_genome_score", C.POINTER(C.c_float)), ("max_genome_score_col", C.POINTER(C.c_float)),
                ("scoresMaxMappings", C.POINTER(C.c_int32)), ("owner", C.c_void_p)]


class ScoreStats(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("lookups", C.c_uint64), ("pairs", C.c_uint64), ("cells", C.c_uint64),
                ("fallback_rows", C.c_uint64), ("launches", C.c_uint64), ("fwd_entries", C.c_uint64), ("retry_rows", C.c_uint64), ("kernel_ms", C.c_double), ("total_ms", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class EdgesStruct(C.Structure):
    _fields_ = [("count", C.c_uint64), ("src", C.POINTER(C.c_uint32)), ("dst", C.POINTER(C.c_uint32)), ("score", C.POINTER(C.c_float)),
                ("cells", C.c_uint64), ("owner", C.c_void_p)]


class ShardKeys(C.Structure):
    _fields_ = [("d_send", C.c_void_p), ("send_counts", C.c_uint64 * 32)]


class ShardInfo(C.Structure):
    _fields_ = [("entries", C.c_uint64), ("multi", C.c_uint64), ("kmers", C.c_uint64), ("d_gene_counts", C.c_void_p)]


class ShardArrays(C.Structure):
    _fields_ = [("d_post", C.c_void_p), ("d_heads", C.c_void_p), ("d_multi", C.c_void_p), ("seg", C.c_uint64), ("mseg", C.c_uint64)]


_lib = None
_lib_path = None


def load(path=None):
    """Loads the engine library.  There is no fallback: a missing library is an error."""
    global _lib, _lib_path
    if path is None and _lib is not None:
        return _lib
    path = path or os.environ.get("PANDELOS_B200_LIB") or ENGINE_LIB
    if _lib is not None and _lib_path == path:
        return _lib
    if not os.path.exists(path):
        raise OSError("%s not built: run `python -m pandelos_b200.build engine` (needs nvcc)" % path)
    L = C.CDLL(path)
    L.pd_last_error.restype = C.c_char_p
    L.pd_device_count.restype = C.c_int
    for name in ("pd_build", "pd_build_device"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.POINTER(Options), C.POINTER(C.c_void_p)]
    L.pd_free.argtypes = [C.c_void_p]
    L.pd_trim.argtypes = []
    L.pd_info.argtypes = [C.c_void_p, C.POINTER(IndexInfo)]
    L.pd_gene_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.pd_entries.argtypes = [C.c_void_p] * 6
    L.pd_compute_scores.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(ScoresStruct)]
    L.pd_scores_release.argtypes = [C.c_void_p, C.POINTER(ScoresStruct)]
    L.pd_last_score_stats.argtypes = [C.POINTER(ScoresStruct), C.POINTER(ScoreStats)]
    L.pd_genome_edges.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(EdgesStruct)]
    L.pd_edges_release.argtypes = [C.c_void_p, C.POINTER(EdgesStruct)]
    L.pd_score_partition_device.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(ScoreStats)]
    L.pd_partition_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p]
    L.pd_devices.argtypes = [C.c_void_p]
    L.pd_genome_device.argtypes = [C.c_void_p, C.c_uint32]
    L.pd_build_shard.restype = C.c_int
    L.pd_build_shard.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.POINTER(Options), C.c_uint32, C.c_uint32,
                                 C.POINTER(C.c_void_p), C.POINTER(ShardKeys)]
    L.pd_shard_recv.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    L.pd_shard_sort.argtypes = [C.c_void_p, C.POINTER(ShardInfo)]
    L.pd_shard_buffers.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(ShardArrays)]
    L.pd_shard_multi.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    L.pd_shard_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pd_shard_finish.argtypes = [C.c_void_p]
    _lib, _lib_path = L, path
    return L


def _check(rc):
    if rc != PD_OK:
        raise PdError(rc, (load().pd_last_error() or b"").decode("utf-8", "replace"))


class Scores:
    """One computeScores(genome) result; field names as in Scores.java:3-35 (arrays are numpy copies)."""

    FIELDS = ("scores", "percs", "tr_percs", "row", "column", "first_seq_genome", "second_seq_genome")

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @classmethod
    def from_struct(cls, st, copy=True):
        n, S, rows, G = st.scoresCount, st.S, st.rows, st.G

        def arr(p, count, dt):
            if count == 0:
                return np.zeros(0, dtype=dt)
            a = np.ctypeslib.as_array(p, shape=(count,))
            return a.astype(dt, copy=True) if copy else a

        return cls(
            scoresCount=n,
            scores=arr(st.scores, n, np.float32), percs=arr(st.percs, n, np.float32), tr_percs=arr(st.tr_percs, n, np.float32),
            row=arr(st.row, n, np.int32), column=arr(st.column, n, np.int32),
            first_seq_genome=arr(st.first_seq_genome, n, np.int32), second_seq_genome=arr(st.second_seq_genome, n, np.int32),
            max_genome_score=arr(st.max_genome_score, rows * G, np.float32).reshape(rows, G),
            max_genome_score_col=arr(st.max_genome_score_col, S, np.float32),
            scoresMaxMappings=arr(st.scoresMaxMappings, S, np.int32),
        )

    def canonical(self):
        """Cells sorted by (row, column): the cell order is not part of the contract (Pangenes.java:98-176)."""
        order = np.lexsort((self.column, self.row))
        return {f: getattr(self, f)[order] for f in self.FIELDS}


class PangeneIData:
    """What PangeneIData.readFromFile holds after parsing (PangeneIData.java:30-75), in packed form:
    residues uint8[total], offsets uint64[S+1], sequenceGenome uint32[S]; optional names."""

    def __init__(self, residues, offsets, sequenceGenome, sequenceName=None, genomeNames=None):
        self.residues = np.ascontiguousarray(residues, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.sequenceGenome = np.ascontiguousarray(sequenceGenome, dtype=np.uint32)
        self.sequenceName = sequenceName
        self.genomeNames = genomeNames

    @property
    def sequences_count(self):
        return len(self.sequenceGenome)


class PangeneNative:
    """Device-resident index + scoring calls: `new PangeneNative(k, pid)` / `generateScoresPart(g)`."""

    def __init__(self, k, data, device=-1, verbose=False, contexts=0, hash_log2=0, cell_capacity=0, keep_sorted=False,
                 residues_device_ptr=None, shard=None, devices=0, table_device_ptrs=None):
        """shard = (rank, world): this process builds slice `rank` of an index that `world` processes build together
        (pd_build_shard); the caller exchanges the slices and calls shard_buffers / shard_finish
        (pandelos_b200.multigpu.build_sharded does all of it over torch.distributed)."""
        L = load()
        self._L = L
        self._data = data
        self.k = int(k)
        opt = Options(int(device), int(verbose), int(contexts), int(hash_log2), int(cell_capacity), int(keep_sorted), int(devices),
                      1 if (table_device_ptrs is not None and residues_device_ptr is not None) else 0, 0)
        # the gene table: host arrays, or (table_device_ptrs = (offsets, genome ids), with device residues) resident in HBM too
        if opt.table_on_device:
            off_arg, gid_arg = C.c_void_p(int(table_device_ptrs[0])), C.c_void_p(int(table_device_ptrs[1]))
        else:
            off_arg, gid_arg = data.offsets.ctypes.data, data.sequenceGenome.ctypes.data
        h = C.c_void_p()
        S = data.sequences_count
        self.shard_keys = None
        if shard is not None:
            rank, world = shard
            self.shard_keys = ShardKeys()
            on_dev = residues_device_ptr is not None
            rc = L.pd_build_shard(C.c_void_p(int(residues_device_ptr)) if on_dev else data.residues.ctypes.data, int(on_dev),
                                  off_arg, gid_arg, S, self.k, C.byref(opt), int(rank), int(world),
                                  C.byref(h), C.byref(self.shard_keys))
        elif residues_device_ptr is not None:
            rc = L.pd_build_device(C.c_void_p(int(residues_device_ptr)), off_arg, gid_arg, S, self.k, C.byref(opt), C.byref(h))
        else:
            rc = L.pd_build(data.residues.ctypes.data, data.offsets.ctypes.data, data.sequenceGenome.ctypes.data, S, self.k,
                            C.byref(opt), C.byref(h))
        self._h = None
        _check(rc)
        self._h = h
        self.info = IndexInfo()
        _check(L.pd_info(self._h, C.byref(self.info)))
        self.last_stats = None

    @staticmethod
    def printComplexity(k, data):
        """PangeneNative.printComplexity (PangeneNative.java:10-12): the cost report only."""
        PangeneNative(k, data, verbose=True).close()

    def close(self):
        if getattr(self, "_h", None):
            self._L.pd_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- computeScores
    def generateScoresPart(self, genome, multithread=True):
        st = ScoresStruct()
        _check(self._L.pd_compute_scores(self._h, int(genome), C.byref(st)))
        try:
            stats = ScoreStats()
            self._L.pd_last_score_stats(C.byref(st), C.byref(stats))
            self.last_stats = stats
            return Scores.from_struct(st)
        finally:
            self._L.pd_scores_release(self._h, C.byref(st))

    def compute_scores_raw(self, genome):
        """Zero-copy variant: returns (ScoresStruct, release callable); arrays live in the library's pinned buffers."""
        st = ScoresStruct()
        _check(self._L.pd_compute_scores(self._h, int(genome), C.byref(st)))
        return st, (lambda: self._L.pd_scores_release(self._h, C.byref(st)))

    def genomeEdges(self, genome):
        """The pool task of Pangenes.java:60-183 for one genome, filter included, on the device: returns
        (src uint32[], dst uint32[], score float32[], cells) = the addConnection calls (inter-genome BBH edges once,
        as (row gene, column gene); intra-genome edges with src < dst)."""
        e = EdgesStruct()
        _check(self._L.pd_genome_edges(self._h, int(genome), C.byref(e)))
        try:
            n = int(e.count)

            def arr(p, dt):
                return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)

            return arr(e.src, np.uint32), arr(e.dst, np.uint32), arr(e.score, np.float32), int(e.cells)
        finally:
            self._L.pd_edges_release(self._h, C.byref(e))

    def genome_edges_raw(self, genome):
        """Zero-copy variant of genomeEdges: returns (EdgesStruct, release callable)."""
        e = EdgesStruct()
        _check(self._L.pd_genome_edges(self._h, int(genome), C.byref(e)))
        return e, (lambda: self._L.pd_edges_release(self._h, C.byref(e)))

    # ---- sharded build, steps 2 and 3 (include/pandelos_b200.h)
    def shard_recv(self, n_recv):
        p = C.c_void_p()
        _check(self._L.pd_shard_recv(self._h, int(n_recv), C.byref(p)))
        return p.value

    def shard_sort(self):
        si = ShardInfo()
        _check(self._L.pd_shard_sort(self._h, C.byref(si)))
        return si

    def shard_buffers(self, max_entries, max_multi):
        a = ShardArrays()
        _check(self._L.pd_shard_buffers(self._h, int(max_entries), int(max_multi), C.byref(a)))
        return a

    def shard_multi(self, max_multi):
        p = C.c_void_p()
        _check(self._L.pd_shard_multi(self._h, int(max_multi), C.byref(p)))
        return p.value

    def shard_groups(self, entries_of_rank, multi_of_rank):
        e = np.ascontiguousarray(entries_of_rank, dtype=np.uint64)
        m = np.ascontiguousarray(multi_of_rank, dtype=np.uint64)
        b = np.zeros(len(e) + 1, np.uint32)
        _check(self._L.pd_shard_groups(self._h, e.ctypes.data, m.ctypes.data, b.ctypes.data))
        return b

    def shard_finish(self):
        _check(self._L.pd_shard_finish(self._h))
        _check(self._L.pd_info(self._h, C.byref(self.info)))

    def devices(self):
        return int(self._L.pd_devices(self._h))

    def genome_device(self, genome):
        return int(self._L.pd_genome_device(self._h, int(genome)))

    # ---- diagnostics / partitions
    def gene_stats(self):
        S = self.info.S
        kl = np.zeros(S, np.uint32)
        tv = np.zeros(S, np.uint64)
        _check(self._L.pd_gene_stats(self._h, kl.ctypes.data, tv.ctypes.data))
        return kl, tv

    def entries(self, with_groups=True):
        U = self.info.U
        rank = np.zeros(U, np.uint64); seq = np.zeros(U, np.uint32); cnt = np.zeros(U, np.uint32)
        gs = np.zeros(U, np.uint32); gl = np.zeros(U, np.uint32)
        if with_groups:
            _check(self._L.pd_entries(self._h, rank.ctypes.data, seq.ctypes.data, cnt.ctypes.data, gs.ctypes.data, gl.ctypes.data))
        else:
            _check(self._L.pd_entries(self._h, None, seq.ctypes.data, cnt.ctypes.data, None, None))
        return rank, seq, cnt, gs, gl

    def partition_rows(self, parts, snap_to_genomes=False):
        b = np.zeros(parts + 1, np.uint32)
        _check(self._L.pd_partition_rows(self._h, parts, int(snap_to_genomes), b.ctypes.data))
        return b

    def score_partition_device(self, row_begin, row_end, best_hit_ptr=None, rows_per_launch=0):
        """Device-resident scoring of genes [row_begin, row_end); best_hit_ptr = device pointer to (rows x G) floats or None."""
        stats = ScoreStats()
        _check(self._L.pd_score_partition_device(self._h, int(row_begin), int(row_end), int(rows_per_launch),
                                                 C.c_void_p(int(best_hit_ptr)) if best_hit_ptr else None, C.byref(stats)))
        self.last_stats = stats
        return stats
