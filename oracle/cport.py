"""TEST INFRASTRUCTURE — ctypes binding of oracle/pangenes_oracle.c (the CPU "port" of the reference hot path)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")


class ScoresStruct(C.Structure):
    """Flat copy of infoasys.cli.pangenes.Scores (reference Scores.java:3-35); shared by cport and refjni."""

    _fields_ = [
        ("scoresCount", C.c_int32), ("S", C.c_int32), ("rows", C.c_int32), ("G", C.c_int32),
        ("scores", C.POINTER(C.c_float)), ("percs", C.POINTER(C.c_float)), ("tr_percs", C.POINTER(C.c_float)),
        ("row", C.POINTER(C.c_int32)), ("column", C.POINTER(C.c_int32)),
        ("first_seq_genome", C.POINTER(C.c_int32)), ("second_seq_genome", C.POINTER(C.c_int32)),
        ("max_genome_score", C.POINTER(C.c_float)), ("max_genome_score_col", C.POINTER(C.c_float)),
        ("scoresMaxMappings", C.POINTER(C.c_int32)),
    ]


class Scores:
    """numpy view of one computeScores(genome) result; field names follow Scores.java."""

    FIELDS = ("scores", "percs", "tr_percs", "row", "column", "first_seq_genome", "second_seq_genome")

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @classmethod
    def from_struct(cls, st):
        n, S, rows, G = st.scoresCount, st.S, st.rows, st.G

        def arr(p, count, dt):
            if count == 0:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(p, shape=(count,)).astype(dt, copy=True)

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
        """Cells sorted by (row, column) — the reference's cell order is not part of the contract
        (all consumers in Pangenes.java:98-176 are max/min/set-insert)."""
        order = np.lexsort((self.column, self.row))
        return {f: getattr(self, f)[order] for f in self.FIELDS}


def build_lib():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "pangenes_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_lib())
        L.po_build.restype = C.c_void_p
        L.po_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32]
        L.po_free.argtypes = [C.c_void_p]
        for name, rt in (("po_genomes", C.c_uint32), ("po_alphabet_base", C.c_uint32), ("po_num_kmers", C.c_uint64),
                         ("po_num_entries", C.c_uint64), ("po_total_lookups", C.c_uint64)):
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [C.c_void_p]
        L.po_entries.argtypes = [C.c_void_p] * 4
        L.po_groups.argtypes = [C.c_void_p] * 3
        L.po_gene_stats.argtypes = [C.c_void_p] * 3
        L.po_compute_scores.restype = C.POINTER(ScoresStruct)
        L.po_compute_scores.argtypes = [C.c_void_p, C.c_uint32]
        L.po_scores_free.argtypes = [C.POINTER(ScoresStruct)]
        L.po_candidate_pairs.restype = C.c_uint64
        L.po_candidate_pairs.argtypes = [C.c_void_p, C.c_uint32]
        _lib = L
    return _lib


class OracleIndex:
    """po_build(...) handle.  `residues` uint8, `offsets` uint64[S+1], `genome_of` uint32[S]."""

    def __init__(self, residues, offsets, genome_of, k):
        self.residues = np.ascontiguousarray(residues, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.genome_of = np.ascontiguousarray(genome_of, dtype=np.uint32)
        self.S = len(self.genome_of)
        self.k = int(k)
        self._h = lib().po_build(self.residues.ctypes.data, self.offsets.ctypes.data, self.genome_of.ctypes.data, self.S, self.k)
        if not self._h:
            raise ValueError("oracle refused the input (k <= 0 or base^k overflow)")

    def close(self):
        if self._h:
            lib().po_free(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @property
    def genomes(self):
        return lib().po_genomes(self._h)

    @property
    def base(self):
        return lib().po_alphabet_base(self._h)

    @property
    def num_kmers(self):
        return lib().po_num_kmers(self._h)

    @property
    def num_entries(self):
        return lib().po_num_entries(self._h)

    @property
    def total_lookups(self):
        return lib().po_total_lookups(self._h)

    def entries(self):
        U = self.num_entries
        rank = np.zeros(U, np.uint64); seq = np.zeros(U, np.uint32); cnt = np.zeros(U, np.uint32)
        lib().po_entries(self._h, rank.ctypes.data, seq.ctypes.data, cnt.ctypes.data)
        return rank, seq, cnt

    def groups(self):
        U = self.num_entries
        gs = np.zeros(U, np.uint32); gl = np.zeros(U, np.uint32)
        lib().po_groups(self._h, gs.ctypes.data, gl.ctypes.data)
        return gs, gl

    def gene_stats(self):
        kl = np.zeros(self.S, np.uint32); tv = np.zeros(self.S, np.uint64)
        lib().po_gene_stats(self._h, kl.ctypes.data, tv.ctypes.data)
        return kl, tv

    def compute_scores(self, genome):
        p = lib().po_compute_scores(self._h, genome)
        try:
            return Scores.from_struct(p.contents)
        finally:
            lib().po_scores_free(p)

    def candidate_pairs(self, genome):
        return lib().po_candidate_pairs(self._h, genome)
