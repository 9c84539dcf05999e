"""Synthetic `.faa`-shaped workloads of the BASELINE.json config shapes (ctypes binding of csrc/synth.cpp).

Measurement/test input only: family-structured proteins delivered in the packed form the C ABI
consumes — residue bytes, gene offsets, genome ids — i.e. what `PangeneIData.readFromFile`
(reference ig/infoasys/cli/pangenes/PangeneIData.java:30-75) holds after parsing.
"""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpdsynth.so")

# (genomes, genes per genome, mean length, substitution rate, seed) — BASELINE.json `configs`, SURVEY.md §8d
SHAPES = {
    "salmonella7": (7, 4500, 310.0, 0.05, 1),
    "ecoli10": (10, 5000, 315.0, 0.08, 2),
    "xanthomonas14": (14, 4500, 330.0, 0.10, 3),
    "mycoplasma64": (64, 800, 180.0, 0.25, 4),
    "scaleout1000": (1000, 4000, 300.0, 0.08, 5),
}


class _Params(C.Structure):
    _fields_ = [("genomes", C.c_uint32), ("genes_per_genome", C.c_uint32), ("mean_len", C.c_double), ("mu", C.c_double),
                ("seed", C.c_uint64), ("low_complexity", C.c_double), ("threads", C.c_uint32)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            from . import build
            build.build_synth()
        L = C.CDLL(_LIB)
        L.pd_synth_plan_new.restype = C.c_void_p
        L.pd_synth_plan_new.argtypes = [C.POINTER(_Params)]
        L.pd_synth_genes.restype = C.c_uint32
        L.pd_synth_genes.argtypes = [C.c_void_p]
        L.pd_synth_residues.restype = C.c_uint64
        L.pd_synth_residues.argtypes = [C.c_void_p]
        L.pd_synth_fill.argtypes = [C.c_void_p] * 5
        L.pd_synth_plan_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class Workload:
    """Packed protein set: residues uint8[total], offsets uint64[S+1], genome_of uint32[S], family_of uint32[S]."""

    def __init__(self, residues, offsets, genome_of, family_of=None, name=""):
        self.residues, self.offsets, self.genome_of, self.family_of, self.name = residues, offsets, genome_of, family_of, name

    @property
    def S(self):
        return len(self.genome_of)

    @property
    def G(self):
        return int(self.genome_of.max()) + 1 if len(self.genome_of) else 0

    def sequence(self, i):
        return self.residues[int(self.offsets[i]):int(self.offsets[i + 1])].tobytes().decode("latin-1")

    def subset_genomes(self, n):
        """First n genomes (genes of a genome are contiguous in generated workloads)."""
        keep = int(np.searchsorted(self.genome_of, n, side="left"))
        end = int(self.offsets[keep])
        return Workload(self.residues[:end], self.offsets[:keep + 1], self.genome_of[:keep],
                        None if self.family_of is None else self.family_of[:keep], "%s[:%d]" % (self.name, n))

    def write_faa(self, path):
        """Two lines per gene: `genome<TAB>gene-name<TAB>product`, then the sequence (PangeneIData.java:47-64)."""
        with open(path, "w") as f:
            per = {}
            for i in range(self.S):
                g = int(self.genome_of[i])
                n = per.get(g, 0)
                per[g] = n + 1
                fam = -1 if self.family_of is None else int(self.family_of[i])
                f.write("G%d\tg%d_%d@G%d:1\tfam%d\n%s\n" % (g, g, n, g, fam, self.sequence(i)))


def generate(genomes, genes_per_genome, mean_len, mu, seed, low_complexity=0.0, threads=0, name="synthetic"):
    L = _load()
    p = _Params(genomes, genes_per_genome, mean_len, mu, seed, low_complexity, threads)
    plan = L.pd_synth_plan_new(C.byref(p))
    try:
        S, total = L.pd_synth_genes(plan), L.pd_synth_residues(plan)
        residues = np.empty(total, np.uint8)
        offsets = np.empty(S + 1, np.uint64)
        genome_of = np.empty(S, np.uint32)
        family_of = np.empty(S, np.uint32)
        L.pd_synth_fill(plan, residues.ctypes.data, offsets.ctypes.data, genome_of.ctypes.data, family_of.ctypes.data)
    finally:
        L.pd_synth_plan_free(plan)
    return Workload(residues, offsets, genome_of, family_of, name)


def shape(name, scale=1.0, genomes=None, **kw):
    """One of SHAPES, optionally with fewer genomes or `scale` x genes per genome (for quick tests)."""
    G, M, ell, mu, seed = SHAPES[name]
    if genomes is not None:
        G = genomes
    return generate(G, max(1, int(M * scale)), ell, mu, seed, name=name, **kw)


def from_sequences(seqs, genome_ids, name="literal"):
    """Pack python strings (latin-1) into a Workload; for hand-written fixtures."""
    bs = [s.encode("latin-1") for s in seqs]
    offsets = np.zeros(len(bs) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    residues = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return Workload(residues, offsets, np.asarray(genome_ids, dtype=np.uint32), None, name)


def calculate_k(w):
    """k = floor(log_a(total residues) / H_a) with a = alphabet size (reference calculate_k.py:23-30).

    Same double-precision expression order as the reference: math.log(x, a) and a running sum over the
    alphabet in first-appearance order (dict insertion order)."""
    res = w.residues
    total = int(len(res))
    if total == 0:
        raise ValueError("empty input")
    counts = np.bincount(res, minlength=256)
    # alphabet in first-appearance order (the reference's dict insertion order); found from growing prefixes
    present = set(np.flatnonzero(counts).tolist())
    order, seen, pos, chunk = [], set(), 0, 1 << 16
    while len(seen) < len(present):
        part = res[pos:pos + chunk]
        vals, first = np.unique(part, return_index=True)
        for i in np.argsort(first):
            b = int(vals[i])
            if b not in seen:
                seen.add(b)
                order.append(b)
        pos += chunk
        chunk *= 4
    a = len(order)
    if a < 2:
        raise ValueError("alphabet of one letter: entropy is zero")
    ent = 0.0
    for b in order:
        c = int(counts[b])
        ent += -math.log(c / total, a) * (c / total)
    return math.floor(math.log(total, a) / ent)
