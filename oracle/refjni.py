"""TEST INFRASTRUCTURE — drives a PangeneNative-shaped JNI library through oracle/fakejni.cpp.

``RefJni(path)`` loads any shared object exporting the two JNI symbols of
``/root/reference/ig/native/pangene_native.h:16-25``: the unmodified reference build
(``oracle/_ref/libnative_ref.so``) or the B200 drop-in ``libnative.so``.  The reference keeps ONE process-global index
(``library.cpp:73``), so one RefJni instance per library per process.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .cport import Scores, ScoresStruct

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
FAKEJNI = os.path.join(REF_DIR, "libfakejni.so")
REFERENCE_LIB = os.path.join(REF_DIR, "libnative_ref.so")
REFERENCE_SRC = "/root/reference/ig/native/library.cpp"


def build_ref():
    """(Re)build oracle/_ref from the reference sources where they lie; no-op where /root/reference is absent."""
    if os.path.exists(REFERENCE_SRC):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)
    return os.path.exists(FAKEJNI) and os.path.exists(REFERENCE_LIB)


def available():
    return os.path.exists(FAKEJNI) and os.path.exists(REFERENCE_LIB)


_fj = None


def fj():
    global _fj
    if _fj is None:
        L = C.CDLL(FAKEJNI)
        L.fj_open.restype = C.c_void_p
        L.fj_open.argtypes = [C.c_char_p]
        L.fj_data_new.restype = C.c_void_p
        L.fj_data_new.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.fj_data_free.argtypes = [C.c_void_p]
        L.fj_preprocess.restype = C.c_double
        L.fj_preprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.fj_compute_scores.restype = C.POINTER(ScoresStruct)
        L.fj_compute_scores.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.fj_scores_free.argtypes = [C.POINTER(ScoresStruct)]
        L.fj_compute_scores_pool.restype = C.c_double
        L.fj_compute_scores_pool.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.fj_compute_scores_list.restype = C.c_double
        L.fj_compute_scores_list.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_void_p]
        _fj = L
    return _fj


class RefJni:
    def __init__(self, libpath=REFERENCE_LIB):
        self._lib = fj().fj_open(libpath.encode())
        if not self._lib:
            raise OSError("cannot load JNI library %s" % libpath)
        self._data = None
        self.preprocess_seconds = None

    def preprocess(self, residues, offsets, genome_of, k, only_complexity=False, quiet=True):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        genome_of = np.ascontiguousarray(genome_of, dtype=np.uint32)
        if self._data:
            fj().fj_data_free(self._data)
        self._data = fj().fj_data_new(residues.ctypes.data, offsets.ctypes.data, genome_of.ctypes.data, len(genome_of))
        self.preprocess_seconds = fj().fj_preprocess(self._lib, self._data, int(k), int(only_complexity), int(quiet))
        return self.preprocess_seconds

    def compute_scores(self, genome, quiet=True):
        p = fj().fj_compute_scores(self._lib, int(genome), int(quiet))
        try:
            return Scores.from_struct(p.contents)
        finally:
            fj().fj_scores_free(p)

    def compute_scores_pool(self, g_begin, g_end, threads, quiet=True):
        """Thread pool over genomes as Pangenes.java:54-66; returns (wall seconds, total cells)."""
        cells = C.c_int64(0)
        t = fj().fj_compute_scores_pool(self._lib, int(g_begin), int(g_end), int(threads), int(quiet), C.byref(cells))
        return t, cells.value

    def compute_scores_list(self, genomes, threads, quiet=True):
        """The same pool over an explicit list of genomes; returns (wall seconds, cells per genome as int64[])."""
        gl = np.ascontiguousarray(genomes, dtype=np.int32)
        per = np.zeros(len(gl), np.int64)
        cells = C.c_int64(0)
        t = fj().fj_compute_scores_list(self._lib, gl.ctypes.data, len(gl), int(threads), int(quiet), C.byref(cells), per.ctypes.data)
        return t, per
