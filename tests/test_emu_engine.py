"""CPU: kernel LOGIC of the engine against the oracle, through the same C ABI and ctypes binding the GPU tests use.

The .cu sources are compiled by g++ against tests/emu/cuemu.h (a fiber-based SIMT emulator: real barriers, warp
shuffles/ballots/match, atomics) — test infrastructure only; it proves nothing about performance and is never a
fallback of the product library.  Inputs are tiny (the emulator runs one thread at a time)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

import build_emu  # noqa: E402
import fixtures  # noqa: E402
import golden_util  # noqa: E402
from parity import check_scores, check_workload  # noqa: E402
from pandelos_b200 import native, synth  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def emu_lib():
    path = build_emu.build()
    native.load(path)
    yield
    native._lib = None  # later modules load the product library again


@pytest.mark.parametrize("name", sorted(fixtures.LITERAL))
def test_literal_fixtures(name):
    w, k = fixtures.literal(name)
    check_workload(w, k)


@pytest.mark.parametrize("name", golden_util.names())
def test_golden_vectors(name):
    """Engine logic against the outputs of the unmodified reference."""
    w, k, gold = golden_util.load(name)
    pn = native.PangeneNative(k, native.PangeneIData(w.residues, w.offsets, w.genome_of))
    try:
        assert pn.info.G == len(gold)
        for g, ref in enumerate(gold):
            check_scores(pn.generateScoresPart(g), ref, "%s genome %d" % (name, g))
    finally:
        pn.close()


def test_family_workload_multi_tile_sort():
    # > 8192 k-mers: several radix tiles, several scan tiles
    w = synth.generate(4, 80, 90.0, 0.1, 41)
    st = check_workload(w, 3)
    assert st["N"] > 3 * 8192 and st["cells"] > 0


def test_overflow_levels_and_dense_path(monkeypatch):
    """Tiny tables: rows climb level 1 -> level 2 -> dense global accumulators and still match bit for bit."""
    monkeypatch.setenv("PD_SMEM_TOP", "512")
    w, k = fixtures.random_workload(51, genes=70, genomes=4, max_len=45)
    st = check_workload(w, k, hash_log2=4)
    assert st["fallback_rows"] > 0
    w = synth.generate(3, 60, 80.0, 0.1, 52, low_complexity=0.6)
    st = check_workload(w, 3, hash_log2=5)
    assert st["fallback_rows"] > 0
    # the same rows through the global-memory fallback kernel (what indices above 112 K genes use)
    monkeypatch.setenv("PD_DENSE", "global")
    st = check_workload(w, 3, hash_log2=5)
    assert st["fallback_rows"] > 0
    monkeypatch.delenv("PD_DENSE")
    # a smaller index takes over the parked score context of the larger one (dense accumulators laid out for another S)
    w, k = fixtures.random_workload(51, genes=70, genomes=4, max_len=45)
    st = check_workload(w, k, hash_log2=4)
    assert st["fallback_rows"] > 0


def test_table_sizes_that_are_not_powers_of_two(monkeypatch):
    """The level tables' tier-1 size is any multiple of 32 (the default top level has 4992 slots): the slot function
    floor(c * T1 / S), the clearing and the finalize scan on small odd sizes, with and without overflow into the retry level."""
    w = synth.generate(4, 70, 85.0, 0.1, 53)
    for levels in ("96:5:128:64:40,160:6:128:96:200,224:6:128:96:0,352:8:128:128:0",
                   "32:5:128:32:8,96:5:128:64:64,160:5:128:64:0,480:7:256:128:0"):
        monkeypatch.setenv("PD_LEVELS", levels)
        st = check_workload(w, 3)
        assert st["cells"] > 0
    monkeypatch.delenv("PD_LEVELS")


def test_cell_buffer_regrow():
    w, k = fixtures.random_workload(54, genes=60, genomes=3, max_len=40)
    st = check_workload(w, k, cell_capacity=7)
    assert st["cells"] > 7


def test_partition_rows_and_device_scoring():
    w = synth.generate(4, 50, 80.0, 0.1, 55)
    pn = native.PangeneNative(3, native.PangeneIData(w.residues, w.offsets, w.genome_of))
    try:
        S = pn.info.S
        for parts in (1, 2, 3, 8):
            b = pn.partition_rows(parts)
            assert b[0] == 0 and b[-1] == S and (b[1:] >= b[:-1]).all()
        b = pn.partition_rows(2, snap_to_genomes=True)
        assert b[1] in set([0, S] + [int((w.genome_of < g).sum()) for g in range(w.G + 1)])
        total = 0
        cells = 0
        for g in range(pn.info.G):
            cells += pn.generateScoresPart(g).scoresCount
            total += pn.last_stats.pairs
        st = pn.score_partition_device(0, S, rows_per_launch=37)
        assert st.pairs == total and st.cells == cells and st.rows == S and st.lookups == pn.info.lookups
    finally:
        pn.close()


def test_errors():
    w, _ = fixtures.literal("identical_pair")
    data = native.PangeneIData(w.residues, w.offsets, w.genome_of)
    with pytest.raises(native.PdError) as e:
        native.PangeneNative(0, data)
    assert e.value.code == native.PD_ERR_INVALID
    with pytest.raises(native.PdError) as e:
        native.PangeneNative(20, data)
    assert e.value.code == native.PD_ERR_UNSUPPORTED
    pn = native.PangeneNative(3, data)
    with pytest.raises(native.PdError):
        pn.generateScoresPart(7)
    pn.close()
