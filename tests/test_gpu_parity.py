"""GPU (-m gpu): the sm_100a engine against the oracle through the C ABI — bit-exact k-mer entries, groups, cost
model, cells (float32 bit patterns), best-hit tables — on fixtures, the reference's golden vectors, seeded
workloads at sizes the oracle finishes in seconds, and size-independent properties on a config-sized input."""
import numpy as np
import pytest

import fixtures
import golden_util
from parity import check_scores, check_workload
from pandelos_b200 import native, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def product_lib(engine_lib):
    native._lib = None
    native.load(native.ENGINE_LIB)
    yield


@pytest.mark.parametrize("name", sorted(fixtures.LITERAL))
def test_literal_fixtures(name):
    w, k = fixtures.literal(name)
    check_workload(w, k)


@pytest.mark.parametrize("name", golden_util.names())
def test_reference_golden_vectors(name):
    w, k, gold = golden_util.load(name)
    pn = native.PangeneNative(k, native.PangeneIData(w.residues, w.offsets, w.genome_of))
    try:
        assert pn.info.G == len(gold)
        for g, ref in enumerate(gold):
            check_scores(pn.generateScoresPart(g), ref, "%s genome %d" % (name, g))
    finally:
        pn.close()


@pytest.mark.parametrize("seed,k", [(201, 2), (202, 3), (203, 4), (204, 5)])
def test_random_dense_sharing(seed, k):
    w, _ = fixtures.random_workload(seed, genes=400, genomes=6, max_len=120, alphabet="ACDEFG", k=k)
    check_workload(w, k)


@pytest.mark.parametrize("shape,scale,k", [("salmonella7", 0.12, 4), ("mycoplasma64", 0.1, 4), ("ecoli10", 0.08, 5)])
def test_config_shapes_scaled(shape, scale, k):
    """The BASELINE.json shapes with fewer genes per genome: full index + every genome's cells vs the oracle."""
    w = synth.shape(shape, scale=scale)
    st = check_workload(w, k)
    assert st["cells"] > 0 and st["fallback_rows"] == 0


def test_large_index_level_config_on_a_small_workload(monkeypatch):
    """Indices below 2^18 genes take a larger top-level table; the configuration of the large ones (two 512-thread CTAs per
    SM, 4096 + 2048 slots) is forced here so that parity covers it at oracle sizes too."""
    monkeypatch.setenv("PD_LEVELS", "512:9:128:256:256,2048:11:256:512:1536,4096:11:512:1024:0,2048:14:1024:1024:0")
    w = synth.shape("ecoli10", scale=0.08)
    st = check_workload(w, 5, index=False)
    assert st["cells"] > 0


def test_low_complexity_multiplicities():
    """poly-A / poly-Q runs: counts > 1 everywhere, long posting lists, MULTI rows."""
    w = synth.generate(8, 300, 200.0, 0.1, 61, low_complexity=0.5)
    st = check_workload(w, 4)
    assert st["cells"] > 0


def test_huge_posting_lists():
    """A poly-A k-mer shared by more genes than kHugeList (2048): the CTA-wide walk of one list, plus rows whose
    distinct columns overflow the first table and are re-run on the retry level."""
    w = synth.generate(8, 900, 100.0, 0.1, 66, low_complexity=0.9)
    data = native.PangeneIData(w.residues, w.offsets, w.genome_of)
    pn = native.PangeneNative(4, data, keep_sorted=True)
    _, _, _, _, gl = pn.entries()
    pn.close()
    assert gl.max() > 2048
    st = check_workload(w, 4, genomes=[0, 3, 5], index=False)
    assert st["cells"] > 0


def test_long_genes_multi_segment_forward_lists():
    """Genes with more shared k-mers than one staged segment of forward entries (fcap = 1024 at the largest level)."""
    w = synth.generate(5, 60, 2500.0, 0.05, 67)
    st = check_workload(w, 5)
    assert st["cells"] > 0
    kl = np.diff(w.offsets.astype(np.int64))
    assert kl.max() > 3000


def test_overflow_levels_and_dense_path(monkeypatch):
    monkeypatch.setenv("PD_SMEM_TOP", "2048")
    w = synth.generate(6, 200, 150.0, 0.1, 62, low_complexity=0.3)
    st = check_workload(w, 4, hash_log2=6, index=False)
    assert st["fallback_rows"] > 0


def test_dense_fallback_kernels_shared_and_global(monkeypatch):
    """Rows wider than every hash table: small indices count them in shared memory (16-bit counter per gene), large ones
    in global memory; both kernels on the same rows, with repeated k-mers (the correction arrays) and without."""
    monkeypatch.setenv("PD_SMEM_TOP", "2048")
    for low in (0.0, 0.4):
        w = synth.generate(6, 220, 150.0, 0.1, 65, low_complexity=low)
        for kind in ("smem", "global"):
            monkeypatch.setenv("PD_DENSE", kind)
            st = check_workload(w, 4, hash_log2=6, index=False)
            assert st["fallback_rows"] > 0


def test_contexts_taken_over_by_indices_of_other_sizes(monkeypatch):
    """Score contexts (result buffers, dense accumulators, side tables) are parked when an index dies and taken over by
    the next one: a larger index, then a smaller one, then a larger one again, each through the dense fallback."""
    monkeypatch.setenv("PD_SMEM_TOP", "2048")
    for genes, seed in ((300, 71), (90, 72), (400, 73), (90, 74)):
        w = synth.generate(5, genes, 150.0, 0.1, seed, low_complexity=0.3)
        st = check_workload(w, 4, hash_log2=6, index=False)
        assert st["fallback_rows"] > 0


def _scores_digest(s):
    """Order-independent digest of every field of a Scores object (cell order is free)."""
    n = s.scoresCount
    f32 = lambda a: np.ascontiguousarray(a[:n], dtype=np.float32).view(np.uint32).astype(np.uint64)
    i32 = lambda a: np.ascontiguousarray(a[:n]).astype(np.int64).astype(np.uint64)
    h = (i32(s.row) * np.uint64(0x9E3779B97F4A7C15)) ^ (i32(s.column) * np.uint64(0xC2B2AE3D27D4EB4F))
    h = h * np.uint64(0x165667B19E3779F9) + f32(s.scores) + (f32(s.percs) << np.uint64(7)) + (f32(s.tr_percs) << np.uint64(13))
    h = h ^ (i32(s.first_seq_genome) << np.uint64(40)) ^ (i32(s.second_seq_genome) << np.uint64(52))
    return (int(n), int(h.sum(dtype=np.uint64)), np.ascontiguousarray(s.max_genome_score).tobytes(),
            np.ascontiguousarray(s.max_genome_score_col).tobytes(), np.ascontiguousarray(s.scoresMaxMappings).tobytes())


def test_concurrent_calls_stress():
    """Four host threads, every genome several times each, results landing in pinned memory that the host learns about by
    polling a completion number (no stream synchronisation): every call must return exactly what a lone call returns."""
    import threading
    w = synth.shape("salmonella7", scale=0.5)
    pn = native.PangeneNative(5, native.PangeneIData(w.residues, w.offsets, w.genome_of), contexts=4)
    try:
        G = pn.info.G
        want = [_scores_digest(pn.generateScoresPart(g)) for g in range(G)]
        assert sum(d[0] for d in want) > 0
        bad, errs = [], []

        def work(t):
            try:
                for rep in range(6):
                    for g in range(G):
                        gg = (g + t) % G
                        if _scores_digest(pn.generateScoresPart(gg)) != want[gg]:
                            bad.append((t, rep, gg))
            except Exception as e:  # pragma: no cover
                errs.append(e)

        th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        assert not bad, bad[:5]
    finally:
        pn.close()


def test_cell_buffer_regrow_and_concurrent_calls():
    import threading
    w = synth.generate(6, 200, 150.0, 0.1, 64)
    data = native.PangeneIData(w.residues, w.offsets, w.genome_of)
    pn = native.PangeneNative(4, data, cell_capacity=100, contexts=3)
    from oracle import cport
    o = cport.OracleIndex(w.residues, w.offsets, w.genome_of, 4)
    want = [o.compute_scores(g) for g in range(pn.info.G)]
    got = [None] * pn.info.G
    errs = []

    def work(g):
        try:
            got[g] = pn.generateScoresPart(g)   # Pangenes.java:60-66: one pool task per genome, concurrent
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(g,)) for g in range(pn.info.G)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    for g in range(pn.info.G):
        check_scores(got[g], want[g], "genome %d" % g)
    pn.close()


def test_full_size_properties():
    """E. coli-shaped config at full size (10 x ~5,000 proteins, k from calculate_k): properties that need no oracle —
    every cell has its mirror with identical score bits and swapped percs; best-hit tables equal the maxima of the
    cells; colmax_g[c] == BH[c][g]; the device partition path reproduces pairs/cells; one sampled genome vs the oracle."""
    w = synth.shape("ecoli10")
    k = synth.calculate_k(w)
    data = native.PangeneIData(w.residues, w.offsets, w.genome_of)
    pn = native.PangeneNative(k, data)
    S, G = pn.info.S, pn.info.G
    BH = np.zeros((S, G), np.float32)
    keys, vals = [], []
    pairs = cells = 0
    colmax = []
    for g in range(G):
        s = pn.generateScoresPart(g)
        pairs += pn.last_stats.pairs
        cells += s.scoresCount
        rows = np.nonzero(w.genome_of == g)[0]
        BH[rows] = s.max_genome_score
        colmax.append(s.max_genome_score_col)
        assert (s.first_seq_genome == g).all() and (s.second_seq_genome == w.genome_of[s.column].astype(np.int32)).all()
        assert (s.row != s.column).all() and (s.scores > 0).all() and (s.scores <= 1).all()
        # best hits are the maxima of the emitted cells
        m = np.zeros((S, G), np.float32)
        np.maximum.at(m, (s.row, s.second_seq_genome), s.scores)
        assert (m[rows] == s.max_genome_score).all()
        keys.append(s.row.astype(np.int64) * S + s.column)
        vals.append(np.stack([s.scores.view(np.uint32), s.percs.view(np.uint32), s.tr_percs.view(np.uint32)], 1))
    keys = np.concatenate(keys)
    vals = np.concatenate(vals)
    order = np.argsort(keys)
    keys, vals = keys[order], vals[order]
    assert len(np.unique(keys)) == len(keys)
    mirror = (keys % S) * S + keys // S
    pos = np.searchsorted(keys, mirror)
    assert (keys[pos] == mirror).all()
    assert (vals[pos][:, 0] == vals[:, 0]).all() and (vals[pos][:, 1] == vals[:, 2]).all() and (vals[pos][:, 2] == vals[:, 1]).all()
    for g in range(G):
        assert (colmax[g] == BH[:, g]).all()
    st = pn.score_partition_device(0, S)
    assert st.pairs == pairs and st.cells == cells and st.lookups == pn.info.lookups
    from oracle import cport
    o = cport.OracleIndex(w.residues, w.offsets, w.genome_of, k)
    assert o.total_lookups == pn.info.lookups and o.num_entries == pn.info.U
    check_scores(pn.generateScoresPart(3), o.compute_scores(3), "ecoli10 genome 3")
    pn.close()


def test_jni_dropin_against_unmodified_reference_library():
    """libnative.so (JNI shim over the engine) and the UNMODIFIED reference library, both driven through the same
    fake JNIEnv (oracle/fakejni.cpp): every Scores field the Java side reads must be identical."""
    import os
    from oracle import refjni
    from pandelos_b200 import build
    if not refjni.available() or not os.path.exists(build.JNI_LIB):
        pytest.skip("oracle/_ref or libnative.so not built")
    w = synth.generate(5, 120, 150.0, 0.1, 71, low_complexity=0.2)
    k = 4
    ref = refjni.RefJni()
    ref.preprocess(w.residues, w.offsets, w.genome_of, k)
    mine = refjni.RefJni(build.JNI_LIB)
    mine.preprocess(w.residues, w.offsets, w.genome_of, k)
    cells = 0
    for g in range(w.G):
        a, b = mine.compute_scores(g), ref.compute_scores(g)
        check_scores(a, b, "jni genome %d" % g)
        cells += a.scoresCount
    assert cells > 0


def test_several_devices_serve_the_same_results():
    """pd_options.devices = 2: one replica per device, genomes dealt out by posting-list volume; every per-genome result
    must be what a single-device index returns, and the partition call must see the same job."""
    import ctypes as C
    if native.load().pd_device_count() < 2:
        pytest.skip("one device visible")
    w = synth.shape("mycoplasma64", scale=0.3)
    k = 4
    data = native.PangeneIData(w.residues, w.offsets, w.genome_of)
    one = native.PangeneNative(k, data)
    two = native.PangeneNative(k, data, devices=2)
    try:
        assert two.devices() == 2 and one.devices() == 1
        owners = [two.genome_device(g) for g in range(one.info.G)]
        assert set(owners) == {0, 1} and owners == sorted(owners)
        for g in range(one.info.G):
            assert _scores_digest(two.generateScoresPart(g)) == _scores_digest(one.generateScoresPart(g)), g
            a, b = two.genomeEdges(g), one.genomeEdges(g)
            assert sorted(zip(a[0].tolist(), a[1].tolist(), a[2].view(np.uint32).tolist())) == sorted(zip(b[0].tolist(), b[1].tolist(), b[2].view(np.uint32).tolist()))
        s2, s1 = two.score_partition_device(0, one.info.S), one.score_partition_device(0, one.info.S)
        assert (s2.pairs, s2.cells, s2.lookups, s2.rows) == (s1.pairs, s1.cells, s1.lookups, s1.rows)
    finally:
        one.close()
        two.close()
    native.load().pd_trim()


def test_release_after_free_and_trim():
    """A Scores object released after its index was freed (a host that tears down in the wrong order) and pd_trim between
    indices: no crash, the next index works and returns the same result."""
    import ctypes as C
    w = synth.generate(4, 80, 90.0, 0.1, 81)
    data = native.PangeneIData(w.residues, w.offsets, w.genome_of)
    L = native.load()
    pn = native.PangeneNative(4, data)
    want = _scores_digest(pn.generateScoresPart(1))
    st, rel = pn.compute_scores_raw(2)
    pn.close()
    rel()
    L.pd_trim()
    pn = native.PangeneNative(4, data)
    assert _scores_digest(pn.generateScoresPart(1)) == want
    pn.close()
