"""The Java host's network construction (Pangenes.java:98-176, PangeneNet.java) — restated in oracle/pangenes_java.py —
against the device filter of pd_genome_edges.  CPU tests run the kernels' logic through the SIMT emulator build."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

import fixtures  # noqa: E402
from oracle import cport, pangenes_java  # noqa: E402
from pandelos_b200 import native, synth  # noqa: E402


def engine_net(pn, genome_of):
    net = pangenes_java.PangeneNet()
    for g in range(pn.info.G):
        src, dst, sc, _ = pn.genomeEdges(g)
        for a, b, s in zip(src.tolist(), dst.tolist(), sc):
            net.add_connection(a, b, s)
            if genome_of[a] != genome_of[b]:
                net.add_connection(b, a, s)     # Pangenes.java:103-104
    return net


def oracle_net(w, k):
    o = cport.OracleIndex(w.residues, w.offsets, w.genome_of, k)
    try:
        return pangenes_java.run(o.compute_scores, o.genomes)
    finally:
        o.close()


def check(w, k):
    pn = native.PangeneNative(k, native.PangeneIData(w.residues, w.offsets, w.genome_of))
    try:
        got = engine_net(pn, w.genome_of)
    finally:
        pn.close()
    want = oracle_net(w, k)
    assert got.edge_set() == want.edge_set()
    assert got.lines() == want.lines()
    return len(want.lines())


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    native.load(build_emu.build())
    yield
    native._lib = None


def test_native_cli_fast_number_formatter_matches_the_java_rule():
    """The CLI's .net writer formats scores with to_chars; it must print what Double.toString prints (CPU only)."""
    import subprocess
    from pandelos_b200 import build
    build.build_host()
    r = subprocess.run([build.CLI_BIN, "--selftest-format", "300000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


def test_java_double_to_string_known_values():
    j = pangenes_java.java_double_to_string
    assert j(np.float32(1 / 3)) == "0.3333333432674408" and j(np.float32(0.5)) == "0.5" and j(np.float32(1.0)) == "1.0"
    assert j(np.float32(1e-4)) == "9.999999747378752E-5" and j(np.float32(0.001)) == "0.0010000000474974513"


@pytest.mark.parametrize("name", ["tail_merge_kat", "identical_pair", "interleaved_genomes", "single_genome", "repeats"])
def test_filter_on_literal_fixtures_emulated(emu_lib, name):
    w, k = fixtures.literal(name)
    check(w, k)


def test_filter_on_families_emulated(emu_lib):
    w = synth.generate(5, 40, 70.0, 0.12, 93)
    assert check(w, 3) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("shape,scale,k", [("salmonella7", 0.15, 4), ("mycoplasma64", 0.12, 4)])
def test_filter_on_config_shapes_gpu(engine_lib, shape, scale, k):
    native._lib = None
    native.load(native.ENGINE_LIB)
    w = synth.shape(shape, scale=scale)
    assert check(w, k) > 100


@pytest.mark.gpu
def test_native_cli_writes_the_reference_net(engine_lib, tmp_path):
    """pangenes -i -k -o (the line of pandelos.sh:73): .faa in, .net out, text identical to the restated Java writer."""
    import subprocess
    from pandelos_b200 import build
    cli = build.build_host()
    w = synth.generate(6, 150, 160.0, 0.1, 95)
    faa = str(tmp_path / "in.faa")
    w.write_faa(faa)
    out = str(tmp_path / "out.net")
    r = subprocess.run([cli, "-i", faa, "-k", "4", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Total cost:" in r.stdout and not any(ln.startswith("F{ ") for ln in r.stdout.splitlines())
    seqs, gids, _, _ = pangenes_java.read_faa(faa)
    w2 = synth.from_sequences(seqs, gids)
    assert (w2.residues == w.residues).all() and (w2.genome_of == w.genome_of).all()
    want = oracle_net(w2, 4).lines()
    got = open(out).read().splitlines()
    assert got == want and len(got) > 100
    # -c prints the cost report only (Pangenes.java:33-36); a missing option is a usage error
    r = subprocess.run([cli, "-i", faa, "-k", "4", "-o", out + ".x", "-c"], capture_output=True, text=True)
    assert r.returncode == 0 and "Total cost:" in r.stdout and not os.path.exists(out + ".x")
    assert subprocess.run([cli, "-i", faa], capture_output=True).returncode != 0


GOLD = os.path.join(HERE, "golden", "net")


def _golden_workload():
    seqs, gids, _, _ = pangenes_java.read_faa(os.path.join(GOLD, "family5.faa"))
    return synth.from_sequences(seqs, gids), 4


def test_golden_net_from_the_reference_emulated(emu_lib):
    """tests/golden/net/family5.net was produced from the UNMODIFIED reference library's scores (make_net_golden.py):
    the C oracle + restated filter, and the engine's device filter, must both reproduce it line for line."""
    w, k = _golden_workload()
    want = open(os.path.join(GOLD, "family5.net")).read().splitlines()
    assert oracle_net(w, k).lines() == want
    pn = native.PangeneNative(k, native.PangeneIData(w.residues, w.offsets, w.genome_of))
    try:
        assert engine_net(pn, w.genome_of).lines() == want
    finally:
        pn.close()


@pytest.mark.skipif(not os.path.exists("/root/reference/netclu_ng.py"), reason="the reference's netclu_ng.py is only in the build container")
def test_golden_clus_is_what_netclu_ng_makes_of_the_golden_net(tmp_path):
    import subprocess
    r = subprocess.run([sys.executable, "/root/reference/netclu_ng.py", os.path.join(GOLD, "family5.faa"), os.path.join(GOLD, "family5.net")],
                       capture_output=True, text=True, check=True)
    fams = sorted(set(ln.replace("F{ ", "").replace("}", "").replace(" ;", "").strip() for ln in r.stdout.splitlines() if ln.startswith("F{ ")))
    assert fams == open(os.path.join(GOLD, "family5.clus")).read().splitlines()


@pytest.mark.gpu
def test_native_cli_on_the_golden_input(engine_lib, tmp_path):
    import subprocess
    from pandelos_b200 import build
    out = str(tmp_path / "family5.net")
    r = subprocess.run([build.build_host(), "-i", os.path.join(GOLD, "family5.faa"), "-k", "4", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out).read() == open(os.path.join(GOLD, "family5.net")).read()
    # --clus: the families netclu_ng.py + pandelos.sh:79 make of that network (the reference chain leaves "name " for a
    # gene outside the network; the committed golden has the lines stripped)
    clus = str(tmp_path / "family5.clus")
    r = subprocess.run([build.CLI_BIN, "-i", os.path.join(GOLD, "family5.faa"), "-k", "4", "-o", out, "--clus", clus], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = open(clus).read().splitlines()
    assert got == sorted(got) and sorted(ln.strip() for ln in got) == open(os.path.join(GOLD, "family5.clus")).read().splitlines()


def test_native_cli_over_the_emulated_engine(emu_lib, tmp_path):
    """csrc/host/pangenes_main.cpp linked with the emulator build of the engine (same C ABI): the CLI's .faa reader, its
    genome tasks, edge bookkeeping, .net writer and --clus clustering run on the CPU against the reference-made goldens."""
    import subprocess
    import build_emu
    from pandelos_b200 import build
    lib = build_emu.build()
    exe = str(tmp_path / "pangenes_emu")
    host = os.path.join(build.CSRC, "host")
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I", build.INCLUDE, "-I", host, "-o", exe,
                    os.path.join(host, "pangenes_main.cpp"), lib, "-Wl,-rpath," + os.path.dirname(lib), "-pthread"], check=True)
    out, clus = str(tmp_path / "family5.net"), str(tmp_path / "family5.clus")
    r = subprocess.run([exe, "-i", os.path.join(GOLD, "family5.faa"), "-k", "4", "-o", out, "--clus", clus], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert open(out).read() == open(os.path.join(GOLD, "family5.net")).read()
    got = open(clus).read().splitlines()
    assert got == sorted(got) and sorted(ln.strip() for ln in got) == open(os.path.join(GOLD, "family5.clus")).read().splitlines()
    assert "Families: " in r.stdout and "Girvan-Newman" in r.stdout
