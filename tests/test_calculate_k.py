"""calculate_k (native drop-in for the reference's calculate_k.py) against golden stdout of the
unmodified reference script (tests/golden/make_calculate_k_golden.py), against the package's own vectorised formula, and
its error behaviour.  CPU only."""
import glob
import os
import subprocess

import numpy as np
import pytest

from pandelos_b200 import build, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "calculate_k")


@pytest.fixture(scope="module")
def calck():
    build.build_host()
    assert os.path.exists(build.CALCK_BIN)
    return build.CALCK_BIN


def run(binary, path):
    return subprocess.run([binary, path], capture_output=True, text=True, timeout=120)


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.faa"))))
def test_stdout_identical_to_reference_script(calck, name):
    r = run(calck, os.path.join(GOLD, name + ".faa"))
    assert r.returncode == 0, r.stderr
    with open(os.path.join(GOLD, name + ".out")) as f:
        assert r.stdout == f.read()


def test_k_line_is_what_pandelos_sh_greps(calck):
    # pandelos.sh:68: k=`grep -E "^k =" "$tmp" | sed s/k\ =\ //g`
    r = run(calck, os.path.join(GOLD, "long.faa"))
    ks = [ln for ln in r.stdout.splitlines() if ln.startswith("k =")]
    assert len(ks) == 1 and int(ks[0].replace("k = ", "")) == 4


def write_faa(w, path):
    with open(path, "w") as f:
        for s in range(len(w.genome_of)):
            f.write("g%d\tgene%d\tp\n" % (int(w.genome_of[s]), s))
            f.write(w.residues[int(w.offsets[s]):int(w.offsets[s + 1])].tobytes().decode("latin-1") + "\n")


@pytest.mark.parametrize("shape,scale", [("salmonella7", 0.2), ("mycoplasma64", 0.2)])
def test_same_k_as_the_workload_formula(calck, tmp_path, shape, scale):
    w = synth.shape(shape, scale=scale)
    path = str(tmp_path / "w.faa")
    write_faa(w, path)
    r = run(calck, path)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "total length  %d" % len(w.residues)
    assert int(lines[-1].replace("k = ", "")) == synth.calculate_k(w)
    # alphabet counts, whatever the order
    counts = np.bincount(w.residues, minlength=256)
    body = lines[1][len("alphabet {"):-1]
    got = {kv.split(": ")[0].strip("'\""): int(kv.split(": ")[1]) for kv in body.split(", ")}
    assert got == {chr(b): int(counts[b]) for b in np.flatnonzero(counts)}


def test_errors(calck, tmp_path):
    assert run(calck, str(tmp_path / "missing.faa")).returncode == 1
    one = tmp_path / "one_letter.faa"
    one.write_text("g\tx\tp\nAAAA\n")
    r = run(calck, str(one))
    assert r.returncode == 1 and "ZeroDivisionError" in r.stderr  # math.log(x, 1) in the reference
    empty = tmp_path / "empty.faa"
    empty.write_text("")
    r = run(calck, str(empty))
    assert r.returncode == 1 and "math domain error" in r.stderr
    high = tmp_path / "latin1.faa"
    high.write_bytes(b"g\tx\tp\nMK\xe9V\n")
    assert run(calck, str(high)).returncode == 1
