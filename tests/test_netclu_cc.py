"""netclu_cc (native connected-component half of the reference's netclu_ng.py) against digests of
the unmodified script's stdout (tests/golden/make_netclu_golden.py): the families it prints as they stand, the
components it leaves for the Girvan-Newman split, the singletons; and — where the script is present — the whole
two-step pipeline of INTEGRATION.md against the script run once on the full network.  CPU only."""
import glob
import json
import os
import subprocess
import sys

import pytest

from pandelos_b200 import build

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "netclu")
SCRIPT = "/root/reference/netclu_ng.py"
CASES = sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLD, "*.json")))


@pytest.fixture(scope="module")
def netclu():
    build.build_host()
    assert os.path.exists(build.NETCLU_BIN)
    return build.NETCLU_BIN


def run(binary, name, rest):
    return subprocess.run([binary, os.path.join(GOLD, name + ".faa"), os.path.join(GOLD, name + ".net"), "-r", rest],
                          capture_output=True, text=True, timeout=120)


def clus(f_lines):
    # pandelos.sh:79: grep "F{ " | sed s/F{\ //g | sed s/}//g | sed s/\ \;//g | sort | uniq
    return sorted(set(ln.replace("F{ ", "").replace("}", "").replace(" ;", "").strip() for ln in f_lines))


def test_cases_present():
    assert len(CASES) >= 4


@pytest.mark.parametrize("name", CASES)
def test_families_components_and_singletons_match_the_script(netclu, name, tmp_path):
    want = json.load(open(os.path.join(GOLD, name + ".json")))
    rest = str(tmp_path / "rest.net")
    r = run(netclu, name, rest)
    assert r.returncode == (3 if want["split"] else 0), r.stderr
    out = r.stdout.splitlines()
    assert all(ln.startswith("F{ ") for ln in out)
    assert sorted(ln for ln in out if not ln.endswith(" }")) == want["kept"]
    assert sorted(ln for ln in out if ln.endswith(" }")) == want["singletons"]
    # rest.net: exactly the lines of the split components, verbatim, in file order
    members = set(s for comp in want["split"] for s in comp)
    src = [ln for ln in open(os.path.join(GOLD, name + ".net")).read().splitlines() if int(ln.split("\t")[0]) in members]
    assert open(rest).read().splitlines() == src
    for ln in src:
        assert int(ln.split("\t")[1]) in members


@pytest.mark.skipif(not os.path.exists(SCRIPT), reason="the reference's netclu_ng.py is only in the build container")
@pytest.mark.parametrize("name", CASES)
def test_two_step_pipeline_gives_the_scripts_clus(netclu, name, tmp_path):
    want = json.load(open(os.path.join(GOLD, name + ".json")))
    rest = str(tmp_path / "rest.net")
    r = run(netclu, name, rest)
    lines = r.stdout.splitlines()
    if r.returncode == 3:
        s = subprocess.run([sys.executable, SCRIPT, os.path.join(GOLD, name + ".faa"), rest], capture_output=True, text=True, check=True)
        split_lines = [ln for ln in s.stdout.splitlines() if ln.startswith("F{ ") and not ln.endswith(" }")]
        assert sorted(split_lines) == want["split_families"]
        lines += split_lines
    assert clus(lines) == clus(want["kept"] + want["split_families"] + want["singletons"])


def test_family5_clus_is_the_pangenes_golden(netclu, tmp_path):
    # the same digest, post-processed as pandelos.sh does, is the committed .clus of the Pangenes golden network
    want = json.load(open(os.path.join(GOLD, "family5.json")))
    gold = open(os.path.join(os.path.dirname(GOLD), "net", "family5.clus")).read().splitlines()
    assert clus(want["kept"] + want["split_families"] + want["singletons"]) == gold


def test_errors(netclu, tmp_path):
    faa = tmp_path / "a.faa"
    faa.write_text("G0\ta\td\nAAA\nG1\tb\td\nAAA\n")
    net = tmp_path / "a.net"
    net.write_text("0\t5\t1.0\n")                       # gene id outside the .faa
    assert subprocess.run([netclu, str(faa), str(net)], capture_output=True).returncode == 1
    net.write_text("0\t1\t1.0\n")
    r = subprocess.run([netclu, str(faa), str(net)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "F{ a ; b}\n"
    net.write_text("")
    r = subprocess.run([netclu, str(faa), str(net)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "F{ a }\nF{ b }\n"
    assert subprocess.run([netclu, str(tmp_path / "missing.faa"), str(net)], capture_output=True).returncode == 1
    faa.write_text("G0\ta\nAAA\n")                      # two columns: the script raises IndexError
    assert subprocess.run([netclu, str(faa), str(net)], capture_output=True).returncode == 1
