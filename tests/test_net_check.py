"""net_check: the `check` command of the reference's Rust harness (benchmark/test-framework/src/main.rs:129-169,
verify.rs:48-86) restated — unordered pairs, first of duplicate pairs kept, float32 weights, tolerance 0.001, the harness'
own lines.  No cargo here, so the expected text is written from the source, not captured from the harness.  CPU only."""
import os
import subprocess

import pytest

from pandelos_b200 import build

NET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "net")


@pytest.fixture(scope="module")
def net_check():
    build.build_host()
    assert os.path.exists(build.NETCHECK_BIN)
    return build.NETCHECK_BIN


def run(binary, a, b):
    r = subprocess.run([binary, str(a), str(b)], capture_output=True, text=True)
    return r.returncode, r.stdout.splitlines()


def test_pairs_are_unordered_first_duplicate_wins_float32_tolerance(net_check, tmp_path):
    a, b = tmp_path / "a.net", tmp_path / "b.net"
    a.write_text("0\t1\t0.5\n2\t1\t0.25\n1\t0\t0.9\n3\t4\t1.0\n5\t6\t0.3333333432674408\n")
    b.write_text("1 0 0.5005\n1\t2\t0.2515\n7\t3\t0.125\n6\t5\t0.3333333333333333\n")
    rc, out = run(net_check, a, b)
    assert rc == 2
    assert out == ["1 <-> 2 = 0.25 ~ 0.2515",          # |0.25 - 0.2515| > 0.001
                   "MissingA 3 <-> 4 weight: 1",       # f32 Display: shortest digits, no exponent, no trailing .0
                   "MissingB 3 <-> 7 weight: 0.125",
                   "Values 1+1+(1) / 3"]               # (0,1): 0.5 kept (the later 0.9 is a duplicate pair), 0.5 ~ 0.5005 within 0.001
    rc, out = run(net_check, b, a)
    assert rc == 2 and out == ["1 <-> 2 = 0.2515 ~ 0.25", "MissingA 3 <-> 7 weight: 0.125", "MissingB 3 <-> 4 weight: 1", "Values 1+1+(1) / 3"]


def test_same_network_in_another_line_order_and_direction(net_check, tmp_path):
    src = open(os.path.join(NET, "family5.net")).read().splitlines()
    flipped = ["%s\t%s\t%s" % (c[1], c[0], c[2]) for c in (ln.split("\t") for ln in reversed(src))]
    other = tmp_path / "flipped.net"
    other.write_text("\n".join(flipped) + "\n")
    rc, out = run(net_check, os.path.join(NET, "family5.net"), other)
    pairs = len(set(tuple(sorted(map(int, ln.split("\t")[:2]))) for ln in src))
    assert rc == 0 and out == ["Values 0+0+(0) / %d" % pairs]


def test_missing_file_is_an_empty_network_and_malformed_lines_fail(net_check, tmp_path):
    a = tmp_path / "a.net"
    a.write_text("5\t6\t0.3333333432674408\n")
    rc, out = run(net_check, a, tmp_path / "nothing.net")
    assert rc == 2 and out == ["MissingA 5 <-> 6 weight: 0.33333334", "Values 1+0+(0) / 0"]
    rc, out = run(net_check, tmp_path / "nothing.net", a)
    assert rc == 2 and out == ["MissingB 5 <-> 6 weight: 0.33333334", "Values 0+1+(0) / 0"]
    a.write_text("5\t6\n")
    assert run(net_check, a, a)[0] == 1
    assert subprocess.run([net_check, str(a)], capture_output=True).returncode == 1
