"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/pandelos_b200.h declares.
No compute calls here — without a device every entry point must refuse loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_gpu
from pandelos_b200 import build, native

HEADER = os.path.join(ROOT, "include", "pandelos_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pd_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build_engine()


def test_header_symbols_exported(lib_path):
    names = declared_symbols()
    assert "pd_build" in names and "pd_compute_scores" in names and len(names) >= 12
    L = C.CDLL(lib_path)
    for n in names:
        assert hasattr(L, n), "missing export %s" % n


def test_library_is_sm100a_only(lib_path):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_no_device_is_an_error_not_a_fallback(lib_path):
    native.load(lib_path)
    data = native.PangeneIData(np.frombuffer(b"ACDEFGHIKL", np.uint8), np.array([0, 10], np.uint64), np.array([0], np.uint32))
    with pytest.raises(native.PdError) as e:
        native.PangeneNative(3, data)
    assert e.value.code == native.PD_ERR_NO_DEVICE
    native._lib = None


def test_jni_library_exports_reference_symbols():
    """libnative.so must export exactly the two symbols of the reference's pangene_native.h:16-25."""
    path = build.build_jni()
    if path is None:
        pytest.skip("no jni.h available (JAVA_HOME unset and reference absent) and no prebuilt libnative.so")
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    syms = set(re.findall(r" T (Java_\w+)", out))
    assert syms == {"Java_infoasys_cli_pangenes_PangeneNative_preprocessSequences",
                    "Java_infoasys_cli_pangenes_PangeneNative_computeScores"}
