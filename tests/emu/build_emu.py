"""TEST INFRASTRUCTURE — builds the engine's .cu sources with g++ against the SIMT logic emulator (cuemu.h) into
tests/emu/_build/libpandelos_emu.so.  Never shipped, never loaded by the pandelos_b200 package on its own: tests
point `native.load()` at it explicitly to check kernel LOGIC against the oracle in a container without a GPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pandelos_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libpandelos_emu.so")


def build(force=False):
    srcs = [os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "c_api.cu"), os.path.join(HERE, "cuemu.cpp")]
    deps = srcs + [os.path.join(HERE, "cuemu.h"), os.path.join(ROOT, "include", "pandelos_b200.h")] + \
        [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-DPD_EMU", "-x", "c++", "-I", os.path.join(ROOT, "include"),
           "-I", CSRC, "-I", HERE, "-o", OUT] + srcs + ["-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + r.stderr[-4000:])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
