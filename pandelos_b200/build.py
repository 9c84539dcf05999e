"""In-tree builds (no JIT cache): every artefact lands next to its sources so it travels with a repo snapshot.

    python -m pandelos_b200.build [all|engine|jni|host|synth]

* libpandelos_b200.so   sm_100a kernels + engine + the plain C ABI of include/pandelos_b200.h
* libnative.so          drop-in for the reference's JNI library (loads under System.loadLibrary("native")); needs
                        JNI headers: $JAVA_HOME/include, else the reference's vendored copy by include path
* pangenes              native CLI with the reference's Pangenes flags
* calculate_k           native drop-in for the reference's calculate_k.py (CPU only)
* libpdsynth.so         synthetic workload generator
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("PD_CXX", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

ENGINE_LIB = os.path.join(HERE, "libpandelos_b200.so")
JNI_LIB = os.path.join(HERE, "libnative.so")
CLI_BIN = os.path.join(HERE, "pangenes")
CALCK_BIN = os.path.join(HERE, "calculate_k")
NETCLU_BIN = os.path.join(HERE, "netclu_cc")
NETCHECK_BIN = os.path.join(HERE, "net_check")
SYNTH_LIB = os.path.join(HERE, "libpdsynth.so")


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: %s" % cmd[0])
    return r.stdout + r.stderr


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(*dirs, exts=(".cu", ".cuh", ".cpp", ".h", ".hpp")):
    out = []
    for d in dirs:
        for base, _, files in os.walk(d):
            out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return out


def build_synth(force=False):
    src = os.path.join(CSRC, "synth.cpp")
    if force or _stale(SYNTH_LIB, [src]):
        _run([CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-o", SYNTH_LIB, src])
    return SYNTH_LIB


def build_engine(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("engine.cu", "c_api.cu")]
    deps = _sources(CSRC, INCLUDE)
    if force or _stale(ENGINE_LIB, deps):
        cmd = [NVCC] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-pthread", "-shared",
                              "-I", INCLUDE, "-I", CSRC, "-o", ENGINE_LIB] + os.environ.get("PD_NVCC_FLAGS", "").split() + srcs + ["-lcudart"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        out = _run(cmd)
        if verbose:
            print(out)
    return ENGINE_LIB


def jni_include_dirs():
    jh = os.environ.get("JAVA_HOME")
    if jh and os.path.exists(os.path.join(jh, "include", "jni.h")):
        return [os.path.join(jh, "include"), os.path.join(jh, "include", "linux")]
    ref = "/root/reference/ig/native/jni"
    if os.path.exists(os.path.join(ref, "jni.h")):
        return [ref, os.path.join(ref, "linux")]
    return None


def build_jni(force=False):
    """libnative.so = JNI shim linked against libpandelos_b200.so ($ORIGIN rpath). Skipped (prebuilt kept) without jni.h."""
    inc = jni_include_dirs()
    src = os.path.join(CSRC, "jni_shim.cpp")
    if inc is None or not os.path.exists(src):
        if not os.path.exists(JNI_LIB):
            sys.stderr.write("pandelos_b200.build: no jni.h (JAVA_HOME unset, reference absent); libnative.so not built\n")
        return JNI_LIB if os.path.exists(JNI_LIB) else None
    build_engine(force=False)
    if force or _stale(JNI_LIB, [src, ENGINE_LIB] + _sources(INCLUDE)):
        cmd = [CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I", INCLUDE] + sum((["-I", d] for d in inc), []) + \
              ["-o", JNI_LIB, src, "-L", HERE, "-lpandelos_b200", "-Wl,-rpath,$ORIGIN", "-pthread"]
        _run(cmd)
    return JNI_LIB


def build_host(force=False):
    """The native hosts: `pangenes` (links the engine); `calculate_k`, `netclu_cc` and `net_check` (CPU only, no engine)."""
    build_engine(force=False)
    hdir = os.path.join(CSRC, "host")
    if not os.path.isdir(hdir):
        return None
    headers = _sources(hdir, INCLUDE, exts=(".h", ".hpp"))
    src = os.path.join(hdir, "pangenes_main.cpp")
    if force or _stale(CLI_BIN, [src] + headers + [ENGINE_LIB]):
        _run([CXX, "-std=c++17", "-O2", "-Wall", "-ffp-contract=off", "-I", INCLUDE, "-I", hdir, "-o", CLI_BIN, src,
              "-L", HERE, "-lpandelos_b200", "-Wl,-rpath,$ORIGIN", "-pthread"])
    src = os.path.join(hdir, "calculate_k_main.cpp")
    if force or _stale(CALCK_BIN, [src] + headers):
        _run([CXX, "-std=c++17", "-O2", "-Wall", "-I", hdir, "-o", CALCK_BIN, src])
    src = os.path.join(hdir, "netclu_cc_main.cpp")
    if force or _stale(NETCLU_BIN, [src] + headers):
        # -ffp-contract=off: girvan_newman.h repeats networkx' float sums operation by operation
        _run([CXX, "-std=c++17", "-O2", "-Wall", "-ffp-contract=off", "-pthread", "-I", hdir, "-o", NETCLU_BIN, src])
    src = os.path.join(hdir, "net_check_main.cpp")
    if force or _stale(NETCHECK_BIN, [src] + headers):
        _run([CXX, "-std=c++17", "-O2", "-Wall", "-I", hdir, "-o", NETCHECK_BIN, src])
    return CLI_BIN


def build_all(force=False):
    build_synth(force)
    build_engine(force)
    build_jni(force)
    build_host(force)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    force = "--force" in sys.argv
    {"all": build_all, "engine": build_engine, "jni": build_jni, "host": build_host, "synth": build_synth}[what](force)
