"""Build recipe for the native pieces (run here on CPU: nvcc cross-compiles sm_100a).

  libaby3cu.so  -- CUDA kernels + C ABI (include/aby3cu.h), one translation unit
  libsh3.so     -- the C++ sh3 facade (aby3_b200/sh3) + its C harness, links libaby3cu
(The CPU checker under the repo's top-level checker directory has its own Makefile
and is built by __graft_entry__.build() / tests/conftest.py, not from here.)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "aby3_b200")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def _sources(d, exts):
    out = []
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return out


def build_cuda(force=False):
    so = os.path.join(PKG, "libaby3cu.so")
    srcs = _sources(os.path.join(PKG, "csrc"), (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "aby3cu.h")]
    if force or _newer(so, srcs):
        # ABY3CU_NVCC_DEFS="-DABY3CU_GEMM_STAGES=4 ..." : experiment switches of the kernels
        extra = os.environ.get("ABY3CU_NVCC_DEFS", "").split()
        _run([NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", *extra,
              "-o", so, os.path.join(PKG, "csrc", "aby3cu_all.cu")])
    return so


def build_sh3(force=False):
    so = os.path.join(PKG, "libsh3.so")
    d = os.path.join(PKG, "sh3")
    if not os.path.isdir(d):
        return None
    srcs = _sources(d, (".cpp", ".h")) + [os.path.join(ROOT, "include", "aby3cu.h")]
    cpps = [s for s in srcs if s.endswith(".cpp")]
    # headers of the app layers the harness instantiates (aby3-ML, aby3-Basic)
    srcs += _sources(os.path.join(PKG, "ml"), (".h",)) + _sources(os.path.join(PKG, "basic"), (".h",))
    if not cpps:
        return None
    if force or _newer(so, srcs + [os.path.join(PKG, "libaby3cu.so")]):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall",
              "-I", ROOT, "-I", os.path.join(ROOT, "include"), "-o", so, *cpps,
              "-L", PKG, "-laby3cu", "-Wl,-rpath,$ORIGIN"])
    return so


def build_compat():
    """compat/_build/libcompat_apps.so: the reference's application sources against the facade headers (compat/Makefile).
    Rebuilt whenever the facade changes -- it embeds the facade's class layouts.  Needs the reference tree; elsewhere the
    prebuilt library is used as it is."""
    ref = os.environ.get("ABY3_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "aby3-ML")) or not os.path.isdir(os.path.join(ROOT, "compat")):
        return None
    _run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "compat"), "REF=" + ref])
    return os.path.join(ROOT, "compat", "_build", "libcompat_apps.so")


def build_all(force=False):
    out = [build_cuda(force), build_sh3(force)]
    c = build_compat()
    return out + ([c] if c else [])


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv):
        print("built", p)
