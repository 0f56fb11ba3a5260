"""Build recipe of libagofrt.so (nvcc, sm_100a only).  In-tree output: analisi_b200/libagofrt.so.

    python -m analisi_b200.build [--force]

nvcc cross-compiles without a GPU.  The flags matter for parity: no fast-math, no FMA contraction
on the host side (the threshold table is computed with the reference's own expression), -lineinfo
so ncu's source page maps to the .cu files.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libagofrt.so")
SOURCES = ["agofrt_kernels.cu", "agofrt_cabi.cu"]
HEADERS = [os.path.join(CSRC, "agofrt_kernels.cuh"), os.path.join(ROOT, "include", "agofrt.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-fno-fast-math",
    "-Xptxas", "-v",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defs=None, out=None):
    """Compile libagofrt.so if it is missing or older than its sources; return its path.

    ``defs`` (e.g. ["-DAGOFRT_IPT=4"]) and ``out`` build a tuning variant next to the product library."""
    global LIB
    if out is not None:
        LIB = os.path.join(HERE, out)
        force = True
    if not force and not stale():
        return LIB
    objs = []
    log = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", ".o"))
        o = o if out is None else o.replace(".o", "." + out + ".o")
        cmd = [nvcc()] + NVCC_FLAGS + list(defs or []) + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c",
                                                          os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append(r.stdout)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout)
        objs.append(o)
    cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


# ---- host side: the C++ classes mirroring the reference API (include/analisi, src/host), the CLI and
# the pybind11 module.  Plain g++ (no CUDA headers needed: they only see include/agofrt.h). ----------
HOST_SOURCES = ["device.cpp", "trajectory.cpp", "trajectory_numpy.cpp"]
CLI = os.path.join(ROOT, "bin", "analisi")
HOST_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-Wno-sign-compare", "-Wno-unused-variable"]


def host_cxx():
    # $CXX in this image is /opt/gcc/bin/g++, a wrapper that links libstdc++ statically: a python
    # extension built with it crashes as soon as another libstdc++ is in the process.  ANALISI_CXX overrides.
    for cand in (os.environ.get("ANALISI_CXX"), "/usr/bin/g++", "g++"):
        if cand and (not os.path.isabs(cand) or os.path.exists(cand)):
            return cand
    raise RuntimeError("g++ not found")


def pyext_path():
    import sysconfig
    return os.path.join(ROOT, "python", "pyanalisi" + sysconfig.get_config_var("EXT_SUFFIX"))


def _host_deps():
    deps = [os.path.abspath(__file__), LIB, os.path.join(ROOT, "include", "agofrt.h")]
    for d in (os.path.join(ROOT, "include", "analisi"), os.path.join(ROOT, "src", "host")):
        deps += [os.path.join(d, f) for f in os.listdir(d)]
    return deps


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build failed:\n" + " ".join(cmd) + "\n" + r.stdout)
    return r.stdout


def build_host(force=False):
    """Compile bin/analisi (the CLI) and python/pyanalisi*.so (pybind11) against libagofrt.so; return both paths."""
    build()
    ext = pyext_path()
    srcs = [os.path.join(ROOT, "src", "host", s) for s in HOST_SOURCES]
    link = ["-L", HERE, "-lagofrt", "-Wl,-rpath,$ORIGIN/../analisi_b200"]
    inc = ["-I", os.path.join(ROOT, "include")]
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    main = os.path.join(ROOT, "cli", "main.cpp")
    deps = _host_deps()
    if force or not os.path.exists(CLI) or any(os.path.getmtime(d) > os.path.getmtime(CLI) for d in deps + [main]):
        _run([host_cxx()] + HOST_FLAGS + inc + [main] + srcs + link + ["-o", CLI])
    mod = os.path.join(ROOT, "python", "pyanalisi.cpp")
    if force or not os.path.exists(ext) or any(os.path.getmtime(d) > os.path.getmtime(ext) for d in deps + [mod]):
        import pybind11
        import sysconfig
        _run([host_cxx()] + HOST_FLAGS + ["-shared", "-fvisibility=hidden", "-DANALISI_WITH_PYBIND11"] + inc +
             ["-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"], mod] + srcs + link + ["-o", ext])
    return CLI, ext


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--host" in sys.argv:
        print(build_host(force="--force" in sys.argv))
