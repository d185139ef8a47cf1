"""In-tree build of libshgpu.so (CUDA sm_100a + C-ABI).  nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libshgpu.so")
SOURCES = ["shgpu_api.cu", "shape_tables.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                 # no implicit FMA contraction: every DFMA is an explicit fma()
    "--extended-lambda",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-shared",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    return False


def build(force=False, verbose=False):
    """Compile libshgpu.so in-tree if missing or stale; return its path."""
    if not force and not needs_build():
        return SO
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    cmd = [_nvcc(), "-ccbin", ccbin] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return SO


SHLMP = os.path.join(HERE, "shlmp")


def build_host(force=False):
    """Compile the shlmp input-script front-end (plain C++, links libshgpu.so)."""
    src = os.path.join(HERE, "host", "shlmp.cpp")
    if not force and os.path.exists(SHLMP) and os.path.getmtime(SHLMP) > max(os.path.getmtime(src), os.path.getmtime(SO)):
        return SHLMP
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    cmd = [gxx, "-O2", "-std=c++17", "-pthread", "-o", SHLMP, src, "-L" + HERE, "-lshgpu", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return SHLMP


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
