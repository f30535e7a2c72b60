"""Build libtsp_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

Every csrc/*.cu is compiled to its own object (in parallel, only when it or a header changed) and linked."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["api.cu", "percentile.cu", "fir.cu", "band.cu", "spm.cu", "fast.cu", "binned.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(os.path.dirname(HERE), "include", "tsp_b200.h"),
           os.path.abspath(__file__)]
LIB = os.path.join(HERE, "libtsp_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-O2,-Wall"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(LIB, [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + HEADERS)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not force and not _stale(obj, [path] + HEADERS):
        return obj, ""
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
    return obj, res.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(lambda s: _compile(s, force, verbose), SOURCES))
    if verbose:
        sys.stderr.write("".join(log for _, log in results))
    cmd = [NVCC, "--shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"] + \
          [obj for obj, _ in results] + ["-ldl", "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libtsp_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
