"""Build recipe for the engine: nvcc -> chroma_b200/libb200clover.so (sm_100a only, in-tree).

Usage: python -m chroma_b200.build [--force] [--verbose]
"""
import concurrent.futures
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# B200_BUILD_TAG=<tag> builds a tuning variant next to the product library (libb200clover_<tag>.so, own object dir);
# chroma_b200/lib.py loads it when B200_LIB_TAG=<tag> is set.  Variants are built HERE (nvcc cross-compiles without a GPU)
# so that GPU time is spent measuring, not compiling.
_TAG = os.environ.get("B200_BUILD_TAG", "")
OUT = os.path.join(HERE, "libb200clover%s.so" % ("_" + _TAG if _TAG else ""))
OBJDIR = os.path.join(HERE, "_obj" + ("_" + _TAG if _TAG else ""))
SOURCES = ["api.cu", "engine_d.cu", "engine_f.cu", "engine_mixed.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr", "-ccbin", "/usr/bin/g++",
]
# tuning knobs (defaults live in the sources): B200_DSLASH_BLOCK, B200_DSLASH_MINBLOCKS
for _k in ("B200_DSLASH_BLOCK", "B200_DSLASH_MINBLOCKS", "B200_DSLASH_BLOCK_F", "B200_DSLASH_MINBLOCKS_F",
           "B200_MRHS_NRB", "B200_MRHS_MINB", "B200_MRHS_NRB_F", "B200_MRHS_MINB_F", "B200_MRHS_PREFETCH", "B200_MRHS_PREFETCH_F", "B200_MRHS_L1", "B200_MRHS_DEPTH", "B200_MRHS_CLOVER_LATE", "B200_CLOV_MINB", "B200_RED_RELEASE", "B200_RED_PROBE"):
    if os.environ.get(_k):
        FLAGS += ["-D%s=%s" % (_k, os.environ[_k])]


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def up_to_date():
    return os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_source_mtime()


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    os.makedirs(OBJDIR, exist_ok=True)
    t_start = time.time()
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
                                                 "-Xcompiler", "-fPIC", "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.utime(OUT, (t_start, t_start))      # a source edited WHILE this build ran must make the library stale
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
