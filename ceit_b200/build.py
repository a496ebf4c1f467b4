"""Build libceit_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m ceit_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libceit_b200.so")
SOURCES = ["api.cu", "plan.cpp"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "--expt-relaxed-constexpr",
] + (["-DCEIT_TILE_CLOCKS"] if os.environ.get("CEIT_TILE_CLOCKS") else [])


STAMP = LIB + ".srchash"


def source_hash():
    """Content hash of every native source + the public header + the flags (mtimes do not survive a snapshot)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def is_current():
    try:
        with open(STAMP) as fh:
            return os.path.exists(LIB) and fh.read().strip() == source_hash()
    except OSError:
        return False


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and is_current():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for s in SOURCES:
        o = os.path.join(LIBDIR, s.rsplit(".", 1)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        subprocess.run(cmd, check=True)
        objs.append(o)
    subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"], check=True)
    with open(STAMP, "w") as fh:
        fh.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
