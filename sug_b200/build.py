"""Build libsug_b200.so in-tree with nvcc for sm_100a (B200).  No torch headers are involved:
the library is plain CUDA runtime code behind the C ABI of include/sug_b200.h."""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsug_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsug_b200.so cannot be built")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link libsug_b200.so.  Returns its path."""
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()

    def fresh():
        return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig
    if not force and fresh():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    # one builder at a time (torchrun starts one process per GPU): the others wait on the lock and then find
    # the finished library; the .so is linked under a temporary name and renamed, so a concurrent ctypes.CDLL
    # never opens a half-written file
    with open(os.path.join(OBJ, "lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():
                return LIB
            return _build_locked(dig, stamp, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(dig: str, stamp: str, verbose: bool) -> str:
    nvcc = _nvcc()
    logs = {}

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([nvcc, *NVCC_FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        logs[src] = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    fd, tmp = tempfile.mkstemp(prefix="libsug_b200.", suffix=".so.tmp", dir=HERE)
    os.close(fd)
    r = subprocess.run([nvcc, "-shared", "-o", tmp, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        os.unlink(tmp)
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.chmod(tmp, 0o755)
    os.replace(tmp, LIB)
    with open(os.path.join(OBJ, "ptxas.log"), "w") as fh:
        for s in sorted(logs):
            fh.write(f"==== {s}\n{logs[s]}\n")
    with open(stamp, "w") as fh:
        fh.write(dig)
    if verbose:
        print(open(os.path.join(OBJ, "ptxas.log")).read())
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
