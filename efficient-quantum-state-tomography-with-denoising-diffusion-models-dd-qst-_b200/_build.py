"""In-tree build of libddqst.so (sm_100a only).  nvcc cross-compiles without a GPU.

Staleness is decided by CONTENT, not mtime: every object carries a ``.hash`` side file = sha256 of (its source, every
header under csrc/, include/ddqst.h, the nvcc flags, the nvcc version string), and the library carries a ``.stamp`` =
sha256 of the object hashes.  A tree that arrives with objects built from other sources (a stale checkout, a snapshot
pushed to a GPU box) is therefore rebuilt, and ``stamp_matches()`` lets ``_lib.load()`` refuse a mismatching library
when no compiler is available.  ``build(force=True)`` recompiles everything regardless."""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libddqst.so")
STAMP = LIB + ".stamp"
SOURCES = ["api.cu", "simt.cu", "recon.cu", "eig.cu", "sampler_tc.cu", "train.cu", "train_tc.cu", "train_fused.cu", "mlp.cu",
           "dataset.cu", "synth.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# developer builds: DDQST_NVCC_DEFINES="DDQST_JL_PROFILE ..." adds -D switches (part of the content hash like every other flag)
NVCC_FLAGS += ["-D" + d for d in os.environ.get("DDQST_NVCC_DEFINES", "").split()]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers():
    hs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    hs.append(os.path.join(HERE, "..", "include", "ddqst.h"))
    return hs


def _common_digest() -> "hashlib._Hash":
    h = hashlib.sha256()
    for p in _headers():
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h


def _object_hash(src: str, common) -> str:
    h = common.copy()
    h.update(src.encode())
    h.update(open(os.path.join(CSRC, src), "rb").read())
    return h.hexdigest()


def expected_stamp() -> str:
    common = _common_digest()
    return hashlib.sha256("".join(_object_hash(s, common) for s in _sources()).encode()).hexdigest()


def stamp_matches() -> bool:
    try:
        return os.path.exists(LIB) and open(STAMP).read().strip() == expected_stamp()
    except OSError:
        return False


def _read(path):
    try:
        return open(path).read().strip()
    except OSError:
        return None


def build(force: bool = False, verbose: bool = False) -> str:
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)          # torchrun ranks all call load(): one builds, the rest find it current
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> str:
    common = _common_digest()
    objs, jobs, hashes = [], [], []
    for src in _sources():
        o = os.path.join(CSRC, src[:-3] + ".o")
        want = _object_hash(src, common)
        objs.append(o)
        hashes.append(want)
        if force or not os.path.exists(o) or _read(o + ".hash") != want:
            jobs.append((os.path.join(CSRC, src), o, want))

    def compile_one(job):
        s, o, want = job
        r = subprocess.run([_nvcc(), *NVCC_FLAGS, "-c", s, "-o", o], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{r.stdout}\n{r.stderr}")
        with open(o + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        with open(o + ".hash", "w") as f:
            f.write(want)
        if verbose:
            sys.stderr.write(r.stderr)

    with ThreadPoolExecutor(max_workers=min(12, os.cpu_count() or 4)) as ex:
        list(ex.map(compile_one, jobs))
    stamp = hashlib.sha256("".join(hashes).encode()).hexdigest()
    if force or jobs or not os.path.exists(LIB) or _read(STAMP) != stamp:
        r = subprocess.run([_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                            "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        with open(STAMP, "w") as f:
            f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
