"""In-tree build of allocnet_b200/libmincob.so (nvcc, sm_100a only).

Eight (S, LPT) kernel objects + the host API object are compiled in parallel and linked into one
shared library next to this file, so it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.environ.get("MINCOB_BUILD_DIR") or os.path.join(HERE, "build")
LIB = os.environ.get("MINCOB_BUILD_OUT") or os.path.join(HERE, "libmincob.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
INST = [(S, L) for S in (3, 4) for L in (5, 8, 16, 32)]
# resident blocks per SM the optimize kernel is compiled for (caps registers per thread); override for
# experiments with MINCOB_MINB3 / MINCOB_MINB4 in the environment
EXTRA = os.environ.get("MINCOB_EXTRA_FLAGS", "").split()
MINB = {3: int(os.environ.get("MINCOB_MINB3", "3")), 4: int(os.environ.get("MINCOB_MINB4", "2"))}
# threads per block per LPT: the LPT = 5 objects (six trajectories per warp) use one-warp blocks so that shared memory,
# not the block granularity, decides how many warps fit (11 per SM at N = 5, K = 16); MINB scales to the same register cap
THREADS = {5: int(os.environ.get("MINCOB_THREADS5", "32")), 8: int(os.environ.get("MINCOB_THREADS8", "128"))}


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libmincob.so cannot be built (there is no CPU fallback)")
    return exe


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, "..", "include", "mincob.h")]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def _run(cmd, log):
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"build step failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return r.stderr


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for S, L in INST:
        o = os.path.join(OBJ, f"kernels_s{S}_l{L}.o")
        thr = THREADS.get(L, 128)
        minb = MINB.get(S, 2) * 128 // thr
        jobs.append(([nvcc, *ARCH, *FLAGS, f"-DMINCOB_S={S}", f"-DMINCOB_LPT={L}", f"-DMINCOB_MINB={minb}", f"-DMINCOB_THREADS={thr}", *EXTRA, "-c",
                      os.path.join(CSRC, "kernels_inst.cu"), "-o", o], o))
    o = os.path.join(OBJ, "mincob.o")
    jobs.append(([nvcc, *ARCH, *FLAGS, *EXTRA, "-c", os.path.join(CSRC, "mincob.cu"), "-o", o], o))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        outs = list(ex.map(lambda j: _run(j[0], j[1] + ".log"), jobs))
    if verbose:
        for (cmd, o), out in zip(jobs, outs):
            print(os.path.basename(o))
            print("\n".join(l for l in out.splitlines() if "registers" in l or "spill" in l or "Compiling" in l))
    _run([nvcc, *ARCH, "-shared", "-o", LIB, *[j[1] for j in jobs], "-ldl"], os.path.join(OBJ, "link.log"))
    return LIB


STRICT_LIB = os.path.join(HERE, "libmincob_strict.so")


def build_strict_library(force: bool = False) -> str:
    """The -DMINCOB_STRICT=1 variant (reference-form divisions / square roots in the L-BFGS decisions), S = 3 with 5 and 8
    lanes only: a test artefact that the GPU parity suite compares the shipped build with."""
    if not force and os.path.exists(STRICT_LIB) and all(os.path.getmtime(s) <= os.path.getmtime(STRICT_LIB) for s in _sources()):
        return STRICT_LIB
    nvcc = _nvcc()
    obj = OBJ + "_strict"
    os.makedirs(obj, exist_ok=True)
    jobs = []
    for S, L in INST:
        thr = THREADS.get(L, 128)
        minb = MINB.get(S, 2) * 128 // thr
        o = os.path.join(obj, f"kernels_s{S}_l{L}.o")
        jobs.append(([nvcc, *ARCH, *FLAGS, "-DMINCOB_STRICT=1", f"-DMINCOB_S={S}", f"-DMINCOB_LPT={L}", f"-DMINCOB_MINB={minb}",
                      f"-DMINCOB_THREADS={thr}", "-c", os.path.join(CSRC, "kernels_inst.cu"), "-o", o], o))
    o = os.path.join(obj, "mincob.o")
    jobs.append(([nvcc, *ARCH, *FLAGS, "-c", os.path.join(CSRC, "mincob.cu"), "-o", o], o))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        list(ex.map(lambda j: _run(j[0], j[1] + ".log"), jobs))
    _run([nvcc, *ARCH, "-shared", "-o", STRICT_LIB, *[j[1] for j in jobs], "-ldl"], os.path.join(obj, "link.log"))
    return STRICT_LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    print(build_strict_library(force="--force" in sys.argv))
