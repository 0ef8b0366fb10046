"""Builds crypto12381_b200/libc12381_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m crypto12381_b200.build [--force]

One object per translation unit (compiled in parallel), linked into one shared library whose exported symbols are
exactly the C ABI of include/c12381_cuda.h.  Objects are rebuilt when their source, any header in csrc/, or the
flags change."""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libc12381_cuda.so")
UNITS = ["ctx", "msm_common", "msm_g1", "msm_g2", "pairing", "miracl_pod", "hash"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _headers_digest() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".inc", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    return h.hexdigest()


def _compile(unit: str, digest: str, force: bool) -> bool:
    src = os.path.join(CSRC, unit + ".cu")
    obj = os.path.join(OBJ, unit + ".o")
    stamp = os.path.join(OBJ, unit + ".stamp")
    with open(src, "rb") as fh:
        want = hashlib.sha256(digest.encode() + fh.read()).hexdigest()
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return False
    log = os.path.join(OBJ, unit + ".ptxas.log")
    with open(log, "w") as lf:
        r = subprocess.run([_nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj], stdout=lf, stderr=subprocess.STDOUT, timeout=1800)
    if r.returncode != 0:
        sys.stderr.write(open(log).read()[-8000:])
        raise RuntimeError(f"nvcc failed on {unit}.cu")
    with open(stamp, "w") as fh:
        fh.write(want)
    return True


def build(force: bool = False, verbose: bool = True, variant: str = "", defines=()) -> str:
    """variant/defines: build an experimental twin `libc12381_cuda_<variant>.so` with extra -D flags (A/B measurements;
    selected at load time with the environment variable C12381_LIB_VARIANT)."""
    global OBJ, NVCC_FLAGS
    lib = LIB if not variant else LIB.replace(".so", f"_{variant}.so")
    obj_dir, flags = OBJ, NVCC_FLAGS
    if variant:
        OBJ = os.path.join(CSRC, "_obj_" + variant)
        NVCC_FLAGS = NVCC_FLAGS + [f"-D{d}" for d in defines]
    try:
        os.makedirs(OBJ, exist_ok=True)
        digest = _headers_digest()
        with concurrent.futures.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
            rebuilt = list(ex.map(lambda u: _compile(u, digest, force), UNITS))
        if any(rebuilt) or not os.path.exists(lib):
            objs = [os.path.join(OBJ, u + ".o") for u in UNITS]
            subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs, check=True, timeout=600)
            if verbose:
                print(f"built {lib} ({sum(rebuilt)} unit(s) recompiled)")
        elif verbose:
            print(f"{lib} is up to date")
    finally:
        OBJ, NVCC_FLAGS = obj_dir, flags
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        build(force="--force" in sys.argv, variant=sys.argv[i + 1], defines=sys.argv[i + 2].split(","))
    else:
        build(force="--force" in sys.argv)
