"""In-tree build of the CUDA library: nvcc -> feprogram_b200/libfeprogram_b200.so (sm_100a only).

    python -m feprogram_b200.build [--force] [--verbose]

Cross-compiles without a GPU.  The built .so is git-ignored but travels to the GPU box with the
repository snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "csrc", "build")
LIB = os.path.join(PKG, "libfeprogram_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def _nccl_paths():
    """NCCL header + library: prefer the torch-bundled one (it is what torch.distributed uses)."""
    try:
        import nvidia.nccl as n  # type: ignore
        base = os.path.dirname(n.__file__) if getattr(n, "__file__", None) else list(n.__path__)[0]
        inc, lib = os.path.join(base, "include"), os.path.join(base, "lib")
        if os.path.exists(os.path.join(inc, "nccl.h")):
            return inc, lib
    except Exception:
        pass
    if os.path.exists("/usr/include/nccl.h"):
        return "/usr/include", "/usr/lib/x86_64-linux-gnu"
    return None, None


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(srcs) -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "feprogram_b200.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("FE_EXTRA_DEFS", "").encode())     # developer-only extra -D flags are part of the identity
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp(srcs)
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    nvcc = _nvcc()
    inc, nccl_lib = _nccl_paths()
    extra_inc = ["-I", inc, "-DFE_WITH_NCCL"] if inc else []
    if os.environ.get("FE_EXTRA_DEFS"):       # developer builds only (profiling macros); hashed into the stamp
        extra_inc += os.environ["FE_EXTRA_DEFS"].split()

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra_inc, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(OBJ, src[:-3] + ".ptxas.log"), "w") as f:
            f.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    if inc and nccl_lib:
        # NCCL symbols resolve at load time against the libnccl.so.2 torch has already loaded (same soname);
        # the rpath covers a process that loads this library before torch.
        so = os.path.join(nccl_lib, "libnccl.so.2")
        link += ["-L", nccl_lib, "-l:libnccl.so.2" if os.path.exists(so) else "-lnccl",
                 "-Xlinker", "-rpath", "-Xlinker", nccl_lib]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
