"""Build libcellseg_sm100.so in-tree with nvcc for sm_100a.

The library is the product's only compute path; there is no JIT cache and no
fallback.  `python -m cellseg_b200.build` (or `__graft_entry__.build()`)
compiles every `.cu` under `csrc/` into `neurips-cellseg_b200/libcellseg_sm100.so`.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libcellseg_sm100.so")
STAMP = os.path.join(PKG_DIR, "build", "stamp.txt")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # blend arithmetic must round mul and add separately
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode()); h.update(fh.read())
    with open(os.path.join(PKG_DIR, "..", "include", "cellseg_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libcellseg_sm100.so cannot be built")
    return p


def build(force: bool = False, verbose: bool = False) -> str:
    dig = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB_PATH
    os.makedirs(os.path.dirname(STAMP), exist_ok=True)
    objs, procs, log = [], [], []
    for src in sources():                      # one nvcc per translation unit, all at once
        obj = os.path.join(PKG_DIR, "build", os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc_path(), *NVCC_FLAGS, "-I", CSRC, "-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, err = p.communicate()
        log.append(err)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out + err)
    cmd = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(os.path.join(PKG_DIR, "build", "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(STAMP, "w") as fh:
        fh.write(dig)
    if verbose:
        sys.stderr.write("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
