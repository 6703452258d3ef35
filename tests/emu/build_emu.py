"""TEST INFRASTRUCTURE ONLY: compile csrc/*.cu with g++ against the fiber
emulator (cuda_emu.h) into tests/emu/libcellseg_emu.so so the CPU-only CI can
check kernel logic.  The package never loads this library."""
from __future__ import annotations

import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "neurips-cellseg_b200", "csrc")
LIB = os.path.join(HERE, "libcellseg_emu.so")
STAMP = os.path.join(HERE, "build", "stamp.txt")


def _digest() -> str:
    h = hashlib.sha256()
    for d in (CSRC, HERE):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(f.encode()); h.update(fh.read())
    with open(os.path.join(ROOT, "include", "cellseg_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False) -> str:
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    objs = []
    flags = ["-O2", "-g", "-std=c++17", "-fPIC", "-DCS_EMU", "-I", HERE, "-I", CSRC, "-ffp-contract=off",
             "-Wno-unused-variable", "-Wno-unused-function"]
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    procs = []
    for src in srcs:
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        procs.append((src, subprocess.Popen(["g++", *flags, "-x", "c++", "-c", src, "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    obj = os.path.join(HERE, "build", "cuda_emu.o")
    procs.append(("cuda_emu.cpp", subprocess.Popen(["g++", *flags, "-c", os.path.join(HERE, "cuda_emu.cpp"), "-o", obj],
                                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs.append(obj)
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"emu build failed for {name}:\n{out}")
    r = subprocess.run(["g++", "-shared", "-o", LIB, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu link failed:\n" + r.stderr)
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
