"""Where does the tile kernel spend its time?  Builds libcellseg with -DCS_PHASE_TIMING (clock64 marks by thread 0 of every
CTA, ccl.cu) into a scratch .so, runs the mask -> instance-map chain on the mask families of pp_bench.py and prints the
cycles per phase (summed over CTAs, so shares — not wall time).  Usage (under gpurun): python tools/phase_timing.py"""
import sys, os, subprocess, ctypes as C, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cellseg_b200 import build as B, capi, synth

def build_timing():
    out = os.path.join(B.PKG_DIR, "build", os.environ.get("CS_TIMING_LIB", "libcellseg_timing.so"))
    objs = []
    os.makedirs(os.path.join(B.PKG_DIR, "build", "tobj"), exist_ok=True)
    procs = []
    for src in B.sources():
        obj = os.path.join(B.PKG_DIR, "build", "tobj", os.path.basename(src)[:-3] + ".o")
        procs.append(subprocess.Popen([B.nvcc_path(), *[f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")], "-DCS_PHASE_TIMING", *os.environ.get("CS_EXTRA_DEFS", "").split(), "-I", B.CSRC, "-c", src, "-o", obj]))
        objs.append(obj)
    for p in procs:
        assert p.wait() == 0
    subprocess.check_call([B.nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs])
    return out

NAMES = {0: "load", 1: "empty-tile exit", 2: "resume apply",
         3: "s1 store+count", 4: "s1 all-certified publish", 5: "s1 links+ccl", 6: "s1 nodes+border", 7: "s1 apply", 8: "s1 stop: cache",
         10: "s2 store+count", 11: "s2 all-certified publish", 12: "s2 links+ccl", 13: "s2 nodes+border", 14: "s2 apply", 15: "s2 stop: cache",
         17: "s3 ccl/diag", 18: "s3 roots+cache", 19: "s3 border"}
COUNTS = {20: "s1 tiles all certified", 21: "s1 tiles with a rest", 22: "s2 tiles all certified", 23: "s2 tiles with a rest", 24: "s1 stops", 25: "s2 stops"}
def main():
    cand = os.path.join(B.PKG_DIR, "build", os.environ.get("CS_TIMING_LIB", "libcellseg_timing.so"))
    path = cand if (os.path.exists(cand) and "--rebuild" not in sys.argv and torch.cuda.is_available()) else build_timing()
    if not torch.cuda.is_available():
        print("built", path); return
    capi._lib = capi.bind(C.CDLL(path))
    from cellseg_b200 import postprocess
    lib = capi._lib
    buf = (C.c_ulonglong * 128)()
    cases = {
        "blob 64x1024^2": np.stack([synth.blob_mask(1024, 1024, s) for s in range(64)]),
        "sparse cells 64x1024^2": np.stack([synth.lattice_cells(1024, 1024, seed=s, pitch=200, radius=14) > 0 for s in range(64)]),
    }
    um = [os.path.join(ROOT, "tests", "golden", "unet_masks_4x1024.npz")]
    if os.path.exists(um[0]):
        m4 = np.unpackbits(np.load(um[0])["packed"]).reshape(4, 1024, 1024).astype(bool)
        cases["unet-like 64x1024^2"] = np.tile(m4, (16, 1, 1))
    for name, m in cases.items():
        d = torch.from_numpy(np.ascontiguousarray(m)).cuda()
        out = torch.empty(d.shape, dtype=torch.int32, device="cuda")
        for _ in range(2): postprocess.postprocess_device(d, out=out)
        torch.cuda.synchronize()
        lib.cs_debug_phase_cycles(buf, 1)
        postprocess.postprocess_device(d, out=out)
        torch.cuda.synchronize()
        lib.cs_debug_phase_cycles(buf, 1)
        tot = sum(buf[s * 40 + i] for s in range(3) for i in range(20))
        print(f"== {name}: total {tot / 1e6:.1f} Mcycles (sum over CTAs), {m.shape[0] * ((m.shape[1] + 63) // 64) * ((m.shape[2] + 127) // 128)} tiles")
        for start in range(3):
            sub = sum(buf[start * 40: start * 40 + 20])
            if not sub: continue
            print(f"  k_tile<{start}>: {sub / tot * 100:.1f} % of all")
            for i in range(20):
                v = buf[start * 40 + i]
                if v: print(f"     {NAMES.get(i, i):26s} {v / tot * 100:6.2f} %")
            for i in range(20, 30):
                v = buf[start * 40 + i]
                if v: print(f"     # {COUNTS.get(i, i):24s} {v}")

if __name__ == "__main__":
    main()
