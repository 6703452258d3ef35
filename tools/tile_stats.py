"""CPU-side statistics of the post-processing workload per 64x128 tile (scipy): how many tiles hold border pieces below the
size thresholds (the tiles k_tile has to stop and resume), segments per word, components per tile — for the three mask
families of tools/pp_bench.py.  Usage: python tools/tile_stats.py [unet]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.ndimage as ndi
from cellseg_b200 import synth
from oracle import morph

TH, TW = 64, 128
C4 = ndi.generate_binary_structure(2, 1)

def border_small(tile, thr):
    """number of 4-conn components of `tile` that touch its border and have < thr pixels; and total components"""
    cc, n = ndi.label(tile, C4)
    if n == 0: return 0, 0
    sizes = np.bincount(cc.ravel(), minlength=n + 1)
    b = np.unique(np.concatenate([cc[0], cc[-1], cc[:, 0], cc[:, -1]]))
    b = b[b > 0]
    return int((sizes[b] < thr).sum()), n

def stats(name, m):
    H, W = m.shape
    m1 = morph.remove_small_holes(m, 64)
    m2 = morph.remove_small_objects(m1, 16)
    nt = s1 = s2 = e = 0; u1 = u2 = 0; c1 = c2 = 0
    for y in range(0, H, TH):
        for x in range(0, W, TW):
            nt += 1
            t0 = m[y:y + TH, x:x + TW]
            if not t0.any(): e += 1; continue
            a, n = border_small(~t0, 64); s1 += a > 0; u1 += a; c1 += n
            a, n = border_small(m1[y:y + TH, x:x + TW], 16); s2 += a > 0; u2 += a; c2 += n
    w = np.packbits(m.reshape(H, W // 32, 32), axis=-1, bitorder="little").view(np.uint32)[..., 0] if W % 32 == 0 else None
    seg = ((m[:, 1:] & ~m[:, :-1]).sum() + m[:, 0].sum()) / (H * W / 32)
    print(f"{name}: fg {m.mean():.3f} → {m1.mean():.3f} → {m2.mean():.3f}; tiles {nt}, empty {e / nt:.2f}, stop@holes {s1 / nt:.2f} "
          f"(undecided/tile {u1 / nt:.1f}, bg comps/tile {c1 / nt:.1f}), stop@objects {s2 / nt:.2f} (undecided/tile {u2 / nt:.1f}, "
          f"fg comps/tile {c2 / nt:.1f}); fg segments/word {seg:.2f}; instances {morph.label(m2).max()}")

if __name__ == "__main__":
    stats("blob 1024^2", synth.blob_mask(1024, 1024, 0))
    stats("sparse cells 1024^2", synth.lattice_cells(1024, 1024, seed=0, pitch=200, radius=14) > 0)
    stats("noise p=0.45", synth.noise_mask(1024, 1024, 0.45, 1))
    if "unet" in sys.argv:
        import torch, bench
        from cellseg_b200.unet import build_unet
        from oracle import sliding_window as osw
        net = build_unet(3, seed=2022).eval()
        imgs = bench.synthetic_batch(4, 1024, 0)
        for i in range(4):
            x = osw.scale_to_unit(imgs[i])
            with torch.no_grad():
                out = osw.sliding_window_inference(x, (256, 256), 4, net)
            m = osw.foreground_mask(out)
            np.save(f"gpurun_out/unet_mask_{i}.npy", m)
            stats(f"unet mask {i} (modality {i % 4})", m)
