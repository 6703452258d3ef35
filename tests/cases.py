"""Mask cases shared by the emulator (CPU) and GPU parity tests: adversarial shapes from
SURVEY.md §4 — spirals, combs, checkerboards, diagonals (4- vs 8-connectivity), all-fg, all-bg,
1×N, N×1, and sizes either side of the 64×128 tile and 32-bit word boundaries."""
import numpy as np

from cellseg_b200 import synth


def diag_mask(H, W, anti=False):
    m = np.zeros((H, W), dtype=bool)
    n = min(H, W)
    idx = np.arange(n)
    m[idx, (W - 1 - idx) if anti else idx] = True
    return m


def holes_mask(H, W, seed=0):
    """Foreground slab with random small/large holes and specks, many straddling tile borders."""
    rng = np.random.default_rng(seed)
    m = np.ones((H, W), dtype=bool)
    for _ in range(max(8, H * W // 600)):
        h, w = rng.integers(1, 14, 2)
        y, x = rng.integers(0, max(1, H - h)), rng.integers(0, max(1, W - w))
        m[y:y + h, x:x + w] = False
    for _ in range(max(8, H * W // 900)):
        h, w = rng.integers(1, 7, 2)
        y, x = rng.integers(0, max(1, H - h)), rng.integers(0, max(1, W - w))
        m[y:y + h, x:x + w] ^= True
    return m


def border_specks(H, W):
    """Small objects/holes placed exactly across tile corners (64,128) and word boundaries."""
    m = np.zeros((H, W), dtype=bool)
    for (y, x) in [(62, 126), (63, 127), (64, 128), (60, 30), (63, 31), (64, 0), (0, 127), (126, 254)]:
        for (dy, dx) in [(0, 0), (0, 1), (1, 0), (1, 1), (2, 1), (1, 2), (0, 2), (2, 0), (2, 2), (3, 0), (0, 3), (3, 3), (3, 1), (3, 2), (1, 3), (2, 3)]:
            yy, xx = y + dy, x + dx
            if 0 <= yy < H and 0 <= xx < W:
                m[yy, xx] = True
    # an L-shaped 15-pixel object crossing a tile corner (must be removed at min_size=16) and a 16-pixel one (kept)
    if H > 140 and W > 270:
        m[120:128, 255] = True; m[127, 255:262] = True          # 8 + 7 - 1 = 14 → + 1
        m[119, 255] = True                                        # 15
        m[100:108, 200] = True; m[107, 200:209] = True          # 8 + 9 - 1 = 16
    return m


def all_cases():
    c = []
    c.append(("noise50_70x150", synth.noise_mask(70, 150, 0.5, 1)))
    c.append(("noise30_257x129", synth.noise_mask(257, 129, 0.3, 2)))
    c.append(("noise45_200x300", synth.noise_mask(200, 300, 0.45, 3)))
    c.append(("noise60_130x260", synth.noise_mask(130, 260, 0.6, 4)))
    c.append(("noise59_64x128", synth.noise_mask(64, 128, 0.59, 5)))
    c.append(("noise59_65x129", synth.noise_mask(65, 129, 0.59, 6)))
    c.append(("noise59_63x127", synth.noise_mask(63, 127, 0.59, 7)))
    c.append(("noise80_192x384", synth.noise_mask(192, 384, 0.8, 8)))
    c.append(("blob_300x500", synth.blob_mask(300, 500, 9)))
    c.append(("blob_fine_260x390", synth.blob_mask(260, 390, 10, sigma=2.0)))
    c.append(("spiral_150x270", synth.spiral_mask(150, 270)))
    c.append(("spiral3_129x257", synth.spiral_mask(129, 257, gap=3)))
    c.append(("comb_v_140x260", synth.comb_mask(140, 260, True)))
    c.append(("comb_h_140x260", synth.comb_mask(140, 260, False)))
    c.append(("checker_131x259", synth.checkerboard(131, 259)))
    c.append(("diag_200x200", diag_mask(200, 200)))
    c.append(("antidiag_200x290", diag_mask(200, 290, True)))
    c.append(("x_200x200", diag_mask(200, 200) | diag_mask(200, 200, True)))
    c.append(("all_fg_100x200", np.ones((100, 200), dtype=bool)))
    c.append(("all_bg_100x200", np.zeros((100, 200), dtype=bool)))
    c.append(("row_1x500", synth.noise_mask(1, 500, 0.6, 11)))
    c.append(("col_500x1", synth.noise_mask(500, 1, 0.6, 12)))
    c.append(("px_1x1_fg", np.ones((1, 1), dtype=bool)))
    c.append(("px_1x1_bg", np.zeros((1, 1), dtype=bool)))
    c.append(("holes_190x390", holes_mask(190, 390, 13)))
    c.append(("holes_inv_190x390", ~holes_mask(190, 390, 14)))
    c.append(("border_specks_200x300", border_specks(200, 300)))
    c.append(("border_specks_inv", ~border_specks(200, 300)))
    c.append(("odd_width_97x203", synth.noise_mask(97, 203, 0.55, 15)))
    c.append(("frame_130x258", _frame(130, 258)))
    # sparse fields: most 64x128 tiles are empty (the fast path of k_tile), next to tiles with pieces touching the shared borders
    c.append(("sparse_cells_400x700", sparse_cells(400, 700, 16)))
    c.append(("sparse_inv_border_260x520", sparse_border_pieces(260, 520)))
    c.append(("all_bg_slivers_134x262", np.zeros((134, 262), dtype=bool)))      # edge tiles of 6 rows / 6 columns, corner tile 6x6 < 64 px
    c.append(("one_px_far_corner_134x262", _one_px(134, 262)))
    return c


def sparse_cells(H, W, seed):
    """A few disks and specks on an empty field; some sit on tile borders / corners."""
    rng = np.random.default_rng(seed)
    m = np.zeros((H, W), dtype=bool)
    yy, xx = np.mgrid[0:H, 0:W]
    for cy, cx, r in [(64, 128, 9), (200, 383, 6), (10, 10, 3), (330, 640, 12), (127, 255, 2), (192, 10, 7)]:
        m |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
    for _ in range(12):
        y, x = rng.integers(0, H - 3), rng.integers(0, W - 3)
        m[y:y + rng.integers(1, 4), x:x + rng.integers(1, 4)] = True
    return m


def sparse_border_pieces(H, W):
    """Foreground walls that cut small background pockets (< 64 px) and large ones right at the borders of otherwise empty tiles."""
    m = np.zeros((H, W), dtype=bool)
    m[60:70, 120:124] = True; m[60, 120:136] = True; m[69, 120:136] = True; m[60:70, 135] = True      # closed box around a 8x11 pocket across (64,128)
    m[126:130, 250:262] = True                                                                        # bar across a tile corner
    m[191, 0:40] = True; m[185:192, 39] = True; m[185, 0:40] = True                                   # pocket of 5x39 = 195 px at the image's left edge
    m[128:131, 300] = True
    return m


def _one_px(H, W):
    m = np.zeros((H, W), dtype=bool)
    m[H - 1, W - 1] = True
    return m


def _frame(H, W):
    m = np.zeros((H, W), dtype=bool)
    m[0, :] = m[-1, :] = True
    m[:, 0] = m[:, -1] = True
    m[H // 2, :] = True
    return m
