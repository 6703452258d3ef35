"""CPU tier for the strip-sharded (multi-GPU) post-processing: the cs_strip_* phases on the fiber-emulated kernels.

  1. lockstep: n strips of one image driven phase by phase in one process == the oracle on the whole image;
  2. world_size-2/3 `gloo` processes, one strip each, rows exchanged with torch.distributed — the host logic the
     GPU ranks run over NCCL (cellseg_b200.wsi.run_distributed).
The product module is pointed at the emulated library by monkeypatching capi._lib — test-only."""
import os
import sys

import numpy as np
import pytest
import torch

import emu_lib
from cases import all_cases, holes_mask
from oracle import morph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def wsi(monkeypatch):
    from cellseg_b200 import capi, wsi
    monkeypatch.setattr(capi, "_lib", emu_lib.lib())
    return wsi


def _snake(H, W):
    """One 1-pixel path that crosses every 64-row cut many times, plus small specks sitting on the cuts."""
    from cellseg_b200 import synth
    m = synth.comb_mask(H, W, vertical=True)          # teeth run through all strips, joined at the top
    m[1:, ::4] = False
    for y in range(63, H - 1, 64):                     # 2x3 and 4x5 specks / 3x3 and 9x9 holes across the cuts
        m[y - 1:y + 3, 40:46] = False
        m[y:y + 2, 41:44] = True
        m[y - 1:y + 3, 60:68] = False
        m[y - 1:y + 3, 61:66] = True
        m[y - 4:y + 5, 100:120] = True
        m[y - 1:y + 2, 104:107] = False
        m[y - 4:y + 5, 109:118] = False
    return m


def _run(wsi, mask, n):
    H, W = mask.shape
    rows = wsi.strip_rows(H, n)
    strips = [wsi.StripPost(r1 - r0, W, "cpu") for r0, r1 in rows]
    masks = [torch.from_numpy(np.ascontiguousarray(mask[r0:r1])) for r0, r1 in rows]
    labels, total = wsi.run_lockstep(strips, masks)
    return np.concatenate([l.numpy() for l in labels]), total


CASES = [c for c in all_cases() if c[1].shape[0] >= 129] + [("snake_330x260", _snake(330, 260)), ("holes_400x200", holes_mask(400, 200, 3)),
                                                            ("holes_inv_400x200", ~holes_mask(400, 200, 4))]


@pytest.mark.parametrize("name,mask", CASES, ids=[c[0] for c in CASES])
def test_lockstep_strips_equal_whole_image(wsi, name, mask):
    ref = morph.postprocess_mask(mask)
    nty = -(-mask.shape[0] // 64)
    for n in sorted({2, 3, nty}):
        if n > nty:
            continue
        got, total = _run(wsi, mask, n)
        assert np.array_equal(got, ref), (name, n)
        assert total == ref.max()


def test_strip_rows_partition():
    from cellseg_b200 import wsi
    for H, n in [(10000, 8), (10000, 2), (4096, 4), (130, 8), (64, 3), (1, 1)]:
        rows = wsi.strip_rows(H, n)
        assert rows[0][0] == 0 and rows[-1][1] == H
        assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        assert all(r0 % 64 == 0 and r1 > r0 for r0, r1 in rows)
        assert len(rows) == min(n, -(-H // 64))


def test_strip_argument_errors(wsi):
    s = wsi.StripPost(100, 50, "cpu")                          # 100 rows: not a multiple of 64
    s.begin(torch.zeros((100, 50), dtype=torch.uint8))
    rows = s.filter_export(1)
    from cellseg_b200 import capi
    with pytest.raises(capi.CellSegError, match="multiple of 64"):
        s.filter_resume(1, None, rows[0])


def _worker(rank, world, port, H, W, seed, q):
    try:
        sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import torch.distributed as dist
        import cellseg_b200  # noqa: F401
        import emu_lib as el
        from cellseg_b200 import capi, synth, wsi
        capi._lib = el.lib()
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        mask = synth.blob_mask(H, W, seed, sigma=3.0) ^ synth.noise_mask(H, W, 0.02, seed + 1)
        r0, r1 = wsi.strip_rows(H, world)[rank]
        strip = wsi.StripPost(r1 - r0, W, "cpu")
        labels, total = wsi.run_distributed(strip, torch.from_numpy(np.ascontiguousarray(mask[r0:r1])), rank, world)
        from oracle import morph as om
        ref = om.postprocess_mask(mask)
        ok = bool(np.array_equal(labels.numpy(), ref[r0:r1])) and total == int(ref.max())
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, ok, total))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks_exchange_rows(world):
    import torch.multiprocessing as mp
    emu_lib.lib()                                               # build once in the parent
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, 320, 384, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    try:
        for _ in procs:
            res.append(q.get(timeout=240))
    except Exception:                                            # a rank that never reports: show what the others said and how they ended
        for p in procs:
            p.join(timeout=5)
        codes = [p.exitcode for p in procs]
        for p in procs:
            if p.is_alive():
                p.terminate()
        raise AssertionError(f"world {world}: only {len(res)} of {len(procs)} ranks reported; exit codes {codes}; reports: {res}")
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


@pytest.mark.parametrize("H,W,world", [(10000, 10000, 8), (10000, 10000, 2), (4096, 3000, 4), (1024, 1024, 3), (700, 900, 2)])
def test_window_plan_partitions_and_transfers(wsi, H, W, world):
    """Every window is predicted exactly once, shares are equal to within one window, and the transfers are exactly what each
    strip needs beyond what it owns."""
    from cellseg_b200.sliding_window import make_plan
    plan = make_plan(H, W, 256, 0.25)
    strips, owned, needed, transfers = wsi.window_plan(H, W, world)
    n = plan.ny * plan.nx
    assert owned[0][0] == 0 and owned[-1][1] == n and all(owned[i][1] == owned[i + 1][0] for i in range(world - 1))
    sizes = [b - a for a, b in owned]
    assert max(sizes) - min(sizes) <= 1
    for q, ((r0, r1), (lo, hi)) in enumerate(zip(strips, needed)):
        rows = {i for i, y in enumerate(plan.ys) if y < r1 and y + 256 > r0}          # window rows that intersect the strip
        assert set(range(lo, hi)) == {iy * plan.nx + ix for iy in rows for ix in range(plan.nx)}
        got = set(range(max(lo, owned[q][0]), min(hi, owned[q][1])))
        for o, qq, a, b in transfers:
            if qq == q:
                assert o != q and owned[o][0] <= a < b <= owned[o][1] and not (got & set(range(a, b)))
                got |= set(range(a, b))
        assert got == set(range(lo, hi))
