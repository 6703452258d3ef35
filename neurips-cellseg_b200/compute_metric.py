"""Drop-in for the reference's baseline/compute_metric.py: same function names, argument meaning,
return values, CLI flags and CSV layout; the per-pixel work runs in libcellseg_sm100.so.

  remove_boundary_cells  (:124-133)  → K7: border/present flags per label + prefix-count LUT + one remap pass
  _label_overlap         (:39-70)    → K6: sparse (gt, pred) pair histogram, dense uint64 matrix on request
  _intersection_over_union (:21-37)  → from the sparse triples and their row/column sums (float64 on device)
  _true_positive         (:83-105)   → host: the Hungarian objective there is lexicographic (count of IoU ≥ th
                                        pairs first, since the IoU tie-break sums to < 1), i.e. a maximum-cardinality
                                        matching of the thresholded pair graph — solved on the sparse graph
  eval_tp_fp_fn          (:107-122), dice (:72-81), main (:135-265) incl. the ≥ 25 Mpx 2000² patch branch.
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import traceback
from collections import OrderedDict

import numpy as np

from . import capi

join = os.path.join


# ---------------------------------------------------------------------------
# device helpers (torch is only the allocator / stream provider)
# ---------------------------------------------------------------------------
def _torch():
    import torch
    if not torch.cuda.is_available():
        raise capi.CellSegError("compute_metric needs a CUDA device; cellseg_b200 has no CPU path")
    return torch


def _stream(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _to_device_i32(a):
    torch = _torch()
    if isinstance(a, torch.Tensor):
        return a.to("cuda", torch.int32)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda()


def _check_same_shape(a, b, what):
    """The kernels walk both maps with one geometry: a mismatched pair must fail here, as a Python error the CLI's per-file
    try/except records as `failed` (the reference fails the same file: numpy's boolean index in dice(), compute_metric.py:72-81,
    raises on the mismatch) — never as an out-of-bounds read on the device."""
    sa, sb = tuple(np.shape(a)), tuple(np.shape(b))
    if sa != sb:
        raise ValueError(f"{what}: ground truth {sa} and segmentation {sb} differ in shape")


def _label_max_dev(t):
    """(max, min) label of a 2-D int32 CUDA tensor (rows may be strided)."""
    torch = _torch()
    lib = capi.load()
    out = torch.empty(2, dtype=torch.int32, device=t.device)
    capi.check(lib, lib.cs_label_max(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), out.data_ptr(), _stream(torch)))
    mx, mn = out.tolist()
    if mn < 0:
        raise ValueError("negative labels are not supported")
    return mx, mn


def _relabel_dev(t, border: int, zero_inplace: bool):
    """t: 2-D int32 CUDA tensor with unit column stride → (relabelled contiguous tensor, n_labels)."""
    torch = _torch()
    lib = capi.load()
    assert t.dim() == 2 and t.stride(1) == 1 and t.dtype == torch.int32
    H, W = t.shape
    mx, _ = _label_max_dev(t)
    wsb = lib.cs_relabel_workspace_bytes(mx)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=t.device)
    out = torch.empty((H, W), dtype=torch.int32, device=t.device)
    n = torch.empty(1, dtype=torch.int32, device=t.device)
    capi.check(lib, lib.cs_relabel(t.data_ptr(), H, W, t.stride(0), border, mx, int(zero_inplace), out.data_ptr(), W,
                                   n.data_ptr(), ws.data_ptr(), wsb, _stream(torch)))
    return out, n


def _sparse_overlap_dev(gt, pr, nt: int, npred: int):
    """gt, pr: 2-D int32 CUDA tensors (unit column stride).  → dict of device tensors: triples (pg, pp, pc),
    their count, row/col sums."""
    torch = _torch()
    lib = capi.load()
    H, W = gt.shape
    dev = gt.device
    cap = 1 << 16
    while cap < 8 * (nt + npred + 2) and cap < (1 << 30):
        cap <<= 1
    while True:
        wsb = lib.cs_overlap_workspace_bytes(cap)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        pg = torch.empty(cap, dtype=torch.int32, device=dev)
        pp = torch.empty(cap, dtype=torch.int32, device=dev)
        pc = torch.empty(cap, dtype=torch.int32, device=dev)        # uint32 payload
        misc = torch.empty(2, dtype=torch.int32, device=dev)        # n_pairs, status
        ag = torch.empty(nt + 1, dtype=torch.int32, device=dev)
        ap = torch.empty(npred + 1, dtype=torch.int32, device=dev)
        capi.check(lib, lib.cs_label_overlap_sparse(
            gt.data_ptr(), gt.stride(0), pr.data_ptr(), pr.stride(0), H, W, nt, npred, pg.data_ptr(), pp.data_ptr(), pc.data_ptr(),
            cap, misc.data_ptr(), ag.data_ptr(), ap.data_ptr(), misc.data_ptr() + 4, cap, ws.data_ptr(), wsb, _stream(torch)))
        n_pairs, status = misc.tolist()
        if status & 6:
            raise ValueError("labels out of range (negative, or above the maximum passed)")
        if status & 1:
            if cap >= (1 << 30):
                raise capi.CellSegError("overlap table overflow")
            cap <<= 2
            continue
        return dict(pg=pg, pp=pp, pc=pc, n=n_pairs, cap=cap, misc=misc, ag=ag, ap=ap)


def _max_matching(rows: np.ndarray, cols: np.ndarray) -> int:
    """Size of a maximum-cardinality matching of the bipartite graph with the given edges."""
    if len(rows) == 0:
        return 0
    ur, ri = np.unique(rows, return_inverse=True)
    uc, ci = np.unique(cols, return_inverse=True)
    if len(ur) == len(rows) and len(uc) == len(cols):
        return len(rows)                                   # every row and column has one edge: the usual case at th >= 0.5
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import maximum_bipartite_matching
    g = csr_matrix((np.ones(len(rows), dtype=np.int8), (ri, ci)), shape=(len(ur), len(uc)))
    return int((maximum_bipartite_matching(g, perm_type="column") >= 0).sum())


def _tp_from_sparse(sp, nt: int, npred: int, th: float) -> int:
    """tp of compute_metric.py:83-105 from the sparse triples: thresholding and the one-edge-per-row/column check run on the device
    (cs_iou_edges); the host only sees two integers unless a row or column really has several candidates."""
    if th <= 0:                                            # every pair (incl. non-overlapping ones) is an edge
        return min(nt, npred)
    torch = _torch()
    lib = capi.load()
    dev = sp["pg"].device
    scratch = torch.empty(nt + npred + 2, dtype=torch.int32, device=dev)
    eg = torch.empty(sp["cap"], dtype=torch.int32, device=dev)
    ep = torch.empty(sp["cap"], dtype=torch.int32, device=dev)
    out = torch.empty(2, dtype=torch.int32, device=dev)
    capi.check(lib, lib.cs_iou_edges(sp["pg"].data_ptr(), sp["pp"].data_ptr(), sp["pc"].data_ptr(), sp["misc"].data_ptr(), sp["cap"], nt, npred,
                                     sp["ag"].data_ptr(), sp["ap"].data_ptr(), float(th), scratch.data_ptr(), scratch.data_ptr() + 4 * (nt + 1),
                                     eg.data_ptr(), ep.data_ptr(), out.data_ptr(), _stream(torch)))
    n_edges, multi = out.tolist()
    if not multi:
        return n_edges                                     # every row and column has at most one edge: all of them are matched
    return _max_matching(eg[:n_edges].cpu().numpy(), ep[:n_edges].cpu().numpy())


# ---------------------------------------------------------------------------
# reference-named functions
# ---------------------------------------------------------------------------
def _label_overlap(x, y):
    """Dense uint64 overlap matrix [(1+x.max()), (1+y.max())] (compute_metric.py:39-70)."""
    lib = capi.load()
    x = np.ascontiguousarray(np.asarray(x).ravel(), dtype=np.int32)
    y = np.ascontiguousarray(np.asarray(y).ravel(), dtype=np.int32)
    _check_same_shape(x, y, "_label_overlap")
    nt, npred = int(x.max()), int(y.max())
    ov = np.empty((nt + 1, npred + 1), dtype=np.uint64)
    capi.check(lib, lib.cs_label_overlap_host(x.ctypes.data, y.ctypes.data, 1, x.size, nt, npred, ov.ctypes.data))
    return ov.astype(np.uint, copy=False)


def _intersection_over_union(masks_true, masks_pred):
    """Dense float64 IoU matrix incl. the background row/column (compute_metric.py:21-37)."""
    torch = _torch()
    lib = capi.load()
    _check_same_shape(masks_true, masks_pred, "_intersection_over_union")
    gt, pr = _to_device_i32(masks_true), _to_device_i32(masks_pred)
    gt, pr = gt.reshape(1, -1), pr.reshape(1, -1)
    nt, _ = _label_max_dev(gt)
    npred, _ = _label_max_dev(pr)
    sp = _sparse_overlap_dev(gt, pr, nt, npred)
    iou = torch.empty((nt + 1, npred + 1), dtype=torch.float64, device=gt.device)
    capi.check(lib, lib.cs_iou_dense(sp["pg"].data_ptr(), sp["pp"].data_ptr(), sp["pc"].data_ptr(), sp["misc"].data_ptr(), sp["cap"],
                                     nt, npred, sp["ag"].data_ptr(), sp["ap"].data_ptr(), iou.data_ptr(), _stream(torch)))
    return iou.cpu().numpy()


def dice(gt, seg):
    """compute_metric.py:72-81 on boolean (or any non-zero = foreground) arrays."""
    lib = capi.load()
    g = np.ascontiguousarray(np.asarray(gt) != 0).view(np.uint8)
    s = np.ascontiguousarray(np.asarray(seg) != 0).view(np.uint8)
    out = np.zeros(3, dtype=np.uint64)
    if g.shape != s.shape:
        # the reference only pairs the two arrays (seg[gt]) once it knows that the ground truth is not empty
        ng = ns = 0
        for arr in (g, s):
            capi.check(lib, lib.cs_count_nonzero3_host(arr.ctypes.data, arr.ctypes.data, arr.size, out.ctypes.data))
            ng, ns = ns, int(out[0])
        if ng == 0:
            return 1.0 if ns == 0 else 0.0
        _check_same_shape(g, s, "dice")
    capi.check(lib, lib.cs_count_nonzero3_host(g.ctypes.data, s.ctypes.data, g.size, out.ctypes.data))
    ng, ns, nb = (int(v) for v in out)
    if ng == 0 and ns == 0:
        return 1.0
    if ng == 0 and ns > 0:
        return 0.0
    return 2 * nb / (ng + ns)


def _true_positive(iou, th):
    """compute_metric.py:83-105 on a dense IoU matrix (background row/col already dropped by the caller)."""
    iou = np.asarray(iou)
    if th <= 0:
        return min(iou.shape)
    r, c = np.nonzero(iou >= th)
    return _max_matching(r, c)


def eval_tp_fp_fn(masks_true, masks_pred, threshold=0.5):
    """compute_metric.py:107-122.  Returns (tp, fp, fn); an empty prediction gives (0, 0, 0) like the reference (which does not
    look at the ground truth in that case, whatever its shape)."""
    pr = _to_device_i32(masks_pred)
    if pr.dim() != 2:
        pr = pr.reshape(1, -1)
    num_inst_seg, _ = _label_max_dev(pr)
    if num_inst_seg <= 0:
        return 0, 0, 0
    _check_same_shape(masks_true, masks_pred, "eval_tp_fp_fn")
    gt = _to_device_i32(masks_true)
    if gt.dim() != 2:
        gt = gt.reshape(1, -1)
    num_inst_gt, _ = _label_max_dev(gt)
    sp = _sparse_overlap_dev(gt, pr, num_inst_gt, num_inst_seg)
    tp = _tp_from_sparse(sp, num_inst_gt, num_inst_seg, threshold)
    return tp, num_inst_seg - tp, num_inst_gt - tp


def remove_boundary_cells(mask):
    """compute_metric.py:124-133: cells touching the 2-pixel frame are removed — IN PLACE in `mask`, as in the
    reference — and the survivors are relabelled 1..n in ascending order (returned as a new array)."""
    lib = capi.load()
    work = np.ascontiguousarray(mask, dtype=np.int32)      # `mask` itself when it already is contiguous int32
    out = np.empty(work.shape, dtype=np.int32)
    n = np.zeros(1, dtype=np.int32)
    capi.check(lib, lib.cs_remove_boundary_cells_host(work.ctypes.data, work.shape[0], work.shape[1], out.ctypes.data, n.ctypes.data))
    if work is not mask:
        mask[...] = work                                   # propagate the mutation to the caller's array / view
    if mask.dtype != np.int32 and int(n[0]) <= np.iinfo(mask.dtype).max:
        return out.astype(mask.dtype)                      # relabel_sequential keeps the input dtype when it fits
    return out


def relabel_sequential(label_field):
    """skimage.segmentation.relabel_sequential(label_field)[0] (compute_metric.py:186-187)."""
    lib = capi.load()
    src = np.ascontiguousarray(label_field, dtype=np.int32)
    out = np.empty(src.shape, dtype=np.int32)
    capi.check(lib, lib.cs_relabel_sequential_host(src.ctypes.data, src.shape[0], src.shape[1], out.ctypes.data, None))
    return out


# ---------------------------------------------------------------------------
# one (gt, seg) pair, device-resident (what main() uses)
# ---------------------------------------------------------------------------
def evaluate_pair(gt, seg, threshold=0.5, count_bd_cells=False, roi_size=2000):
    """compute_metric.py:179-241 for one pair of instance maps → dict(true_num, pred_num, tp, fp, fn, dice)."""
    torch = _torch()
    lib = capi.load()
    if len(np.shape(gt)) != 2:
        raise ValueError(f"evaluate_pair: instance maps must be 2-D, got {tuple(np.shape(gt))}")
    _check_same_shape(gt, seg, "evaluate_pair")
    g, s = _to_device_i32(gt), _to_device_i32(seg)
    H, W = g.shape
    cnt = torch.zeros(3, dtype=torch.int64, device=g.device)
    gb, sb = (g != 0).view(torch.uint8), (s != 0).view(torch.uint8)
    capi.check(lib, lib.cs_count_nonzero3(gb.data_ptr(), sb.data_ptr(), gb.numel(), cnt.data_ptr(), _stream(torch)))
    ng, ns, nb = cnt.tolist()
    dice_score = 1.0 if (ng == 0 and ns == 0) else 0.0 if ng == 0 else 2 * nb / (ng + ns)

    def score(gt_t, seg_t):
        if not count_bd_cells:
            # remove_boundary_cells ends in relabel_sequential, so the reference's second relabel_sequential (:186-187) is the identity
            # on its output: one pass gives both the map and the count
            gt_t, n_true = _relabel_dev(gt_t, 2, False)
            seg_t, n_pred = _relabel_dev(seg_t, 2, False)
        else:
            gt_t, n_true = _relabel_dev(gt_t, 0, False)
            seg_t, n_pred = _relabel_dev(seg_t, 0, False)
        n_true, n_pred = int(n_true), int(n_pred)
        if n_pred > 0:
            sp = _sparse_overlap_dev(gt_t, seg_t, n_true, n_pred)
            tp = _tp_from_sparse(sp, n_true, n_pred, threshold)
            return n_true, n_pred, tp, n_pred - tp, n_true - tp
        return n_true, n_pred, 0, 0, 0

    if H * W < 25000000:
        n_true, n_pred, tp, fp, fn = score(g, s)
    else:
        if count_bd_cells:
            raise NameError("name 'gt_roi' is not defined")   # the reference's patch branch fails this way (:223-227)
        n_h, n_w = -(-H // roi_size), -(-W // roi_size)
        gp = torch.zeros((n_h * roi_size, n_w * roi_size), dtype=torch.int32, device=g.device)
        sp_ = torch.zeros_like(gp)
        gp[:H, :W] = g; sp_[:H, :W] = s
        n_true = n_pred = tp = fp = fn = 0
        for i in range(n_h):
            for j in range(n_w):
                sl = (slice(roi_size * i, roi_size * (i + 1)), slice(roi_size * j, roi_size * (j + 1)))
                a, b, c, d, e = score(gp[sl], sp_[sl])
                n_true += a; n_pred += b; tp += c; fp += d; fn += e
    return dict(true_num=n_true, pred_num=n_pred, tp=tp, fp=fp, fn=fn, dice=dice_score)


def dist_env():
    """(rank, world, local_rank) of a `torchrun` launch; (0, 1, 0) for a plain process."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


_OWN_GROUP = False        # this module created the process group (release_group() tears it down)


def gather_rows(rows: list, rank: int, world: int) -> list:
    """File-sharded CLIs: every rank evaluated names[rank::world]; returns all (name, ...) rows in name order on every rank.
    Image pairs are independent units, so this is the only exchange of the multi-GPU metric path: a gloo all-gather of a few
    Python tuples per file (no data-path collective).  The process group is created on first use and kept until release_group():
    creating and destroying one per call raced on the rendezvous port (the second group's pairs met the first group's teardown —
    'Connection reset by peer' on one rank, the other waiting for ever)."""
    if world == 1:
        return sorted(rows)
    import torch.distributed as dist
    global _OWN_GROUP
    if not dist.is_initialized():
        import datetime
        # MASTER_ADDR / MASTER_PORT from torchrun; a rank that died before getting here must not hold the others for gloo's 30 minutes
        dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=300))
        _OWN_GROUP = True
    parts = [None] * world
    dist.all_gather_object(parts, rows)
    return sorted(r for part in parts for r in part)


def release_group():
    """Tear down the process group gather_rows created (after a barrier: the rank that hosts the rendezvous store leaves last)."""
    global _OWN_GROUP
    if _OWN_GROUP:
        import torch.distributed as dist
        try:
            dist.barrier()
        finally:
            dist.destroy_process_group()
            _OWN_GROUP = False


def main():
    try:
        _main()
    finally:
        release_group()


def _main():
    from .tiffio import imread
    parser = argparse.ArgumentParser('Compute F1 score for cell segmentation results', add_help=False)
    parser.add_argument('-g', '--gt_path', default='labelsTr_GT', type=str, help='path to ground truth')
    parser.add_argument('-s', '--seg_path', type=str, default='', help='path to segmentation results; file names are the same as ground truth', required=False)
    parser.add_argument('--gt_suffix', default='.tif', type=str, help='suffix of ground truth names')
    parser.add_argument('--seg_suffix', default='_label.tiff', type=str, help='suffix of segmentation names')
    parser.add_argument('-thre', '--thresholds', nargs='+', default=[0.5], type=float, help='threshold to count correct cells')
    parser.add_argument('-o', '--output_path', default='./', type=str, help='path where to save metrics')
    parser.add_argument('-n', '--save_name', default='demo', type=str, help='name of the csv file')
    parser.add_argument('--count_bd_cells', default=False, action='store_true', required=False, help='remove the boundary cells when computing metrics by default')
    args = parser.parse_args()

    import pandas as pd
    names = sorted(os.listdir(args.seg_path))
    names = [i for i in names if i.endswith(args.seg_suffix)]
    # under torchrun the files are dealt out over the ranks (one GPU each); rank 0 writes the CSV
    rank, world, local = dist_env()
    if world > 1:
        t = _torch()
        if t.cuda.is_available():
            t.cuda.set_device(local % t.cuda.device_count())
        names = names[rank::world]
    columns = ('names', 'true_num', 'pred_num', 'correct_num(TP)', 'missed_num(FN)', 'wrong_num(FP)', 'precision', 'recall', 'dice', 'F1')
    for threshold in args.thresholds:
        if rank == 0:
            print('compute metrics at threshold:', threshold)
        metrics = OrderedDict((k, []) for k in columns)
        failed = []
        for name in names:
            try:
                gt_name = name.split(args.seg_suffix)[0] + args.gt_suffix
                gt = imread(join(args.gt_path, gt_name))
                seg = imread(join(args.seg_path, name))
                r = evaluate_pair(gt, seg, threshold, args.count_bd_cells)
                tp = r['tp']
                if tp == 0:
                    precision = recall = f1 = 0
                else:
                    precision = tp / r['pred_num']
                    recall = tp / r['true_num']
                    f1 = 2 * (precision * recall) / (precision + recall)
                metrics['names'].append(name)
                metrics['true_num'].append(r['true_num'])
                metrics['pred_num'].append(r['pred_num'])
                metrics['correct_num(TP)'].append(tp)
                metrics['missed_num(FN)'].append(r['fn'])
                metrics['wrong_num(FP)'].append(r['fp'])
                metrics['precision'].append(np.round(precision, 4))
                metrics['recall'].append(np.round(recall, 4))
                metrics['dice'].append(np.round(r['dice'], 4))
                metrics['F1'].append(np.round(f1, 4))
            except Exception:
                print('!' * 20)
                print(name, 'evaluation error!')
                traceback.print_exc()
                failed.append(name)
        if world > 1:
            # one exchange for both lists: rows are tagged 0, failed names 1
            both = gather_rows([(0,) + tuple(metrics[k][i] for k in columns) for i in range(len(metrics['names']))] + [(1, f) for f in failed],
                               rank, world)
            rows = [r[1:] for r in both if r[0] == 0]
            failed = [r[1] for r in both if r[0] == 1]
            metrics = OrderedDict((k, [r[j] for r in rows]) for j, k in enumerate(columns))
            if rank != 0:
                continue
        seg_metric_df = pd.DataFrame(metrics)
        if args.save_name.endswith('.csv'):
            args.save_name = args.save_name.split('.csv')[0]
        save_name = args.save_name + '-' + str(threshold) + '.csv'
        seg_metric_df.to_csv(join(args.output_path, save_name), index=False)
        print('threshold:', threshold, 'mean F1 Score:', np.mean(metrics['F1']), 'median F1 Score:', np.median(metrics['F1']))
        print('failed cases:', failed)


if __name__ == '__main__':
    main()
