"""Run under torchrun (one rank per GPU): strip-sharded whole-slide image → instance map over NCCL, checked against
the single-GPU engine on rank 0.  Usage: torchrun --nproc-per-node N tools/wsi_check.py [H] [W]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import cellseg_b200
from cellseg_b200 import sliding_window as sw, synth, wsi
from cellseg_b200.unet import build_unet

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.deterministic = True
model = build_unet(3, seed=2022).eval().cuda()
img = synth.modality_image(H, W, seed=0, modality=0)
seg = wsi.StripSegmenter(model, H, W, rank, world, chunk=1, cuda_graph=False)
s0, s1 = seg.source_rows()
labels, total = seg(torch.from_numpy(np.ascontiguousarray(img[s0:s1])).cuda())
torch.cuda.synchronize()
ref, cnt = sw.SegmentationEngine(model, chunk=1)(torch.from_numpy(img[None]).cuda())
ok = bool((labels == ref[0, seg.r0:seg.r1]).all()) and total == int(cnt[0])
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"wsi_check {H}x{W} over {world} ranks: total={total} single-GPU={int(cnt[0])} bit-identical={bool(flag.item())}", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
