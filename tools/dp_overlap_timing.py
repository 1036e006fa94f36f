import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import vit3d_b200
from oracle import vit3d_oracle as O
from vit3d_b200 import functional as F
from vit3d_b200.optim import FusedSGD
from vit3d_b200.graphs import GraphedTrainStep
from vit3d_b200.models.modeling import VisionTransformer
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
cfg = vit3d_b200.north_star_config(18)
B = 256
x = torch.randn(B, 1, 128, 128, 5, device=dev); y = (torch.rand(B, device=dev) > 0.5).float()
for overlap in (False, True, False, True):
    m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision="bf16").to(dev); m.train()
    opt = FusedSGD(m.parameters(), lr=0.01, momentum=0.9)
    step = GraphedTrainStep(m, opt, warmup=2, data_parallel=True, overlap=overlap)
    for _ in range(5): step(x, y, 1.0)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): step(x, y, 1.0)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 30], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"world {world} overlap={overlap}: {float(t):.3f} ms/step, segmented={step._graphs is not None}", flush=True)
    del step, opt, m
dist.destroy_process_group()
