"""Multi-GPU correctness (run under torchrun, one rank per GPU):
  * DP-N gradients (GradReducer + global_pos_weight, in-place arena accumulation) == single-process gradients on the
    concatenated batch (SURVEY.md test plan v);
  * graphed DP steps with the overlapped (two-graph) gradient all-reduce == the single all-reduce schedule;
  * ShardedEnsemble output == TransformerEnsemble output on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import vit3d_b200
from oracle import vit3d_oracle as O
from vit3d_b200 import functional as F
from vit3d_b200.dist import GradReducer, ShardedEnsemble, global_pos_weight
from vit3d_b200.optim import FusedSGD
from vit3d_b200.models.modeling import TransformerEnsemble, VisionTransformer

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
for prec, rtol in (("fp32", 2e-3), ("bf16", 6e-2)):
    cfg = vit3d_b200.get_config(16, 512, 2, 256, 8, dropout_rate=0.0)
    sd = O.init_state_dict(cfg, seed=42)
    B = 4 * world
    x = O.synth_volumes(B, seed=3).to(dev)
    y = O.synth_labels(B).to(dev)
    # single-process reference on the whole batch
    ref = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision=prec).to(dev)
    ref.load_state_dict(sd); ref.train()
    F.enable_direct_grads(False)
    loss_ref = ref(x, y, O.sklearn_pos_weight(y.cpu()))
    loss_ref.backward()
    # DP on the shard of this rank
    m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision=prec).to(dev)
    m.load_state_dict(sd); m.train()
    opt = FusedSGD(m.parameters(), lr=0.0)          # builds the flat arena, turns in-place accumulation on
    red = GradReducer(m, arena=opt.arena)
    xs, ys = x[rank::world], y[rank::world]
    for overlap in (True,):
        red.prepare()
        loss = m(xs, ys, global_pos_weight(ys))
        loss.backward()
        red.finish()
    gmax = max(float(p.grad.norm()) for p in ref.parameters())
    worst = 0.0
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        e = float((p.grad - q.grad).norm()); r = float(q.grad.norm())
        if e > rtol * r + 2e-4 * gmax:
            ok = False
            print(f"[rank {rank}] {prec} MISMATCH {n}: err {e:.3e} ref {r:.3e}")
        worst = max(worst, e / (r + 1e-4 * gmax))
    if rank == 0:
        print(f"DP-{world} {prec}: grads match single-process big batch, worst rel err {worst:.2e}")
    red.remove()
# graphed data-parallel steps: two graphs with the all-reduce of the upper arena half overlapped == one graph + one
# all-reduce (same dropout masks: both runs start from the same seed / step counter)
from vit3d_b200.graphs import GraphedTrainStep
cfg = vit3d_b200.get_config(16, 512, 4, 256, 8, dropout_rate=0.1)
sd = O.init_state_dict(cfg, seed=11)
B = 8 * world
x = O.synth_volumes(B, seed=9).to(dev)[rank::world].contiguous()
y = O.synth_labels(B).to(dev)[rank::world].contiguous()
F.enable_direct_grads(True)
final = {}
for overlap in (True, False, "again"):          # "again": the single all-reduce schedule a second time = run-to-run floor
    torch.manual_seed(123)
    F._STATE["step"] = 0
    m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision="bf16").to(dev)
    m.load_state_dict(sd); m.train()
    opt = FusedSGD(m.parameters(), lr=0.05, momentum=0.9)
    step = GraphedTrainStep(m, opt, warmup=1, data_parallel=True, overlap=overlap is True)
    for _ in range(4):
        loss = step(x, y, 1.3)
    torch.cuda.synchronize()
    final[overlap] = (opt.arena.flat.clone(), float(loss), step._graphs is not None)
d = float((final[True][0] - final[False][0]).abs().max())
floor = float((final["again"][0] - final[False][0]).abs().max())     # split-K reductions are not order-deterministic
scale = float(final[False][0].abs().max())
same_everywhere = final[True][0].clone()
dist.broadcast(same_everywhere, 0)
drift = float((same_everywhere - final[True][0]).abs().max())
if not final[True][2] or final[False][2] or d > 3 * floor + 1e-6 * scale or drift != 0.0 or not (final[True][1] == final[True][1]):
    ok = False
if rank == 0:
    print(f"graphed DP-{world}: overlapped two-graph step vs single all-reduce after 4+1 steps: max weight diff {d:.2e} "
          f"(run-to-run floor of the single schedule {floor:.2e}, scale {scale:.2e}), ranks identical: {drift == 0.0}, segmented: {final[True][2]}")
# sharded ensemble
cfgs = [vit3d_b200.north_star_config(c) for c in (5, 9, 11)]
members = [VisionTransformer(c, 128, zero_head=True, num_classes=1, precision="bf16") for c in cfgs]
ens = TransformerEnsemble(*members, in_features=1)
ens.load_state_dict(O.ensemble_state_dict([O.init_state_dict(c, seed=42 + j) for j, c in enumerate(cfgs)], seed=7))
ens.to(dev).eval()
x = O.synth_volumes(13, seed=5).to(dev)
with torch.no_grad():
    full = ens(x)
    sh = ShardedEnsemble(ens, costs=[O.fwd_flops_per_volume(c) for c in cfgs])(x)
err = float((full - sh).abs().max())
if err > 1e-6:
    ok = False
if rank == 0:
    print(f"ShardedEnsemble over {world} ranks vs single GPU: max abs diff {err:.2e}")
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI-GPU CHECK", "PASSED" if float(t) == 1.0 else "FAILED")
dist.destroy_process_group()
sys.exit(0 if float(t) == 1.0 else 1)
