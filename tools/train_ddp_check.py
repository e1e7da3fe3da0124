"""Two (or more) ranks, one GPU each (torchrun): data-parallel head training steps with different crops per rank.
Checks that the ranks hold identical weights after every step and that the all-reduced gradient equals the mean of
the per-rank gradients; prints the step time (max over ranks, CUDA events)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cryovit_b200.train import CryoVITHeadTrainerB200  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
C, D, h, w = 1536, int(sys.argv[1]) if len(sys.argv) > 1 else 128, 32, 32
g = torch.Generator().manual_seed(100 + rank)
feats = (torch.randn(C, D, h, w, generator=g) * 0.5).half().cuda()
labels = (torch.rand(D, 16 * h, 16 * w, generator=g) < 0.1).float()
labels[::5] = -1
labels = labels.cuda()
tr = CryoVITHeadTrainerB200(C)  # same seeded init on every rank
# gradient check: all-reduced bucket == mean of the per-rank buckets
tr.forward_backward(feats, labels, 1.0 / world)
local_g = tr.flat_g.clone()
gathered = [torch.empty_like(local_g) for _ in range(world)]
dist.all_gather(gathered, local_g)
dist.all_reduce(tr.flat_g)
want = torch.stack(gathered).sum(0)
err = ((tr.flat_g - want).norm() / want.norm()).item()
for step in range(3):
    loss = tr.train_step(feats, labels)
    ref = tr.flat_p.clone()
    dist.broadcast(ref, 0)
    same = torch.equal(ref, tr.flat_p)
    if rank == 0 or not same:
        print(f"rank {rank} step {step} loss {float(loss):.5f} weights identical to rank 0: {same}", flush=True)
    assert same
ts = []
for _ in range(3):
    dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    tr.train_step(feats, labels)
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts.append(t.item())
if rank == 0:
    ms = sorted(ts)[1]
    print(f"world {world}: all-reduce vs gathered-sum rel err {err:.2e}; train step {ms:.2f} ms -> "
          f"{world * D * 512 * 512 / ms / 1e6:.3f} Gvoxel/s aggregate ({tr.flat_g.numel() * 4 / 1e6:.1f} MB gradient bucket)", flush=True)
dist.destroy_process_group()
