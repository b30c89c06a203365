"""2-GPU check of the data-parallel training step (run under torchrun): after one step both ranks hold IDENTICAL
parameters, and the all-reduced gradient equals the mean of the two ranks' local gradients (recomputed with world=1
trainers on each rank's shard)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from tests import cases
from tests.test_gpu_train import _inputs
from mvuld_b200 import train

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dev = "cuda"
g, img, txt, labels = _inputs(B=8, seed=100 + rank)           # each rank its own shard
model = cases.make_fusion().to(dev)
if rank != 0:
    # ranks that built / loaded different weights must end up on rank 0's (DDP's constructor broadcast)
    from mvuld_b200 import synth
    synth.randomize_for_parity(model, seed=4242 + rank)
    model.swinbn.running_mean.add_(1.0)
tr = train.FusionTrainer(model, dropout=0.0, lr=1e-3, bucket_mb=4.0)
assert tr.world == world and len(tr.buckets) > 3
p0 = [torch.empty_like(tr.flat_p) for _ in range(world)]
dist.all_gather(p0, tr.flat_p.clone())
rm = [torch.empty_like(model.swinbn.running_mean) for _ in range(world)]
dist.all_gather(rm, model.swinbn.running_mean.clone())
ref0 = cases.make_fusion().to(dev)
assert all(torch.equal(p0[0], q) for q in p0[1:]) and all(torch.equal(rm[0], q) for q in rm[1:]), "replicas not synchronised at construction"
assert torch.equal(model.fc.weight.detach(), ref0.fc.weight.detach()), "rank 0's weights did not win"
loss, _ = tr.step(g.to(dev), img.to(dev), txt.to(dev), labels.to(dev))
gsum = tr.flat_g.clone()                                      # SUM over ranks of (local grad / world) = mean gradient
# local gradient of this rank alone
m2 = cases.make_fusion().to(dev)
t2 = train.FusionTrainer(m2, dropout=0.0, lr=1e-3, world_size=1)
t2.forward_backward(g.to(dev), img.to(dev), txt.to(dev), labels.to(dev))
local = t2.flat_g.clone()
both = [torch.empty_like(local) for _ in range(world)]
dist.all_gather(both, local)
mean = sum(both) / world
err = float((gsum - mean).norm() / mean.norm())
p = tr.flat_p.clone()
ps = [torch.empty_like(p) for _ in range(world)]
dist.all_gather(ps, p)
same = all(torch.equal(ps[0], q) for q in ps[1:])
losses = [torch.empty_like(loss) for _ in range(world)]
dist.all_gather(losses, loss.clone())
if rank == 0:
    print(f"DP check: world={world} buckets={len(tr.buckets)} allreduced-vs-mean grad rel err={err:.3e} params identical across ranks={same} "
          f"losses={[round(float(l), 4) for l in losses]}")
    assert err < 1e-5 and same
dist.destroy_process_group()
