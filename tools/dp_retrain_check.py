"""2+ GPU check of the data-parallel retrain (N4), run under torchrun: every rank computes its local gradients, then the
same step with train_dp enabled; the averaged gradients must equal the mean of the ranks' local gradients and be
identical on every rank.  Prints one line per rank 0."""
import os
import sys
import time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc
from modelcompression_b200 import train_dp

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
model = mc.Darknet(mc.write_yolov2_voc_cfg()).to(dev)
model.set_masks(mc.weight_prune(model, 90.))
train_dp.broadcast_parameters(model)
model.train()
gen = torch.Generator(device=dev).manual_seed(10 + rank)
x = torch.rand(B, 3, 416, 416, device=dev, generator=gen)
g = torch.randn(B, 125, 13, 13, device=dev, generator=gen)
# local gradients
model.zero_grad()
(model(x) * g).sum().backward()
local_g = [p.grad.clone() for p in model.parameters()]
mean_g = []
for t in local_g:
    m = t.clone()
    dist.all_reduce(m)
    mean_g.append(m / world)
# the same step with the in-backward all-reduce
train_dp.enable(model)
model.zero_grad()
(model(x) * g).sum().backward()
worst = 0.0
for p, m in zip(model.parameters(), mean_g):
    worst = max(worst, float((p.grad - m).abs().max() / m.abs().max().clamp_min(1e-30)))
    chk = p.grad.clone()
    dist.broadcast(chk, src=0)
    assert torch.equal(chk, p.grad), "gradients differ between ranks"
# timing: DP step vs local step
opt = torch.optim.SGD(model.parameters(), lr=1e-5, momentum=0.9, weight_decay=5e-4 * B * world)
def step():
    opt.zero_grad(set_to_none=True)
    (model(x) * g).sum().backward()
    opt.step()
for _ in range(2):
    step()
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize(); dist.barrier(); t_dp = (time.perf_counter() - t0) / 5
train_dp.disable(model)
for _ in range(2):
    step()
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize(); dist.barrier(); t_local = (time.perf_counter() - t0) / 5
if rank == 0:
    print("DP retrain on %d GPUs, batch %d/GPU: averaged gradients == mean of local gradients (max rel diff %.2e), identical "
          "on all ranks; step %.2f ms with all-reduce vs %.2f ms without -> %.0f images/s"
          % (world, B, worst, t_dp * 1e3, t_local * 1e3, B * world / t_dp))
dist.barrier()
dist.destroy_process_group()
