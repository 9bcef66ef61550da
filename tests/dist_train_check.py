"""Multi-GPU check of the training slice (run under torchrun on the GPU box, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/dist_train_check.py

``DistributedDataParallel`` around the drop-in ``Generator``: every rank runs the native forward + backward on its shard
of the batch; DDP's NCCL all-reduce averages the parameter gradients (the reference wraps S, G, D_aug the same way,
ST:1186-1193).  The averaged gradients times the world size must equal the gradients of a single process over the whole
batch (same kernels; summation order differs only across the rank boundary -> tolerance 1e-5 relative).
"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

import stylex_b200 as sx
from stylex_b200 import dist as sxd, synthetic

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
rank, world, local = sxd.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
size, cap, per_rank = 32, 8, 2
G = sx.Generator(size, 514, network_capacity=cap).to(dev)
G.load_state_dict(synthetic.make_generator_state(size, seed=7, network_capacity=cap), strict=False)
G.train()
single = copy.deepcopy(G)
n = per_rank * world
styles = sx.styles_def_to_tensor([(synthetic.make_latents(n, 3).to(dev), G.num_layers)])
noise = synthetic.make_noise(size, 7).to(dev)
go = torch.randn(n, 3, size, size, generator=torch.Generator().manual_seed(11)).to(dev)

ddp = DDP(G, device_ids=[local])
lo, hi = rank * per_rank, (rank + 1) * per_rank
rgb = ddp(styles[lo:hi], noise)
(rgb * go[lo:hi]).sum().backward()

ref = single(styles, noise)
(ref * go).sum().backward()
assert torch.equal(ref[lo:hi].detach(), rgb.detach()), "per-sample forward must not depend on the batch it is in"
worst = 0.0
for (k, p), (_, q) in zip(G.named_parameters(), single.named_parameters()):
    assert p.grad is not None and q.grad is not None, k
    err = float((p.grad * world - q.grad).abs().max()) / max(1.0, float(q.grad.abs().max()))
    worst = max(worst, err)
    assert err <= 1e-5, (k, err)
dist.barrier()
print(f"rank {rank}/{world}: DDP(Generator) native forward+backward, NCCL-averaged gradients == single-process gradients "
      f"(max rel err {worst:.2e})")
dist.destroy_process_group()
