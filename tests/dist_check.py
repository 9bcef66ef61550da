"""Multi-GPU check (run under torchrun on the GPU box, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py

Every rank sweeps its latent shard; after the ONE all-gather every rank must hold exactly the tensor a single-rank
sweep produces (same kernels, same data => bit-identical), and the device selection must agree on every rank.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import stylex_b200 as sx
from stylex_b200 import dist as sxd, synthetic

torch.set_grad_enabled(False)
rank, world, local = sxd.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
size = 64
G = sx.Generator(size, 514).to(dev)
G.load_state_dict(synthetic.make_generator_state(size, seed=42), strict=False)
lat = synthetic.make_latents(5, 42).to(dev)          # 5 latents: uneven shards at world 2 (3 + 2)
noise = synthetic.make_noise(size, 42).to(dev)
coef = torch.randn(2, 3, generator=torch.Generator().manual_seed(0)).to(dev)


class Pool:
    def classify_images(self, x):
        f = x[:, :, 5::20, 7::20].reshape(x.shape[0], 3, -1)
        return torch.stack([(f[:, :, 0] * coef[c]).sum(1) + f[:, 0, 1] for c in range(2)], 1) * 5


S = G.num_style_coords
sind = list(range(0, S, 31))
for prec in ("fp32", "bf16"):
    full = sx.attfind_sweep(G, Pool(), lat, noise, precision=prec, sindices=sind, max_batch=32, rank=rank, world_size=world)
    single = sx.attfind_sweep(G, Pool(), lat, noise, precision=prec, sindices=sind, max_batch=32)
    assert full["style_change"].shape == single["style_change"].shape == (5, 2, S, 2)
    assert torch.equal(full["style_change"], single["style_change"]), (rank, prec)
    picks = sx.attfind_select(full["style_change"], full["base_prob"], 5, 0.5)
    ref = sx.attfind_select(single["style_change"], single["base_prob"], 5, 0.5)
    assert picks[0] == ref[0] and picks[1] == ref[1], (rank, prec)
    gathered = [None] * world
    dist.all_gather_object(gathered, picks[1])
    assert all(g == gathered[0] for g in gathered)
dist.barrier()
print(f"rank {rank}/{world}: sharded sweep + all-gather == single-rank sweep (fp32, bf16); selection agrees")
dist.destroy_process_group()
