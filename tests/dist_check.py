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
# the verification pass (attfind_verify_topk) sharded over the ranks: base logits by latent owner, the (latent, column) pair
# list split evenly, small all-gathers -- must give every rank the picks, merged list and hybrid tensor of a one-rank run
sweep = sx.attfind_sweep(G, Pool(), lat, noise, precision="bf16", max_batch=64, rank=rank, world_size=world)
multi = sx.attfind_verify_topk(G, Pool(), lat, noise, sweep, 5, 0.5, precision="fp32", max_batch=32, rank=rank, world_size=world,
                               min_candidates=8)
one = sx.attfind_verify_topk(G, Pool(), lat, noise, sweep, 5, 0.5, precision="fp32", max_batch=32, min_candidates=8)
assert multi[0] == one[0] and multi[1] == one[1], (rank, multi[:2], one[:2])
assert multi[3]["exact_evals"] == one[3]["exact_evals"] and multi[3]["verified"] == one[3]["verified"]
assert torch.equal(multi[3]["style_change"], one[3]["style_change"]) and torch.equal(multi[3]["base_prob"], one[3]["base_prob"])
full32 = sx.attfind_sweep(G, Pool(), lat, noise, precision="fp32", max_batch=64)
ref = sx.attfind_select(full32["style_change"], full32["base_prob"], 5, 0.5)
exact_ok = multi[0] == ref[0] and multi[1] == ref[1]
gathered = [None] * world
dist.all_gather_object(gathered, (multi[0], multi[1]))
assert all(g == gathered[0] for g in gathered)
dist.barrier()
print(f"rank {rank}/{world}: sharded sweep + all-gather == single-rank sweep (fp32, bf16); selection agrees; sharded verification == "
      f"one-rank verification ({multi[3]['exact_evals']} exact evals, verified={multi[3]['verified']}); picks == full fp32 sweep: {exact_ok}")
dist.destroy_process_group()
