"""Debug helper: banded vs single-GPU on a catalog with duplicates; prints which outputs differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from vif_b200 import qxmatch, catalogs as G
from vif_b200.qxmatch import QxmatchParams
from vif_b200.distributed import qxmatch_banded, shard_bounds
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n1, n2 = 600_000, 500_000
full1 = G.fullsky(1, n1, xp="torch", device=dev)
full2 = G.fullsky(2, n2, xp="torch", device=dev)
full1 = (torch.cat((full1[0][:n1//2], full1[0][:n1//2])), torch.cat((full1[1][:n1//2], full1[1][:n1//2])))
full2[0][:1000] = full1[0][:1000]; full2[1][:1000] = full1[1][:1000]
lo, hi = shard_bounds(n1, world, rank)
ra2, dec2 = full2[0].clone(), full2[1].clone()
ident, d, rev = qxmatch_banded(full1[0][lo:hi].contiguous(), full1[1][lo:hi].contiguous(), ra2, dec2, lo, QxmatchParams(device=local))
ref = qxmatch(full1[0], full1[1], full2[0], full2[1], device=local, brute_force=True, exact_grid=True)
def rep(name, a, b):
    bad = torch.nonzero(a != b).flatten()
    print(f"rank {rank} {name}: {bad.numel()} differ", bad[:5].tolist(), a[bad[:5]].tolist(), b[bad[:5]].tolist(), flush=True)
rep("id", ident[0], ref.id[0, lo:hi]); rep("d", d[0], ref.d[0, lo:hi])
rep("rid", rev["rid"], ref.rid[rev["rows"]]); rep("rd", rev["rd"], ref.rd[rev["rows"]])
ref2 = qxmatch(full1[0], full1[1], full2[0], full2[1], device=local, brute_force=True)
rep("ref-vs-brute id", ref.id[0], ref2.id[0]); rep("ref-vs-brute rid", ref.rid, ref2.rid)
dist.destroy_process_group()
