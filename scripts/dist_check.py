"""Run under torchrun (N ranks, one GPU each): row-sharded qxmatch must equal the single-GPU result
bit for bit (ids) on the binned and the brute-force paths. Rank 0 prints OK/FAIL lines."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from vif_b200 import qxmatch, catalogs as G
from vif_b200.qxmatch import QxmatchParams
from vif_b200.distributed import qxmatch_sharded, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok_all = True
ONLY = set(sys.argv[1].split(",")) if len(sys.argv) > 1 else None      # optional subset of the cases


def _sel(cases):
    return [c for c in cases if ONLY is None or c[0] in ONLY]


for name, n1, n2, kw in _sel((("binned", 300_000, 250_000, dict()), ("binned_nth3", 100_000, 120_000, dict(nth=3)),
                         ("brute", 20_000, 15_000, dict(brute_force=True)), ("nomirror", 50_000, 50_000, dict(no_mirror=True)),
                         ("ties_across_shards", 100_000, 80_000, dict()), ("binned_nccl_merge", 300_000, 250_000, dict()),
                         ("ties_nccl_merge", 100_000, 80_000, dict()), ("sparse_unmatched", 2_000, 50_000, dict()),
                         ("fullsky", 400_000, 300_000, dict()))):
    os.environ["XM_MERGE"] = "nccl" if name.endswith("nccl_merge") else "p2p"
    gen = G.fullsky if name == "fullsky" else G.c2
    full1 = gen(1, n1, xp="torch", device=dev)
    full2 = gen(2, n2, xp="torch", device=dev)
    if name.startswith("ties"):          # the two halves of catalog 1 are the same points: exact ties between ranks
        full1 = (torch.cat((full1[0][:n1//2], full1[0][:n1//2])), torch.cat((full1[1][:n1//2], full1[1][:n1//2])))
    lo, hi = shard_bounds(n1, world, rank)
    ra2 = full2[0].clone() if rank == 0 else torch.zeros(n2, dtype=torch.float64, device=dev)
    dec2 = full2[1].clone() if rank == 0 else torch.zeros(n2, dtype=torch.float64, device=dev)
    ident, d, rid, rd = qxmatch_sharded(full1[0][lo:hi].contiguous(), full1[1][lo:hi].contiguous(), ra2, dec2, lo,
                                        QxmatchParams(device=local, **kw))
    ref = qxmatch(full1[0], full1[1], full2[0], full2[1], device=local, **kw)
    ok = torch.equal(ident, ref.id[:, lo:hi]) and torch.equal(d, ref.d[:, lo:hi])
    if not kw.get("no_mirror"):
        ok = ok and torch.equal(rid, ref.rid) and torch.equal(rd.nan_to_num(-1.0), ref.rd.nan_to_num(-1.0))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(("OK  " if flag.item() else "FAIL"), name, "world", world, flush=True)
    ok_all &= bool(flag.item())
    if name in ("binned", "ties_across_shards", "binned_nccl_merge"):
        # merged reverse match left sliced over the ranks; and the round-1 order of the collectives
        lo2, hi2 = shard_bounds(n2, world, rank)
        for label, env in (("sliced", None), ("early_bcast", "XM_DIST_EARLY_BCAST")):
            if env:
                os.environ[env] = "1"
            if rank != 0:
                ra2.zero_(); dec2.zero_()
            _, _, rs, ds = qxmatch_sharded(full1[0][lo:hi].contiguous(), full1[1][lo:hi].contiguous(), ra2, dec2, lo,
                                           QxmatchParams(device=local, **kw), replicate_reverse=(label != "sliced"))
            if env:
                del os.environ[env]
            want = (ref.rid[lo2:hi2], ref.rd[lo2:hi2]) if label == "sliced" else (ref.rid, ref.rd)
            ok = torch.equal(rs, want[0]) and torch.equal(ds.nan_to_num(-1.0), want[1].nan_to_num(-1.0))
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(("OK  " if flag.item() else "FAIL"), name + "_" + label, "world", world, flush=True)
            ok_all &= bool(flag.item())
# declination-band decomposition (qxmatch_banded): both directions sharded; must equal the single-GPU exact search
from vif_b200.distributed import qxmatch_banded
for name, n1, n2, halo in _sel((("banded", 2_000_000, 1_000_000, None), ("banded_dups", 600_000, 500_000, None),
                                 ("banded_fallback", 300_000, 200_000, 0.002))):
    full1 = G.fullsky(1, n1, xp="torch", device=dev)
    full2 = G.fullsky(2, n2, xp="torch", device=dev)
    if name == "banded_dups":            # identical points inside and across the catalogs: ties go to the smallest row
        full1 = (torch.cat((full1[0][:n1//2], full1[0][:n1//2])), torch.cat((full1[1][:n1//2], full1[1][:n1//2])))
        full2[0][:1000] = full1[0][:1000]; full2[1][:1000] = full1[1][:1000]
    lo, hi = shard_bounds(n1, world, rank)
    ra2 = full2[0].clone() if rank == 0 else torch.zeros(n2, dtype=torch.float64, device=dev)
    dec2 = full2[1].clone() if rank == 0 else torch.zeros(n2, dtype=torch.float64, device=dev)
    ident, d, rev = qxmatch_banded(full1[0][lo:hi].contiguous(), full1[1][lo:hi].contiguous(), ra2, dec2, lo,
                                   QxmatchParams(device=local), halo_deg=halo)
    _, _, full = qxmatch_banded(full1[0][lo:hi].contiguous(), full1[1][lo:hi].contiguous(), ra2, dec2, lo,
                                QxmatchParams(device=local), halo_deg=halo, gather_reverse=True, broadcast_cat2=False)
    ref = qxmatch(full1[0], full1[1], full2[0], full2[1], device=local, brute_force=True, exact_grid=True)
    fwd, _, _ = qxmatch_banded(full1[0][lo:hi].contiguous(), full1[1][lo:hi].contiguous(), ra2, dec2, lo,
                               QxmatchParams(device=local), halo_deg=halo, home=False, broadcast_cat2=False)
    k, k1 = rev["primary"], fwd["primary"]
    ok = torch.equal(ident, ref.id[:, lo:hi]) and torch.equal(d, ref.d[:, lo:hi])
    ok = ok and torch.equal(rev["rid"][k], ref.rid[rev["rows"][k]]) and torch.equal(rev["rd"][k], ref.rd[rev["rows"][k]])
    ok = ok and torch.equal(full["rid"], ref.rid) and torch.equal(full["rd"], ref.rd)
    ok = ok and torch.equal(fwd["id"][k1], ref.id[0][fwd["rows"][k1]]) and torch.equal(fwd["d"][k1], ref.d[0][fwd["rows"][k1]])
    cnt = torch.tensor([int(k.sum())], device=dev)
    dist.all_reduce(cnt)
    ok = ok and int(cnt.item()) == n2
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(("OK  " if flag.item() else "FAIL"), name, "world", world, flush=True)
    ok_all &= bool(flag.item())
# partitioned run (replicated catalogs, row blocks of both): must equal the single-GPU result
from vif_b200.distributed import qxmatch_partitioned
for name, n1, n2 in _sel((("partitioned", 300_001, 250_003), ("partitioned_even", 200_000, 100_000))):
    full1 = G.c2(1, n1, xp="torch", device=dev)
    full2 = G.c2(2, n2, xp="torch", device=dev)
    lo, hi = shard_bounds(n1, world, rank)
    ra2 = full2[0].clone() if rank == 0 else torch.zeros(n2, dtype=torch.float64, device=dev)
    dec2 = full2[1].clone() if rank == 0 else torch.zeros(n2, dtype=torch.float64, device=dev)
    ident, d, rid, rd, (lo2, hi2) = qxmatch_partitioned(full1[0][lo:hi].contiguous(), full1[1][lo:hi].contiguous(), ra2,
                                                       dec2, lo, n1, QxmatchParams(device=local))
    ref = qxmatch(full1[0], full1[1], full2[0], full2[1], device=local)
    ok = torch.equal(ident, ref.id[:, lo:hi]) and torch.equal(d, ref.d[:, lo:hi]) and \
        torch.equal(rid, ref.rid[lo2:hi2]) and torch.equal(rd, ref.rd[lo2:hi2])
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(("OK  " if flag.item() else "FAIL"), name, "world", world, flush=True)
    ok_all &= bool(flag.item())
# partitioned self match, nth=5, clustered (C3-shaped)
ra, dec = G.c3_clustered(3, 5, xp="torch", device=dev)
n = ra.numel()
lo, hi = shard_bounds(n, world, rank)
ident, d, _, _, _ = qxmatch_partitioned(ra[lo:hi].contiguous(), dec[lo:hi].contiguous(), None, None, lo, n,
                                        QxmatchParams(device=local, nth=5), self_match=True)
ref = qxmatch(ra, dec, nth=5, device=local)
ok = torch.equal(ident, ref.id[:, lo:hi]) and torch.equal(d.nan_to_num(-1.0), ref.d[:, lo:hi].nan_to_num(-1.0))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(("OK  " if flag.item() else "FAIL"), "partitioned_self_nth5", "world", world, flush=True)
ok_all &= bool(flag.item())
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
