"""Row-sharded multi-GPU qxmatch: one process per GPU, `torch.distributed` (NCCL over NVLink) for
the plumbing. Mirrors how the reference splits work between its std::threads -- contiguous chunks
of catalog-1 rows per worker (qxmatch.hpp:422-459) -- with one exchange step for the reverse match.

  1. catalog 2 is replicated: `dist.broadcast` from `src` (NCCL broadcast over NVSwitch);
  2. the bounding box of catalog 1 is made global with two 2-element all-reduces (min / max), so
     every rank bins with the grid of qxmatch.hpp:232-267 computed from the WHOLE catalogs;
  3. every rank runs the single-GPU pipeline on its row shard against the full catalog 2
     (forward results never leave the rank: rows are independent);
  4. reverse match: nearest catalog-1 row over all shards, the smallest global row index on exact
     ties (the reference's thread=1 rule, qxmatch.hpp:510-513). On GPUs this is ONE kernel over
     NVLink peer memory (xm_rev_merge_p2p): every rank's search writes (rd, rid) straight into a
     peer-mapped exchange buffer, each rank merges its 1/N slice of the rows out of its peers'
     buffers and stores the merged rows into every peer; torch symmetric memory supplies the
     mappings and the two device-side barriers. Fallback (no peer mappings, XM_MERGE=nccl, CPU
     tensors): all-reduce(min) on `rd`, then all-reduce(min) on `rid` masked to the ranks whose
     local distance equals the global minimum.

No collective touches the forward path. `compute` and `bbox` are injectable so that the sharding
and merge logic is testable on CPU with the gloo backend (tests/test_distributed_gloo.py); the
defaults are the CUDA library and there is no CPU implementation in this module.
"""
import os
import time


def shard_bounds(n, world_size, rank):
    """Contiguous row block [begin, end) of `rank` (floor split, remainder to the last ranks)."""
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


def _cuda_bbox(ra, dec):
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    out = (C.c_double * 4)()
    bad = C.c_uint64()
    err = C.create_string_buffer(512)
    import torch
    if ra.numel() == 0:
        # an empty row shard (fewer rows than ranks): neutral element of the min/max all-reduce, so that the
        # rank still enters every collective of the call
        return [float("inf"), float("-inf"), float("inf"), float("-inf")]
    torch.cuda.current_stream(ra.device).synchronize()
    rc = lib.xm_bbox_dev(ra.data_ptr(), dec.data_ptr(), ra.numel(), ra.device.index or 0, None, out,
                         C.byref(bad), err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return [out[0], out[1], out[2], out[3]]


def _cuda_bbox_neg(ra, dec, out):
    """xm_bbox_neg_dev: {-ra_min, ra_max, -dec_min, dec_max, non-finite rows} into the device tensor `out`
    (5 doubles), on torch's current stream, without a host round trip."""
    import ctypes as C
    import torch
    from . import _lib
    err = C.create_string_buffer(512)
    stream = torch.cuda.current_stream(out.device)
    n = ra.numel()
    rc = _lib.load().xm_bbox_neg_dev(ra.data_ptr() if n else None, dec.data_ptr() if n else None, n,
                                     out.device.index or 0, stream.cuda_stream, out.data_ptr(), err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())


def _cuda_compute(ra1, dec1, ra2, dec2, params, out=None):
    from .qxmatch import qxmatch
    return qxmatch(ra1, dec1, ra2, dec2, params=params, out=out)


_COMM = {}          # device index -> side stream that carries the broadcast of catalog 2


def qxmatch_sharded(ra1_local, dec1_local, ra2, dec2, row_offset, params, group=None, src=0,
                    broadcast_cat2=True, compute=None, bbox=None, replicate_reverse=True):
    """Runs one rank's part of a row-sharded qxmatch.

    ra1_local/dec1_local: this rank's rows [row_offset, row_offset + len) of catalog 1.
    ra2/dec2: full-size catalog-2 tensors (contents only need to be valid on `src` when
    broadcast_cat2). Returns (id_local, d_local, rid, rd): forward results for the local rows,
    reverse results replicated on every rank with GLOBAL catalog-1 row ids. Ids are int64 bit
    patterns (NPOS == -1) as in the single-GPU device path.
    replicate_reverse=False leaves the merged reverse match sliced over the ranks, as the forward
    match is: rid / rd then cover catalog-2 rows [shard_bounds(n2, world, rank)) only (half the
    NVLink traffic of the merge).

    On GPUs the call is laid out so that nothing waits for catalog 2 longer than it must:
    the bounding boxes of BOTH catalogs come out of one 8-element all-reduce (rank `src` knows catalog 2),
    the broadcast of catalog 2 is issued behind it on a side stream, and the library is told the boxes
    (xm_params.bbox1 / bbox2) and the event that marks the end of the broadcast
    (xm_params.cat2_ready_event): it indexes the local rows of catalog 1 while catalog 2 is in flight.
    """
    import torch
    import torch.distributed as dist
    from .qxmatch import QxmatchParams

    on_gpu = ra1_local.is_cuda and compute is None
    compute = compute or _cuda_compute
    bbox = bbox or _cuda_bbox
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    timing = os.environ.get("XM_DIST_TIMING") and ra1_local.is_cuda
    marks = []

    def mark(name):
        if timing:
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    mark("start")
    if params.self:
        raise ValueError("self-match shards catalog 1 against itself: pass the full catalog as "
                         "catalog 2 and self=False with an explicit id offset instead")
    late_cat2 = bool(broadcast_cat2 and world > 1 and on_gpu and not params.brute_force
                     and not os.environ.get("XM_DIST_EARLY_BCAST"))
    pending = []
    if broadcast_cat2 and world > 1 and not late_cat2:     # in flight while the bounding box is exchanged
        pending = [dist.broadcast(ra2, src=src, group=group, async_op=True),
                   dist.broadcast(dec2, src=src, group=group, async_op=True)]

    # global bounding box of catalog 1 (qxmatch.hpp:232-238): one all-reduce(max) on
    # (-ra_min, ra_max, -dec_min, dec_max); with late_cat2 catalog 2's box rides along (known to `src` only)
    ready = None
    if late_cat2 and bbox is _cuda_bbox:
        # Everything stays on the device until the one read after the all-reduce, and everything runs on ONE side
        # stream (never torch's legacy default stream: handed that handle, the library would launch on a stream of
        # its own, unordered with the tensors made here): two bounding-box passes (the second on `src` only) into
        # {-min, max, -min, max, bad rows} x 2, the 10-double all-reduce(max), then the broadcast of catalog 2 --
        # behind the all-reduce, because collectives of one communicator run in the order they were issued.
        dev = ra1_local.device
        comm = _COMM.get(dev.index or 0)
        if comm is None:
            comm = _COMM[dev.index or 0] = torch.cuda.Stream(device=dev)
        comm.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(comm):
            ext = torch.full((10,), float("-inf"), dtype=torch.float64, device=dev)
            _cuda_bbox_neg(ra1_local, dec1_local, ext[0:5])
            if rank == src:
                _cuda_bbox_neg(ra2, dec2, ext[5:10])
            dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
            works = [dist.broadcast(ra2, src=src, group=group, async_op=True),
                     dist.broadcast(dec2, src=src, group=group, async_op=True)]
            e = ext.tolist()                   # the one host round trip: waits for the all-reduce, not the broadcast
            for w in works:
                w.wait()                       # orders `comm` behind NCCL's stream; does not block the host
            ready = torch.cuda.Event()
            ready.record(comm)
        if e[4] > 0 or e[9] > 0:
            raise RuntimeError("%s RA and Dec coordinates contain invalid values (infinite or NaN)"
                               % ("first" if e[4] > 0 else "second"))
        ext = e[0:4] + e[5:9]
        extra = {"bbox1": (-ext[0], ext[1], -ext[2], ext[3]), "bbox2": (-ext[4], ext[5], -ext[6], ext[7]),
                 "cat2_ready_event": int(ready.cuda_event)}
    else:
        late_cat2 = False
        bb = bbox(ra1_local, dec1_local)
        ext = torch.tensor([-bb[0], bb[1], -bb[2], bb[3]], dtype=torch.float64, device=ra1_local.device)
        if world > 1:
            if broadcast_cat2 and not pending:
                pending = [dist.broadcast(ra2, src=src, group=group, async_op=True),
                           dist.broadcast(dec2, src=src, group=group, async_op=True)]
            dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
        ext = ext.tolist()
        extra = {"bbox1": (-ext[0], ext[1], -ext[2], ext[3])}
    for w in pending:
        w.wait()
    mark("bbox" if late_cat2 else "broadcast+bbox")
    p = QxmatchParams(**{**params.__dict__, **extra})

    # the reverse results go straight into the peer-mapped exchange buffer of the merge when there is one
    ent = None
    if on_gpu and world > 1 and not p.no_mirror:
        ent = _p2p_buffers(ra2.numel(), ra2.device, world, group)
    if ent is not None:
        n2 = ra2.numel()
        res = compute(ra1_local, dec1_local, ra2, dec2, p,
                      out={"rd": ent[0][:n2], "rid": ent[0][n2:2*n2].view(torch.int64)})
    else:
        res = compute(ra1_local, dec1_local, ra2, dec2, p)
    if ready is not None:
        torch.cuda.current_stream(ra1_local.device).wait_event(ready)      # later users of ra2 / dec2 on this stream
    mark("compute")
    ident, d, rid, rd = res.id, res.d, res.rid, res.rd
    if p.no_mirror:
        return ident, d, rid, rd

    rid = torch.as_tensor(rid).view(torch.int64) if not isinstance(rid, torch.Tensor) else rid
    rd = torch.as_tensor(rd) if not isinstance(rd, torch.Tensor) else rd
    lo2, hi2 = shard_bounds(rid.numel(), world, rank)
    if rid.is_cuda:
        merged = _merge_reverse_p2p(rid, rd, int(row_offset), bool(p.linear), world, group,
                                    replicate=replicate_reverse) if world > 1 else False
        if merged:
            rid, rd = merged
        else:
            _merge_reverse_cuda(rid, rd, int(row_offset), bool(p.linear), world, group)
            if not replicate_reverse:
                rid, rd = rid[lo2:hi2], rd[lo2:hi2]
    else:
        rid, rd = _merge_reverse_torch(rid, rd, int(row_offset), world, group)
        if not replicate_reverse:
            rid, rd = rid[lo2:hi2], rd[lo2:hi2]
    mark("merge")
    if timing and dist.get_rank(group) == 0:
        print("[qxmatch_sharded] " + " ".join("%s=%.3fms" % (marks[k][0], 1e3*(marks[k][1] - marks[k-1][1]))
                                              for k in range(1, len(marks))), flush=True)
    return ident, d, rid, rd


def qxmatch_partitioned(ra1_local, dec1_local, ra2, dec2, row_offset, n1_total, params, group=None, src=0,
                        broadcast_cat2=True, self_match=False):
    """Partitioned run for catalogs small enough to replicate (the benchmark's 1e7 x 1e7): the
    reference's threaded driver gives every thread a chunk of catalog-1 rows for the forward match
    AND a chunk of catalog-2 rows for the reverse match over shared buckets (qxmatch.hpp:422-454);
    here every rank gets the same two chunks over replicated catalogs.

      1. catalog 2: NCCL broadcast from `src`; catalog 1: NCCL all-gather of the row shards;
      2. every rank indexes both whole catalogs (they are each other's candidates) and its two row
         blocks, then matches block against whole catalog (xm_params.has_ranges);
      3. no collective afterwards: id/d cover this rank's catalog-1 rows, rid/rd this rank's
         catalog-2 rows [shard_bounds(n2, world, rank)), with GLOBAL row ids.

    Compared with qxmatch_sharded this trades the reverse search over every catalog-2 row plus two
    all-reduces for one all-gather of catalog 1 -- the better deal while catalog 1 is small; for
    1e9-row catalogs qxmatch_sharded (no replication of catalog 1) is the one to use.
    self_match=True: one catalog matched against itself (ra2/dec2 ignored): the row shards are
    all-gathered, every rank matches its block against the whole catalog (C3: near-linear scaling,
    the search dominates the replicated index build). The binned self match has no reverse pass.
    Returns (id, d, rid, rd, (rev_begin, rev_end)).
    """
    import torch
    import torch.distributed as dist
    from .qxmatch import QxmatchParams, qxmatch

    live = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if live else 1
    rank = dist.get_rank(group) if live else 0
    if broadcast_cat2 and world > 1 and not self_match:
        dist.broadcast(ra2, src=src, group=group)
        dist.broadcast(dec2, src=src, group=group)
    n_loc = ra1_local.numel()
    if world > 1:
        bounds = [shard_bounds(n1_total, world, r) for r in range(world)]
        if bounds[rank] != (int(row_offset), int(row_offset) + n_loc):
            raise ValueError("qxmatch_partitioned expects the row blocks of shard_bounds()")
        mx = max(hi - lo for lo, hi in bounds)
        even = all(hi - lo == mx for lo, hi in bounds)
        send = torch.stack((ra1_local, dec1_local)) if n_loc == mx else \
            torch.nn.functional.pad(torch.stack((ra1_local, dec1_local)), (0, mx - n_loc))
        buf = torch.empty((world, 2, mx), dtype=torch.float64, device=ra1_local.device)
        dist.all_gather_into_tensor(buf, send.contiguous(), group=group)
        if even:
            ra1 = buf[:, 0, :].reshape(-1)
            dec1 = buf[:, 1, :].reshape(-1)
        else:
            ra1 = torch.cat([buf[r, 0, :hi - lo] for r, (lo, hi) in enumerate(bounds)])
            dec1 = torch.cat([buf[r, 1, :hi - lo] for r, (lo, hi) in enumerate(bounds)])
        ra1, dec1 = ra1.contiguous(), dec1.contiguous()
    else:
        ra1, dec1 = ra1_local, dec1_local
    if self_match:
        p = QxmatchParams(**{**params.__dict__, "self": True, "no_mirror": True,
                             "fwd_range": (int(row_offset), n_loc), "rev_range": (0, 0)})
        res = qxmatch(ra1, dec1, ra1, dec1, params=p)
        return res.id, res.d, res.rid, res.rd, (0, 0)
    lo2, hi2 = shard_bounds(ra2.numel(), world, rank)
    p = QxmatchParams(**{**params.__dict__, "fwd_range": (int(row_offset), n_loc), "rev_range": (lo2, hi2 - lo2)})
    res = qxmatch(ra1, dec1, ra2, dec2, params=p)
    return res.id, res.d, res.rid, res.rd, (lo2, hi2)


def band_edges(world, polar_weight=None):
    """Declination bands [deg], equal in area except the two that hold a pole: the spherical search
    hands the rows next to a pole to its exact kernel (xm_sphere.cu), about 9 ms of extra work per cap on
    1e9 x 1e8, so those two bands get `polar_weight` (default 0.65 for 4 or more bands, else 1) of the
    others' area. The outer edges are pushed past the poles so that every finite declination falls inside."""
    import numpy as np
    w = np.ones(world)
    if world >= 4:
        w[0] = w[-1] = 0.65 if polar_weight is None else float(polar_weight)
    z = np.concatenate(([0.0], np.cumsum(w))) / w.sum()
    e = np.degrees(np.arcsin(np.clip(-1.0 + 2.0 * z, -1.0, 1.0)))
    e[0], e[-1] = -90.000001, 90.000001
    return e


def default_halo_deg(n1, n2):
    """Halo width: six mean neighbour spacings of the sparser catalog (the chance that a nearest
    neighbour lies further is exp(-36 pi)); rows that still fail the completeness test make the call
    fall back to the row-sharded algorithm."""
    import math
    rho = max(min(n1, n2), 1) / (4.0 * math.pi)
    return max(math.degrees(6.0 / math.sqrt(rho)), 1e-3)


def _cuda_partition(ra, dec, row0, edges, halo, only):
    """xm_band_partition_dev: (triples (m, 3) float64, counts (nb,) int64) -- see the ABI header."""
    import ctypes as C
    import torch
    from . import _lib
    lib = _lib.load()
    n, nb = ra.numel(), len(edges) - 1
    # triples the output must hold: every row once, plus the rows within `halo` of an edge (estimated from
    # the thinnest band with 50 % head room; a failed call is repeated with the safe bound of 3 per row)
    thin = min(float(edges[k + 1] - edges[k]) for k in range(nb))
    dup = min(2.0, 3.0 * float(halo) / max(thin, 1e-9)) if nb > 1 else 0.0
    share = 1.0 if only < 0 else min(1.0, 3.0 / nb)
    total = C.c_uint64()
    err = C.create_string_buffer(512)
    e = (C.c_double * (nb + 1))(*[float(v) for v in edges])
    torch.cuda.current_stream(ra.device).synchronize()
    for cap in (int(n * share * (1.0 + dup)) + 4096, 3 * n + 16):
        cap = min(cap, 3 * n + 16)
        out = torch.empty((cap, 3), dtype=torch.float64, device=ra.device)
        counts = torch.zeros(nb, dtype=torch.int64, device=ra.device)
        rc = lib.xm_band_partition_dev(ra.data_ptr(), dec.data_ptr(), n, int(row0), e, nb, float(halo), int(only),
                                       out.data_ptr(), cap, counts.data_ptr(), C.byref(total), ra.device.index or 0, None,
                                       err, 512)
        if rc == 0:
            break
        if b"do not fit" not in err.value:
            raise RuntimeError(err.value.decode())
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return out[:total.value], counts


_ROW_MASK = 0x7FFFFFFFFFFFFFFF


def _cuda_finish(idl, d, bits_src, bits_cand, hlim, bad):
    """xm_band_finish_dev: candidate index -> global row in place, completeness count into `bad`."""
    import ctypes as C
    import torch
    from . import _lib
    err = C.create_string_buffer(512)
    torch.cuda.current_stream(idl.device).synchronize()
    rc = _lib.load().xm_band_finish_dev(idl.data_ptr(), d.data_ptr(), bits_src.data_ptr(), bits_cand.data_ptr(), idl.numel(),
                                        float(hlim), bad.data_ptr(), idl.device.index or 0, None, err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())


def _torch_finish(idl, d, bits_src, bits_cand, hlim, bad):
    """Same contract with torch ops (CPU tensors: the gloo tests)."""
    import torch
    g = bits_cand & _ROW_MASK
    idl.copy_(torch.where(idl >= 0, g[idl.clamp_min(0)], idl) if g.numel() else idl)
    bad += ((bits_src < 0) & ~(d < hlim)).sum()


def _cuda_home(got, sent, row_offset, id_out, d_out):
    import ctypes as C
    import torch
    from . import _lib
    err = C.create_string_buffer(512)
    torch.cuda.current_stream(got.device).synchronize()
    rc = _lib.load().xm_band_home_dev(got.data_ptr(), sent.data_ptr(), sent.shape[0], int(row_offset), id_out.data_ptr(),
                                      d_out.data_ptr(), got.device.index or 0, None, err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())


def _torch_home(got, sent, row_offset, id_out, d_out):
    import torch
    sbits = sent[:, 2].contiguous().view(torch.int64)
    own = torch.nonzero(sbits < 0).flatten()
    rows = (sbits[own] & _ROW_MASK) - int(row_offset)
    id_out[0, rows] = got[own, 0].contiguous().view(torch.int64)
    d_out[0, rows] = got[own, 1]


def qxmatch_banded(ra1_local, dec1_local, ra2, dec2, row_offset, params, group=None, src=0, broadcast_cat2=True,
                   halo_deg=None, home=True, gather_reverse=False, compute=None, partition=None, bbox=None):
    """Spatially decomposed multi-GPU qxmatch for catalogs too large to replicate (BASELINE config 5:
    1e9 x 1e8 full sky): BOTH directions shard. The reference's threaded driver gives every worker a
    chunk of catalog-1 rows and a chunk of catalog-2 rows over shared buckets (qxmatch.hpp:422-454);
    across GPUs the buckets cannot be shared, so the rows are regrouped by declination band:

      1. catalog 2 is replicated (NCCL broadcast from `src`, as the north star prescribes) while every
         rank partitions its catalog-1 rows by band (xm_band_partition_dev: order preserving, rows
         within `halo_deg` of an edge go to both sides);
      2. the bands are exchanged with ONE NCCL all-to-all over NVLink (24 bytes per row: ra, dec, global
         row) while rank b selects band b + halo out of catalog 2 (same kernel);
      3. rank b runs the single-GPU search on band x band: the forward match of the band's catalog-1
         rows and the reverse match of the band's catalog-2 rows each see every candidate closer than
         the halo. The search has brute-force semantics (exact neighbours, equal proxies to the smallest
         row; every band holds its rows in ascending global row order, so ties resolve as on one GPU);
      4. completeness (xm_band_finish_dev): a result farther than the halo (or unmatched) could have a
         nearer candidate outside the band. One all-reduce counts such rows; if there are any, the call
         is redone by qxmatch_sharded (never observed on uniform catalogs: exp(-36 pi));
      5. home=True: forward results travel back to the rows' owner ranks (second all-to-all, 16 bytes
         per row) -> (id_local, d_local) as qxmatch_sharded returns them. home=False leaves them with
         the band: fwd = dict(rows, primary, id, d) over the band's catalog-1 triples (rows = global row,
         entries with primary == False are halo copies and carry no result of their own);
         reverse results always stay with the band: rev = dict(rows, primary, rid, rd) (gather_reverse=True
         assembles full-size replicated rid / rd instead).

    Returns (id_local, d_local, rev) or, with home=False, (fwd, None, rev). nth = 1 only."""
    import torch
    import torch.distributed as dist
    from .qxmatch import QxmatchParams

    fallback_kw = {"compute": compute, "bbox": bbox}
    on_gpu = ra1_local.is_cuda
    compute = compute or _cuda_compute
    partition = partition or _cuda_partition
    finish = _cuda_finish if on_gpu else _torch_finish
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if params.nth > 1 or params.self or params.linear or params.no_mirror:
        raise ValueError("qxmatch_banded serves nth = 1 spherical matches with the reverse match on")
    timing = os.environ.get("XM_DIST_TIMING") and on_gpu
    marks = []

    def mark(name):
        if timing:
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    mark("start")
    dev = ra1_local.device
    n_loc, n2 = ra1_local.numel(), ra2.numel()
    pending = []
    if broadcast_cat2 and world > 1:          # in flight while catalog 1 is partitioned
        pending = [dist.broadcast(ra2, src=src, group=group, async_op=True),
                   dist.broadcast(dec2, src=src, group=group, async_op=True)]
    edges = band_edges(world)
    if halo_deg is not None:
        halo = float(halo_deg)
    else:
        # every rank must use the SAME halo (senders duplicate rows by it, receivers test completeness
        # against it): derive it from the largest shard, one 8-byte all-reduce
        nmax = torch.tensor([n_loc], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=group)
        halo = default_halo_deg(int(nmax.item()) * world, n2)

    # catalog 1: partition the local rows, exchange the bands
    send1, cnt_send = partition(ra1_local, dec1_local, int(row_offset), edges, halo, -1)
    cnt_recv = torch.empty_like(cnt_send)
    if world > 1:
        dist.all_to_all_single(cnt_recv, cnt_send, group=group)
    else:
        cnt_recv.copy_(cnt_send)
    s_split, r_split = [int(v) for v in cnt_send.tolist()], [int(v) for v in cnt_recv.tolist()]
    recv1 = torch.empty((sum(r_split), 3), dtype=torch.float64, device=dev)
    for w in pending:
        w.wait()
    mark("partition+broadcast")
    a2a = None
    if world > 1:
        a2a = dist.all_to_all_single(recv1, send1, r_split, s_split, group=group, async_op=True)
    else:
        recv1.copy_(send1)
    # catalog 2: this rank's band + halo out of the replicated catalog (while the all-to-all is in flight)
    sel2, _ = partition(ra2, dec2, 0, edges, halo, rank)
    ra2b, dec2b = sel2[:, 0].contiguous(), sel2[:, 1].contiguous()
    bits2 = sel2[:, 2].contiguous().view(torch.int64)
    if a2a is not None:
        a2a.wait()
    ra1b, dec1b = recv1[:, 0].contiguous(), recv1[:, 1].contiguous()
    bits1 = recv1[:, 2].contiguous().view(torch.int64)
    mark("exchange+select")

    # band x band with brute-force semantics through the exact spherical grid search: the exact nearest
    # neighbours, i.e. what the reference's binned search returns wherever that search is exact
    p = QxmatchParams(**{**params.__dict__, "bbox1": None, "brute_force": True, "exact_grid": True})
    if ra1b.numel() and ra2b.numel():
        res = compute(ra1b, dec1b, ra2b, dec2b, p)
        idl, d = torch.as_tensor(res.id).view(torch.int64)[0], torch.as_tensor(res.d)[0]
        ridl, rd = torch.as_tensor(res.rid).view(torch.int64), torch.as_tensor(res.rd)
    else:                       # an empty band: nothing matched here, the completeness test decides
        idl = torch.full((ra1b.numel(),), -1, dtype=torch.int64, device=dev)
        d = torch.full((ra1b.numel(),), float("nan"), dtype=torch.float64, device=dev)
        ridl = torch.full((ra2b.numel(),), -1, dtype=torch.int64, device=dev)
        rd = torch.full((ra2b.numel(),), float("nan"), dtype=torch.float64, device=dev)
    mark("compute")
    # local candidate index -> global row; completeness
    hlim = halo * 3600.0 * (1.0 - 1e-9)
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    finish(idl, d, bits1, bits2, hlim, bad)
    finish(ridl, rd, bits2, bits1, hlim, bad)
    if world > 1:
        dist.all_reduce(bad, group=group)
    if int(bad.item()) != 0:
        if rank == 0 and os.environ.get("XM_DIST_TIMING"):
            print("[qxmatch_banded] %d rows beyond the halo (%.4g deg): row-sharded fallback" % (int(bad.item()), halo), flush=True)
        ident, dd, rid, rdd = qxmatch_sharded(ra1_local, dec1_local, ra2, dec2, row_offset, p, group=group, src=src,
                                              broadcast_cat2=False, **fallback_kw)
        if gather_reverse:
            rev = {"rid": rid, "rd": rdd}
        else:
            rows = torch.nonzero((dec2 >= edges[rank]) & (dec2 < edges[rank + 1])).flatten()
            rev = {"rows": rows, "primary": torch.ones_like(rows, dtype=torch.bool), "rid": rid[rows], "rd": rdd[rows]}
        if home:
            return ident, dd, rev
        lo = int(row_offset)
        rows1 = torch.arange(lo, lo + n_loc, device=dev)
        return {"rows": rows1, "primary": torch.ones_like(rows1, dtype=torch.bool), "id": ident[0], "d": dd[0]}, None, rev
    mark("complete")

    rev = {"rows": bits2 & _ROW_MASK, "primary": bits2 < 0, "rid": ridl, "rd": rd}
    if home:
        # forward results back to the owners of the rows (same splits, other way round)
        back = torch.stack((idl.view(torch.float64), d), dim=1)
        got = torch.empty((send1.shape[0], 2), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_to_all_single(got, back, s_split, r_split, group=group)
        else:
            got.copy_(back)
        id_local = torch.full((1, n_loc), -1, dtype=torch.int64, device=dev)
        d_local = torch.full((1, n_loc), float("nan"), dtype=torch.float64, device=dev)
        (_cuda_home if on_gpu else _torch_home)(got, send1, row_offset, id_local, d_local)
        out = (id_local, d_local)
        mark("home")
    else:
        out = ({"rows": bits1 & _ROW_MASK, "primary": bits1 < 0, "id": idl, "d": d}, None)

    if gather_reverse:
        keep2 = torch.nonzero(rev["primary"]).flatten()
        rows2, rid2, rd2 = rev["rows"][keep2], rev["rid"][keep2], rev["rd"][keep2]
        rid_full = torch.full((n2,), -1, dtype=torch.int64, device=dev)
        rd_full = torch.full((n2,), float("nan"), dtype=torch.float64, device=dev)
        if world > 1:
            sizes = torch.tensor([keep2.numel()], dtype=torch.int64, device=dev)
            allsz = [torch.empty_like(sizes) for _ in range(world)]
            dist.all_gather(allsz, sizes, group=group)
            mx = max(int(v.item()) for v in allsz)
            pack = torch.zeros((mx, 3), dtype=torch.float64, device=dev)
            pack[:keep2.numel(), 0] = rows2.view(torch.float64)
            pack[:keep2.numel(), 1] = rid2.view(torch.float64)
            pack[:keep2.numel(), 2] = rd2
            allp = torch.empty((world * mx, 3), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allp, pack, group=group)
            allp = allp.view(world, mx, 3)
            for r in range(world):
                m = int(allsz[r].item())
                rr = allp[r, :m, 0].contiguous().view(torch.int64)
                rid_full[rr] = allp[r, :m, 1].contiguous().view(torch.int64)
                rd_full[rr] = allp[r, :m, 2]
        else:
            rid_full[rows2] = rid2
            rd_full[rows2] = rd2
        rev = {"rid": rid_full, "rd": rd_full}
        mark("gather")
    if timing and rank == 0:
        print("[qxmatch_banded] halo=%.4g deg " % halo + " ".join("%s=%.3fms" % (marks[k][0], 1e3*(marks[k][1] - marks[k-1][1]))
                                                                  for k in range(1, len(marks))), flush=True)
    return out[0], out[1], rev


_P2P = {}           # (device, n, world, group) -> symmetric buffers of the peer-memory merge
_P2P_MAX = 4        # distinct catalog-2 sizes that get peer-mapped buffers (32 n + 8 bytes each per rank)
_P2P_OFF = set()    # groups on which symmetric memory could not be set up: NCCL merge from then on


def _merge_reverse_p2p(rid, rd, row_offset, linear, world, group, replicate=True):
    """Reverse merge as ONE kernel over NVLink peer memory (xm_rev_merge_p2p): every rank publishes
    its (rd, rid, first row) in a symmetric buffer, merges its slice of the rows straight out of its
    peers' memory and stores the result into every peer. torch's symmetric memory provides the peer
    mappings and the two device-side barriers. Returns (rid, rd), or False when symmetric memory is
    unavailable (the caller falls back to the NCCL all-reduce merge); XM_MERGE=nccl forces that.
    replicate=False: every rank keeps the rows it merged, [shard_bounds(n, world, rank)), and nothing is
    stored into the peers."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _lib
    n = rid.numel()
    dev = rid.device
    ent = _p2p_buffers(n, dev, world, group)
    if ent is None:
        return False
    exch, out, hx, ho, px, po, side = ent
    lib = _lib.load()
    err = C.create_string_buffer(512)
    cur = torch.cuda.current_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):        # a non-default stream: the library launches on the handle it is given
        if rd.data_ptr() != exch.data_ptr():
            exch[:n].copy_(rd)
        if rid.data_ptr() != exch[n:].data_ptr():
            exch[n:2*n].view(torch.int64).copy_(rid)
        exch[2*n:].view(torch.int64).fill_(row_offset)
        hx.barrier(channel=0)
        me = dist.get_rank(group)
        if replicate:
            pout = po
        else:
            pout = (C.c_void_p*world)(*[(po[q] if q == me else None) for q in range(world)])
        rc = lib.xm_rev_merge_p2p(px, pout, n, me, world, float("inf") if linear else float("nan"),
                                  dev.index or 0, side.cuda_stream, err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        ho.barrier(channel=0)      # every peer has read this rank's exchange buffer (and stored its rows here)
        lo, hi = (0, n) if replicate else shard_bounds(n, world, me)
        res_rd = out[lo:hi].clone()
        res_rid = out[n + lo:n + hi].view(torch.int64).clone()
    cur.wait_stream(side)
    return res_rid, res_rd


def _p2p_buffers(n, dev, world, group):
    """Symmetric (peer-mapped) buffers of the merge for n catalog-2 rows, created on first use --
    a collective call. None when symmetric memory is unavailable or XM_MERGE=nccl."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    gkey = id(group) if group is not None else 0
    if os.environ.get("XM_MERGE", "p2p") == "nccl" or gkey in _P2P_OFF or world > 16:
        return None
    key = (dev.index or 0, n, world, gkey)
    ent = _P2P.get(key)
    if ent is None and len(_P2P) >= _P2P_MAX:
        return None          # the buffers are per catalog-2 size and never freed: further sizes merge through NCCL
    if ent is None:
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            import torch.distributed._symmetric_memory as symm
            g = group if group is not None else dist.group.WORLD
            exch = symm.empty(2*n + 1, dtype=torch.float64, device=dev)
            out = symm.empty(2*n, dtype=torch.float64, device=dev)
            hx, ho = symm.rendezvous(exch, g), symm.rendezvous(out, g)
            px = (C.c_void_p*world)(*[int(q) for q in hx.buffer_ptrs])
            po = (C.c_void_p*world)(*[int(q) for q in ho.buffer_ptrs])
            ent = (exch, out, hx, ho, px, po, torch.cuda.Stream(device=dev))
        except Exception as e:           # noqa: BLE001 -- any failure means "no peer mappings here"
            if dist.get_rank(group) == 0:
                print("[vif_b200] symmetric memory unavailable (%s): NCCL merge" % (str(e).splitlines() or ["?"])[0],
                      flush=True)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # all ranks take the same path
        if int(ok.item()) == 0:
            _P2P_OFF.add(gkey)
            return None
        _P2P[key] = ent
    return ent


def _merge_reverse_cuda(rid, rd, row_offset, linear, world, group):
    """In place, three fused kernels of the library around the two all-reduces
    (xm_rev_merge_pack / _mask / _unpack, include/vif_b200_qxmatch.h)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _lib
    lib = _lib.load()
    n = rid.numel()
    dev = rid.device.index or 0
    stream = torch.cuda.current_stream(rid.device).cuda_stream
    err = C.create_string_buffer(512)
    key = torch.empty_like(rd)

    def check(rc):
        if rc != 0:
            raise RuntimeError(err.value.decode())

    # The library launches on its own (non-blocking) stream when torch's current stream is the
    # legacy default stream (handle 0 == "NULL"), so order the steps on the host: after an
    # all-reduce, synchronising the current stream guarantees the collective has landed.
    cur = torch.cuda.current_stream(rid.device)
    check(lib.xm_rev_merge_pack(rid.data_ptr(), rd.data_ptr(), n, row_offset, key.data_ptr(), dev, stream, err, 512))
    best = key.clone()
    if world > 1:
        dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
    cur.synchronize()
    check(lib.xm_rev_merge_mask(rid.data_ptr(), key.data_ptr(), best.data_ptr(), n, dev, stream, err, 512))
    if world > 1:
        dist.all_reduce(rid, op=dist.ReduceOp.MIN, group=group)
    cur.synchronize()
    unmatched = float("inf") if linear else float("nan")
    check(lib.xm_rev_merge_unpack(rid.data_ptr(), rd.data_ptr(), best.data_ptr(), n, unmatched, dev, stream, err, 512))


def _merge_reverse_torch(rid, rd, row_offset, world, group):
    """Same rule with plain torch ops (CPU tensors: the gloo tests)."""
    import torch
    import torch.distributed as dist
    big = torch.iinfo(torch.int64).max
    rid = torch.where(rid < 0, torch.full_like(rid, big), rid + row_offset)
    rd_key = torch.where(torch.isnan(rd), torch.full_like(rd, float("inf")), rd)
    if world > 1:
        best = rd_key.clone()
        dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
        rid = torch.where(rd_key == best, rid, torch.full_like(rid, big))
        dist.all_reduce(rid, op=dist.ReduceOp.MIN, group=group)
        # distances of unmatched slots stay NaN (spherical) / inf (linear) as in the reference
        rd = torch.where(rid == big, rd, best)
    rid = torch.where(rid == big, torch.full_like(rid, -1), rid)
    return rid, rd
