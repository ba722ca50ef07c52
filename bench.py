#!/usr/bin/env python
"""bench.py -- the headline metric of BASELINE.json: qxmatch sources/sec on config C2
(1e7 x 1e7 synthetic uniform sources over RA[0,10) x Dec[-5,5), nth=1, cell-binned, reverse match on).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (the reference's own CPU qxmatch)

One "step" = one complete qxmatch of the workload. Prints ONE JSON line on rank 0.
  value        whole-job sources/s with inputs resident in HBM (device pipeline only)
  e2e          the same through the host-buffer C-ABI call (xm_qxmatch_f64): H2D of the catalogs
               from pinned memory and D2H of id/d/rid/rd inside the timed region
  roofline     whole-path algorithmic bytes / step time against the measured HBM peak, with
               the dominant kernel (nearest-neighbour filter, live CUDA-event time) beside it; --config c4:
               11 flop/pair against the measured DFMA peak (profiles/fp64_peak.json)
  cpu_baseline the reference (oracle/_ref) or its port timed on the host cores: the FULL workload once on
               all threads (N = 1), a bounded sample on one thread
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "qxmatch_sources_per_sec"
UNIT = "sources/s"
N_FULL = 10_000_000
SAMPLE_N = 250_000          # one-thread CPU sample: sub-box with 1/40 of the area, same source density
REF_ARM_N = 1_000_000       # reference arm, per step: sub-box with 1/10 of the area (a few seconds on 16 cores)
C4_N = 1_000_000


def workload_desc(n1, n2, config="c2"):
    if config == "c4":
        return ("C4: brute_force qxmatch %.0e x %.0e synthetic sources uniform on the full sky, nth=1, reverse match "
                "(rid/rd) on" % (n1, n2))
    if config == "c3":
        return ("C3: self match (self=true) nth=5 on a %.0e-source clustered (Soneira-Peebles) synthetic catalog over "
                "RA[0,10]xDec[-5,5], cell-binned" % n1)
    if config == "c5":
        return ("C5: qxmatch %.0e x %.0e synthetic sources uniform on the full sky, nth=1, binned call answered by "
                "the exact spherical grid search, reverse match (rid/rd) on" % (n1, n2))
    return ("C2: qxmatch %.0e x %.0e synthetic uniform sources over RA[0,10)xDec[-5,5) (100 deg^2), nth=1, "
            "cell-binned, reverse match (rid/rd) on" % (n1, n2))


def gen_catalog(G, config, seed, n, dev, start=0):
    """Rows [start, start+n) of a synthetic catalog on the device, generated in chunks."""
    import torch
    ra = torch.empty(n, dtype=torch.float64, device=dev)
    dec = torch.empty(n, dtype=torch.float64, device=dev)
    step = 50_000_000
    for s0 in range(0, n, step):
        m = min(step, n - s0)
        if config in ("c5", "c4"):
            a, b = G.fullsky(seed, m, xp="torch", device=dev, start=start + s0)
        else:
            a, b = G.c2(seed, m, xp="torch", device=dev, start=start + s0)
        ra[s0:s0 + m] = a
        dec[s0:s0 + m] = b
    return ra, dec


# ---- clocks sampling --------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons read through NVML at three points INSIDE the timed region, by a
    helper thread woken for each read (a free-running polling thread or an `nvidia-smi -lms` child
    contends with the CUDA driver and was measured to slow a 4 ms step by 40 %)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.idx, self.rows, self.h, self.nv, self.max_mhz = gpu_index, [], None, None, 0.0
        self.thread, self.wake, self.quit = None, None, False

    def start(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.nv = pynvml
            # nvmlDeviceGetMaxClockInfo takes ~2.4 ms per call on this driver: read it once, outside
            # the timed region; the per-sample queries (current clock, reasons) take microseconds
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.sample(keep=False)
            # The queries run on a helper thread that sleeps until the timed loop asks for a sample:
            # an NVML query is microseconds as a rule but was seen to block its caller for 7-30 ms
            # (more often with several ranks inside the driver); the timed loop never waits for it.
            import threading
            self.wake = threading.Event()
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _loop(self):
        while True:
            self.wake.wait()
            self.wake.clear()
            if self.quit:
                return
            self.sample()

    def request(self):
        """Asks the helper thread for one sample; returns at once."""
        if self.wake is not None:
            self.wake.set()

    def sample(self, keep=True):
        nv = self.nv
        if nv is None:
            return
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            mx = self.max_mhz
            fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            rs = fn(self.h)
            if keep:
                self.rows.append((float(sm), float(mx), int(rs)))
        except Exception:
            pass

    def stop(self):
        if self.thread is not None:
            self.quit = True
            self.wake.set()
            self.thread.join(timeout=2.0)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"], "samples": 0}
        reasons = sorted(n for n, bit in self.REASONS.items() if any(r[2] & bit for r in self.rows))
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": max(r[1] for r in self.rows),
                "reasons": reasons, "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ---- CPU side (reference arm / cpu_baseline) -----------------------------------------------------
def cpu_sample(n=None):
    """Bounded sample of the C2 workload: both catalogs restricted to a sub-box holding n (default
    SAMPLE_N) rows at the full workload's density (same cell occupancy => same per-source cost)."""
    from vif_b200 import catalogs as G
    n = n or SAMPLE_N
    frac = n / N_FULL
    side = 10.0 * np.sqrt(frac)
    a1, b1 = G.box(1, n, 0.0, side, -5.0, side)
    a2, b2 = G.box(2, n, 0.0, side, -5.0, side)
    return a1, b1, a2, b2, ("%d x %d rows in a %.3f x %.3f deg sub-box (same density as C2, 1/%d of its area)"
                            % (n, n, side, side, round(1 / frac)))


def cpu_runner(threads=None):
    import oracle as O
    cores = threads or os.cpu_count() or 1
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libqxmatch_ref.so")
    if os.path.exists(ref_so) or os.path.isdir("/root/reference/include/vif"):
        try:
            O.ref_qxmatch(np.array([0.0, 0.01]), np.array([0.0, 0.01]), np.array([0.001, 0.009]), np.array([0.002, 0.008]))
            return (lambda a1, b1, a2, b2: O.ref_qxmatch(a1, b1, a2, b2, thread=cores)), "reference", cores
        except Exception:
            pass
    return (lambda a1, b1, a2, b2: O.oracle_qxmatch(a1, b1, a2, b2, thread=cores)), "port", cores


def measure_cpu_baseline(budget_s=12.0, single_thread=True, full=False):
    """The reference's CPU qxmatch on the host cores. full=True (bench default, N = 1): the WHOLE C2
    workload (1e7 x 1e7) once on all host threads -- the same configuration the GPU number is quoted
    on (33 s on the pool's 16-core hosts). Otherwise a bounded sample, up to three repetitions inside
    `budget_s`. Then the sample once on one thread (SURVEY.md section 8d asks for both)."""
    from vif_b200 import catalogs as G
    a1, b1, a2, b2, desc = cpu_sample()
    fn, kind, cores = cpu_runner()
    fn(a1[:20000], b1[:20000], a2[:20000], b2[:20000])
    if full:
        f1, g1 = G.c2(1, N_FULL)
        f2, g2 = G.c2(2, N_FULL)
        t0 = time.perf_counter()
        fn(f1, g1, f2, g2)
        cdt = time.perf_counter() - t0
        out = {"value": N_FULL / cdt, "unit": UNIT, "cores": cores, "kind": kind, "seconds": cdt,
               "sample": "the full workload (%d x %d rows), one run on all host threads" % (N_FULL, N_FULL)}
        del f1, g1, f2, g2
    else:
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 and time.perf_counter() - t0 < budget_s:
            fn(a1, b1, a2, b2)
            reps += 1
        cdt = (time.perf_counter() - t0) / reps
        out = {"value": SAMPLE_N / cdt, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
    if single_thread and cores > 1:
        # a bounded sample (a smaller one is cache-friendlier: 50 000 rows run 2.8x faster per source), once
        fn1, _, _ = cpu_runner(threads=1)
        t0 = time.perf_counter()
        fn1(a1, b1, a2, b2)
        out["value_1_thread"] = SAMPLE_N / (time.perf_counter() - t0)
        out["sample_1_thread"] = desc
    return out


def run_reference_arm(args):
    """The reference's own CPU qxmatch (oracle/_ref: the unmodified header compiled here) on all host
    threads. Each step is a bounded sample of C2 (REF_ARM_N rows per catalog in a sub-box of the same
    density) so that K + W steps end within minutes; the whole 1e7 x 1e7 workload is then run ONCE and
    its figure reported beside the sample's (`full_size`; XM_BENCH_NO_FULL=1 skips it)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = min(args.n or REF_ARM_N, REF_ARM_N)
    a1, b1, a2, b2, desc = cpu_sample(n)
    fn, kind, cores = cpu_runner()
    for _ in range(args.warmup):
        fn(a1, b1, a2, b2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(a1, b1, a2, b2)
    dt = (time.perf_counter() - t0) / args.steps
    v = n / dt
    full = None
    if not os.environ.get("XM_BENCH_NO_FULL") and n == REF_ARM_N:
        from vif_b200 import catalogs as G
        f1, g1 = G.c2(1, N_FULL)
        f2, g2 = G.c2(2, N_FULL)
        t0 = time.perf_counter()
        fn(f1, g1, f2, g2)
        fdt = time.perf_counter() - t0
        full = {"value": N_FULL / fdt, "unit": UNIT, "seconds": fdt, "rows": [N_FULL, N_FULL], "runs": 1}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_desc(N_FULL * (args.gpus if args.scaling == "weak" else 1), N_FULL), "nth": 1,
                   "sample": desc},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "full_size": full,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---- GPU arm ----------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from vif_b200 import catalogs as G
    from vif_b200 import qxmatch
    from vif_b200.qxmatch import QxmatchParams
    from vif_b200.distributed import qxmatch_sharded, qxmatch_partitioned, qxmatch_banded, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries ONE JSON line: NCCL's own log (NCCL_DEBUG, whatever the launcher set) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    # weak scaling (c2 default): every rank brings args.n rows of catalog 1, catalog 2 stays args.n rows
    n1 = args.n * world if args.scaling == "weak" else args.n
    n2 = args.n2 if args.n2 else (args.n // 10 if args.config == "c5" else args.n)
    lo, hi = shard_bounds(n1, world, rank)
    if args.config == "c3":
        levels = len(str(n1)) - 1
        assert 10 ** levels == n1, "c3 needs a power of ten"
        fra, fdec = G.c3_clustered(3, levels, xp="torch", device=dev)      # prefix-indexed generator: whole, then slice
        ra1, dec1 = fra[lo:hi].clone(), fdec[lo:hi].clone()
        del fra, fdec
        torch.cuda.empty_cache()
        ra2, dec2 = ra1, dec1
        n2 = n1
    else:
        ra1, dec1 = gen_catalog(G, args.config, 1, hi - lo, dev, start=lo)  # this rank's rows of catalog 1
    if args.config == "c3":
        pass
    elif rank == 0:
        ra2, dec2 = gen_catalog(G, args.config, 2, n2, dev)
    else:
        ra2 = torch.empty(n2, dtype=torch.float64, device=dev)
        dec2 = torch.empty(n2, dtype=torch.float64, device=dev)
    params = QxmatchParams(nth=5 if args.config == "c3" else 1, device=local, brute_force=(args.config == "c4"))
    stats_acc = {"ms_forward": 0.0, "ms_reverse": 0.0, "ms_search": 0.0, "ms_total": 0.0, "ms_index": 0.0,
                 "ms_bbox": 0.0, "ms_filter_fwd": 0.0, "ms_filter_rev": 0.0, "launches": 0, "evals": 0, "steps": 0,
                 "refined": 0}

    def step(record):
        from vif_b200.qxmatch import last_stats
        if args.config == "c3" and world > 1:
            out = qxmatch_partitioned(ra1, dec1, None, None, lo, n1, params, self_match=True)
        elif args.config == "c3":
            r = qxmatch(ra1, dec1, params=params)           # one-catalog overload = self match
            out = (r.id, r.d, r.rid, r.rd)
        elif world == 1:
            r = qxmatch(ra1, dec1, ra2, dec2, params=params)
            out = (r.id, r.d, r.rid, r.rd)
        elif args.config == "c2" and args.scaling == "strong":
            # fixed 1e7 x 1e7: catalogs small enough to replicate, row blocks of BOTH per rank, no merge
            out = qxmatch_partitioned(ra1, dec1, ra2, dec2, lo, n1, params)
        elif args.config == "c5" and not os.environ.get("XM_C5_SHARDED"):
            # both directions sharded: declination bands, one all-to-all of catalog 1 over NVLink
            fwd, _, rev = qxmatch_banded(ra1, dec1, ra2, dec2, lo, params, home=bool(os.environ.get("XM_C5_HOME")))
            out = (fwd, rev)
        else:
            # reverse results stay sliced over the ranks like the forward ones (XM_C2_REPLICATE=1: on every rank)
            out = qxmatch_sharded(ra1, dec1, ra2, dec2, lo, params,
                                  replicate_reverse=bool(os.environ.get("XM_C2_REPLICATE")))
        if record:
            st = last_stats(local)
            for k in ("ms_forward", "ms_reverse", "ms_search", "ms_total", "ms_index", "ms_bbox", "ms_filter_fwd",
                      "ms_filter_rev"):
                stats_acc[k] += st[k]
            stats_acc["refined"] += st["rows_refined_fwd"] + st["rows_refined_rev"]
            stats_acc["launches"] += st["launches"]
            stats_acc["evals"] += st["evals_fwd"] + st["evals_rev"]
            stats_acc["steps"] += 1
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("XM_BENCH_NO_CLOCKS"):
        sampler.start()
    out = None
    for _ in range(max(args.warmup, 3)):
        out = step(False)       # kept alive like in the timed loop: torch's allocator ends up caching both result sets
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sample_at = {args.steps // 4, args.steps // 2, (3 * args.steps) // 4}      # clock samples inside the timed region
    e0.record()
    for k in range(args.steps):
        if rank == 0 and k in sample_at:
            sampler.request()               # read by the helper thread while this step runs on the GPU
        out = step(True)
        marks[k].record()
    e1.record()
    barrier()
    per_step = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    args.step_ms = {"min": min(per_step), "median": float(np.median(per_step)), "max": max(per_step),
                    "argmax": int(np.argmax(per_step))}
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    launches = torch.tensor([stats_acc["launches"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms.item() / args.steps
    value = n1 / (ms_step * 1e-3)

    # ---- end to end through the host-buffer call -------------------------------------------------
    if args.no_e2e:
        e2e_value, e2e_steps, h2d, d2h = None, 0, 0, 0
        dt = torch.tensor([0.0], dtype=torch.float64, device=dev)
    else:
        e2e_value, e2e_steps, h2d, d2h, dt = run_e2e(args, world, rank, lo, hi, n1, n2, ra1, dec1, ra2, dec2, params,
                                                     dev, barrier, qxmatch, qxmatch_sharded)
    # ---- BASELINE config 5 (1e9 x 1e8 full sky) beside the headline line: strong scaling over the same N ----
    args.c5_strong = None
    if args.config == "c2" and n1 == N_FULL * (world if args.scaling == "weak" else 1) and not os.environ.get("XM_BENCH_NO_C5"):
        del ra1, dec1, ra2, dec2, out
        args.c5_strong = run_c5_strong(world, rank, local, dev, barrier)
    if rank == 0:
        emit(args, world, n1, n2, lo, hi, value, ms_step, launches, clocks, stats_acc, e2e_value, e2e_steps, h2d, d2h, dt)
    if world > 1:
        dist.destroy_process_group()


def run_c5_strong(world, rank, local, dev, barrier, n1=1_000_000_000, n2=100_000_000, steps=3):
    """C5 at its full size on the N GPUs of this run: one GPU answers the binned call with the exact spherical
    search; N > 1 regroup catalog 1 into declination bands (qxmatch_banded: both directions sharded, results stay
    with the band). Device-resident inputs, CUDA-event time of `steps` calls after 2 warm-ups, max over ranks."""
    import torch
    import torch.distributed as dist
    from vif_b200 import catalogs as G, _lib
    from vif_b200 import qxmatch
    from vif_b200.qxmatch import QxmatchParams
    from vif_b200.distributed import qxmatch_banded, shard_bounds
    try:
        _lib.load().xm_release_memory(-1)
        torch.cuda.empty_cache()
        lo, hi = shard_bounds(n1, world, rank)
        need = (hi - lo) * 125 + n2 * 110                    # inputs, results, index and search scratch [bytes]
        ok = torch.tensor([1 if torch.cuda.mem_get_info(dev)[0] > need else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)         # every rank takes the same decision
        if int(ok.item()) == 0:
            return {"skipped": "needs %.0f GB of free HBM per GPU" % (need / 1e9)}
        ra1, dec1 = gen_catalog(G, "c5", 1, hi - lo, dev, start=lo)
        if rank == 0:
            ra2, dec2 = gen_catalog(G, "c5", 2, n2, dev)
        else:
            ra2 = torch.empty(n2, dtype=torch.float64, device=dev)
            dec2 = torch.empty(n2, dtype=torch.float64, device=dev)
        params = QxmatchParams(device=local)

        def step():
            if world == 1:
                return qxmatch(ra1, dec1, ra2, dec2, params=params)
            return qxmatch_banded(ra1, dec1, ra2, dec2, lo, params, home=False)
        for _ in range(2):
            keep = step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            keep = step()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        del keep, ra1, dec1, ra2, dec2
        _lib.load().xm_release_memory(-1)
        torch.cuda.empty_cache()
        return {"workload": workload_desc(n1, n2, "c5"), "n_gpus": world, "steps": steps, "ms_per_step": ms.item(),
                "value": n1 / (ms.item() * 1e-3), "unit": UNIT, "scaling": "strong",
                "bytes": 16 * n1 + 16 * n2 + 16 * n1 + 16 * n2,
                "hbm_GBps_whole_path": (16 * n1 + 16 * n2 + 16 * n1 + 16 * n2) / (ms.item() * 1e-3) / 1e9,
                "parallelism": ("1 rank" if world == 1 else
                                "declination bands, one NCCL all-to-all of catalog 1, results stay with the band")}
    except Exception as e:          # noqa: BLE001 -- the headline line must not depend on this extra
        return {"error": (str(e).splitlines() or ["?"])[0][:300]}


def run_e2e(args, world, rank, lo, hi, n1, n2, ra1, dec1, ra2, dec2, params, dev, barrier, qxmatch, qxmatch_sharded):
    import torch
    import torch.distributed as dist
    h = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in
         (("ra1", ra1), ("dec1", dec1), ("ra2", ra2), ("dec2", dec2))}
    if world > 1:
        dist.broadcast(ra2, src=0)
        dist.broadcast(dec2, src=0)
    for k, v in (("ra1", ra1), ("dec1", dec1), ("ra2", ra2), ("dec2", dec2)):
        h[k].copy_(v)
    nloc = hi - lo
    o = {"id": torch.empty((1, nloc), dtype=torch.int64).pin_memory(), "d": torch.empty((1, nloc), dtype=torch.float64).pin_memory(),
         "rid": torch.empty(n2, dtype=torch.int64).pin_memory(), "rd": torch.empty(n2, dtype=torch.float64).pin_memory()}
    if world == 1:
        onp = {"id": o["id"].numpy().view(np.uint64), "d": o["d"].numpy(), "rid": o["rid"].numpy().view(np.uint64),
               "rd": o["rd"].numpy()}
        hn = {k: v.numpy() for k, v in h.items()}

        def e2e_step():
            qxmatch(hn["ra1"], hn["dec1"], hn["ra2"], hn["dec2"], params=params, out=onp)
        h2d = 16 * n1 + 16 * n2
        d2h = 16 * n1 + 16 * n2
    else:
        def e2e_step():
            a = h["ra1"].to(dev, non_blocking=True)
            b = h["dec1"].to(dev, non_blocking=True)
            if rank == 0:
                ra2.copy_(h["ra2"], non_blocking=True)
                dec2.copy_(h["dec2"], non_blocking=True)
            if args.config == "c2" and args.scaling == "strong":
                from vif_b200.distributed import qxmatch_partitioned
                i_, d_, rid_, rd_, _ = qxmatch_partitioned(a, b, ra2, dec2, lo, n1, params)
                o["rid"][:rid_.numel()].copy_(rid_, non_blocking=True)      # every rank returns its block
                o["rd"][:rd_.numel()].copy_(rd_, non_blocking=True)
            else:
                i_, d_, rid_, rd_ = qxmatch_sharded(a, b, ra2, dec2, lo, params,
                                                    replicate_reverse=bool(os.environ.get("XM_C2_REPLICATE")))
                if rid_.numel() < n2 or rank == 0:       # sliced: every rank returns its block of catalog-2 rows
                    o["rid"][:rid_.numel()].copy_(rid_, non_blocking=True)
                    o["rd"][:rd_.numel()].copy_(rd_, non_blocking=True)
            o["id"].copy_(i_, non_blocking=True)
            o["d"].copy_(d_, non_blocking=True)
            torch.cuda.synchronize()
        h2d = 16 * n1 + 16 * n2
        d2h = 16 * n1 + 16 * n2
    e2e_steps = max(1, min(args.steps, 10))
    args.e2e_pageable = None
    if world == 1 and args.config == "c2":
        # the same call on ordinary (pageable) numpy arrays -- what the drop-in header passes (std::vector
        # storage): the library moves them through its chunked pinned staging pipeline (xm_staging.cu)
        pn = {k: np.array(v) for k, v in hn.items()}
        po = {k: np.empty_like(v) for k, v in onp.items()}
        for v in po.values():
            v.fill(0)
        for _ in range(2):
            qxmatch(pn["ra1"], pn["dec1"], pn["ra2"], pn["dec2"], params=params, out=po)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            qxmatch(pn["ra1"], pn["dec1"], pn["ra2"], pn["dec2"], params=params, out=po)
        pdt = (time.perf_counter() - t0) / e2e_steps
        args.e2e_pageable = {"value": n1 / pdt, "unit": UNIT, "ms_per_step": pdt * 1e3, "steps": e2e_steps,
                             "buffers": "pageable host memory in and out (chunked pinned staging, %s host threads)"
                                        % (os.environ.get("XM_STAGE_THREADS") or "default")}
        del pn
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return n1 / dt.item(), e2e_steps, h2d, d2h, dt


def fp64_peak():
    """Measured DFMA peak of this pool's B200 (scripts/fp64_peak.py -> profiles/fp64_peak.json), else nominal."""
    p = os.path.join(ROOT, "profiles", "fp64_peak.json")
    try:
        return float(json.load(open(p))["tflops"]), "measured (profiles/fp64_peak.json)"
    except Exception:
        return 37.2, "nominal (148 SM x 64 DFMA/clk x 2 x 1.965 GHz)"


def parallelism_desc(args, world, hi, lo):
    if world == 1:
        return "1 rank"
    if args.config == "c2" and args.scaling == "strong":
        return ("catalog 2 NCCL-broadcast, catalog-1 row shards NCCL-all-gathered; each of the %d ranks matches its row "
                "block of catalog 1 (forward) and of catalog 2 (reverse) against the whole other catalog; no merge "
                "collective" % world)
    if args.config == "c5" and not os.environ.get("XM_C5_SHARDED"):
        return ("catalog 2 NCCL-broadcast; catalog-1 rows (%d per rank) regrouped into %d equal-area declination bands by one "
                "NCCL all-to-all over NVLink (order-preserving partition kernel, halo rows to both sides); rank b matches "
                "band b in both directions; results stay with the band, keyed by global row (XM_C5_HOME=1: forward results "
                "return to the rows' owners by a second all-to-all)" % (hi - lo, world))
    return ("catalog-1 rows sharded over %d ranks (%d rows each); bounding boxes of both catalogs by one 8-double "
            "all-reduce, then catalog 2 NCCL-broadcast while the local rows of catalog 1 are indexed; catalog 2 indexed "
            "per rank; reverse match merged by one kernel over NVLink peer memory (xm_rev_merge_p2p; NCCL "
            "all-reduce(min) when peer mappings are unavailable), merged rows %s"
            % (world, hi - lo, "replicated on every rank" if os.environ.get("XM_C2_REPLICATE") else
               "left sliced over the ranks by catalog-2 row block, as the forward results are by catalog-1 row block"))


def emit(args, world, n1, n2, lo, hi, value, ms_step, launches, clocks, stats_acc, e2e_value, e2e_steps, h2d, d2h, dt):
    peaks, how = measured_peaks()
    ns = max(stats_acc["steps"], 1)
    avg = {k: stats_acc[k] / ns for k in ("ms_total", "ms_bbox", "ms_index", "ms_search", "ms_forward", "ms_reverse",
                                           "ms_filter_fwd", "ms_filter_rev")}
    nth = 5 if args.config == "c3" else 1
    mirror = args.config != "c3"
    # whole-path algorithmic bytes, SURVEY.md section 8(d): coordinates in, (id, d) out, (rid, rd) out
    whole_bytes = 16 * n1 + (16 * n2 if args.config != "c3" else 0) + 16 * nth * n1 + (16 * n2 if mirror else 0)
    whole_gbps = whole_bytes / (ms_step * 1e-3) / 1e9
    if args.config == "c4":
        peak, peak_how = fp64_peak()
        tfl = 11.0 * n1 * n2 / (ms_step * 1e-3) / 1e12
        roof = {"bound": "fp64", "kernel": "brute_filter_kernel (forward pass + reverse pass) + brute_refine_kernel",
                "achieved": tfl, "peak": peak, "unit": "TFLOP/s", "frac": tfl / peak, "traffic": None,
                "peak_source": peak_how,
                "convention": "11 FP64 flop per (i, j) pair of SURVEY.md 8(d), n1*n2 pairs (the reference evaluates a "
                              "pair once for both directions); the kernels evaluate every pair once per direction as a "
                              "3-term dot product (1 DMUL + 2 DFMA + 1 DSETP)",
                "dfma_issue": {"pairs_per_s_per_direction": n1 * n2 * 2.0 / (avg["ms_search"] * 1e-3) / 2.0
                               if avg["ms_search"] > 0 else None,
                               "fp64_instr_tflops_equiv": (2.0 * n1 * n2 * 5.0 / (avg["ms_search"] * 1e-3) / 1e12
                                                           if avg["ms_search"] > 0 else None),
                               "note": "2 passes x n1*n2 pairs x (1 DMUL + 2 DFMA = 5 flop) / search time; ncu "
                                       "sm__inst_executed_pipe_fp64 of the same kernel: profiles/r2_brute_filter_ncu.txt"},
                "stage_ms": avg}
    else:
        # dominant kernel: the nearest-neighbour filter (forward + reverse launches of the same kernel, run
        # concurrently on two streams). Algorithmic bytes per launch: source coordinates in, candidate
        # coordinates in, (id, d) out -- DESIGN.md section 5.
        nsrc_f = hi - lo
        bytes_fwd = 16 * nsrc_f + 16 * n2 + 16 * nth * nsrc_f
        bytes_rev = (16 * n2 + 16 * nsrc_f + 16 * n2) if mirror else 0
        ms_s = avg["ms_search"]
        ms_k = avg["ms_filter_fwd"] + avg["ms_filter_rev"]          # the two launches' own CUDA-event times
        if args.config == "c2" and ms_k > 0:
            # bytes per launch / average launch duration (the persistent launches run one after the other)
            k_ach = (bytes_fwd + bytes_rev) / (ms_k * 1e-3) / 1e9
            k_note = ("achieved = algorithmic bytes per launch / the kernel's average launch duration (CUDA events on "
                      "the launching streams; the forward and the reverse launch run one after the other, the reverse "
                      "one beside the forward lane's unpack and exact-list kernels); search_stage_GBps = the same bytes "
                      "over the whole search stage (filter + unpack + exact list, both lanes); traffic = dram bytes of "
                      "ONE launch from ncu, profiles/ring_search_traffic.json")
        else:
            k_ach = (bytes_fwd + bytes_rev) / (ms_s * 1e-3) / 1e9 if ms_s > 0 else 0.0
            k_note = ("achieved = bytes of both launches / time the two search lanes are in flight (they overlap; each "
                      "launch's own CUDA-event time is listed where it was taken)")
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ring_search_traffic.json")
        if os.path.exists(tp) and args.config == "c2" and world == 1 and n1 == N_FULL:
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        kname = {"c5": "sphere_nn_filter_kernel + list-mode sphere_search_kernel (forward and reverse lanes, concurrent)",
                 "c3": "ring_topk_filter_kernel (nth = 5) + list-mode ring_search_exact_kernel"}.get(
                     args.config, "ring_nn32p_kernel (persistent FP32 nearest-neighbour filter; forward launch, then reverse launch)")
        roof = {"bound": "hbm", "achieved": whole_gbps, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": whole_gbps / peaks["hbm_gbs"], "traffic": traffic, "peak_source": how,
                "what": "whole path, SURVEY.md 8(d): B = 16 n1 + 16 n2 + 16 nth n1 + 16 n2 bytes / step time",
                "bytes": whole_bytes,
                "dominant_kernel": {"kernel": kname, "achieved": k_ach, "frac": k_ach / peaks["hbm_gbs"], "unit": "GB/s",
                                    "algorithmic_bytes_per_launch": {"forward": bytes_fwd, "reverse": bytes_rev},
                                    "ms": {"forward_launch": avg["ms_filter_fwd"], "reverse_launch": avg["ms_filter_rev"],
                                           "search_stage": ms_s},
                                    "search_stage_GBps": (bytes_fwd + bytes_rev) / (ms_s * 1e-3) / 1e9 if ms_s > 0 else None,
                                    "note": k_note},
                "stage_ms": avg,
                "distance_evals_per_step": stats_acc["evals"] / ns,
                "rows_to_exact_kernel_per_step": stats_acc["refined"] / ns,
                # the limit that actually binds is issue/latency, not bytes: evaluations by the 11-flop convention
                "fp64": {"achieved_tflops": 11.0 * stats_acc["evals"] / ns / (ms_s * 1e-3) / 1e12 if ms_s > 0 else 0.0,
                         "peak_tflops": fp64_peak()[0], "peak_source": fp64_peak()[1],
                         "note": "the filter evaluates in FP32 (4 FFMA/FMUL per pair); FP64 only for the winner"}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_desc(n1, n2, args.config), "n1": n1, "n2": n2, "nth": nth,
                   "l2": ("no flush: the input coordinates plus the cell index exceed the 126 MB L2" if args.config != "c4" else
                          "no flush: the kernel is FP64-pipe bound; its candidate tiles are staged by TMA from L2"),
                   "parallelism": parallelism_desc(args, world, hi, lo)},
        "e2e": ({"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                 "ms_per_step": dt.item() * 1e3, "steps": e2e_steps} if e2e_value is not None else None),
        "e2e_pageable": getattr(args, "e2e_pageable", None),
        "gpu_launches": int(launches.item()),
        "step_ms_rank0": args.step_ms,
        "clocks": clocks,
        "roofline": roof,
    }
    if getattr(args, "c5_strong", None) is not None:
        line["c5_strong"] = args.c5_strong
    if world == 1 and not args.no_cpu and args.config == "c2":
        line["cpu_baseline"] = measure_cpu_baseline(full=(n1 == N_FULL and not os.environ.get("XM_BENCH_NO_FULL")))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", dest="n", type=int, default=None, help="rows of catalog 1 (default: the full workload)")
    ap.add_argument("--rows2", dest="n2", type=int, default=None, help="rows of catalog 2 (default: n for c2, n/10 for c5)")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 = the headline metric's workload (default); c3 = self nth=5 clustered; c4 = brute force "
                         "1e6 x 1e6 full sky (FP64 roofline); c5 = 1e9 x 1e8 full sky")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="N>1: weak = every rank brings --rows rows of catalog 1 against the same catalog 2 (c2 default: "
                         "the path shards by catalog-1 rows); strong = the total stays --rows (c3/c5 default)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.n is None:
        args.n = {"c5": 1_000_000_000, "c3": 100_000_000, "c4": C4_N}.get(args.config, N_FULL)
    if args.scaling == "auto":
        args.scaling = "weak" if args.config == "c2" else "strong"
    if args.config in ("c3", "c4"):
        args.no_e2e = args.no_e2e or args.config == "c3"
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
