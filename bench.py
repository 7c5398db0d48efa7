#!/usr/bin/env python
"""bench.py -- BIGSI search hot path on B200 (BASELINE.json configs[1], weak-scaled by column shard).

Workload (config.workload): synthetic m=25 000 000, h=3, k=31 index with 50 000 sample columns
PER GPU resident in HBM (156.8 GB/GPU); one STEP = one 10 000-k-mer query run through the hot
path: canonical+murmur3 hash kernel -> fused gather-AND-popcount kernel -> merge -> threshold at
min_kmers = U (an exact query through the count path).  64 distinct queries rotate so that
consecutive steps never touch the same rows (187.5 MB of rows per step > 126 MB L2).

Metric: k-mer row-AND lookups/s, one lookup = gather h rows of one 50 000-column shard and AND
them (18 750 algorithmic bytes).  At N GPUs every rank looks the same k-mers up in its own
column shard (rank 0 broadcasts the row ids, hits are all-gathered), so the whole-job value is
N * U * steps / time.

`--impl reference` times the CPU oracle port of the reference's algorithm (oracle/, OpenMP on all
host cores) on the same workload; see DESIGN.md "Measurement".
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K, H = 31, 3
METRIC = "kmer_row_and_lookups_per_sec"
UNIT = "lookups/s (1 lookup = h=3 row gather+AND over one 50 000-column shard = 18 750 B)"
N_DISTINCT = 64
HIT_CAP = 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=25_000_000, help="Bloom filter size (rows); default = BASELINE")
    ap.add_argument("--cols", type=int, default=50_000, help="sample columns per GPU")
    ap.add_argument("--kmers", type=int, default=10_000, help="unique k-mers per query")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prewarm", type=int, default=300, help="untimed steps before the warm-up (clock ramp)")
    ap.add_argument("--exchange", default="pipelined", choices=["pipelined", "fused", "nccl"],
                    help="N>1: query broadcast + hit all-gather inside the query kernel over NVLink peer memory, "
                         "pipelined (the kernel of query s waits for the shards' hits of query s-1 only; default) or "
                         "lock-step (fused), or as two NCCL collectives per query (nccl)")
    ap.add_argument("--opt", action="append", default=[], help="diagnostics: index option key=value (repeatable)")
    ap.add_argument("--timeline", action="store_true",
                    help="diagnostics: after the timed region print every rank's per-CTA kernel timeline (stderr)")
    return ap.parse_args()


def planted_columns(n_shards, cols):
    """Per shard: 3 all-ones columns and 4 graded columns, as GLOBAL column ids."""
    pc, pt = [], []
    for g in range(n_shards):
        base = g * cols
        for c in (0, 1, cols - 1):
            pc.append(base + c)
            pt.append(0xFFFFFFFF)
        for c, d in ((cols // 2, 0.95), (7, 0.6), (cols // 4 + 1, 0.41), (2 * cols // 3, 0.3)):
            pc.append(base + c)
            pt.append(int(d * 2 ** 32))
    return pc, pt


def make_queries(n, u):
    """n distinct queries of u unique random 31-mers (seed q+1): uint8 [n, u, K]."""
    out = np.empty((n, u, K), dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for q in range(n):
        rng = np.random.default_rng(q + 1)
        arr = acgt[rng.integers(0, 4, size=(u, K))]
        assert len(np.unique(arr, axis=0)) == u  # set(kmers) semantics: all unique already
        out[q] = arr
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, device_index, period=0.004):
        super().__init__(daemon=True)
        self.period = period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ---------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------
class CpuArm:
    """The oracle port of the reference's algorithm on the same synthetic workload.  Rows the
    queries touch are regenerated from the synthetic index's pure function BEFORE timing (the
    reference's in-memory store would hold them already)."""

    def __init__(self, args, n_queries):
        from oracle import oracle as O

        self.O = O
        self.cores = O.lib().oracle_num_threads()
        pc, pt = planted_columns(1, args.cols)
        self.spec = O.SynthSpec(0, 1, pc, pt)
        self.m, self.cols, self.u = args.m, args.cols, args.kmers
        self.queries = make_queries(n_queries, args.kmers)
        self.stores = []
        for q in range(n_queries):
            r = O.hash_kmers(self.queries[q], K, H, self.m)
            uniq, inv = np.unique(r.reshape(-1), return_inverse=True)
            store = self.spec.rows(uniq, 0, self.cols)
            self.stores.append((store, np.ascontiguousarray(inv.reshape(r.shape), dtype=np.int64), uniq))

    def step(self, q):
        """hash + per-k-mer AND + per-column count + threshold (graph/index.py:62-80, graph/bigsi.py:35-44,211-242)."""
        O = self.O
        store, slot, uniq = self.stores[q]
        r = O.hash_kmers(self.queries[q], K, H, self.m)  # canonical + murmur3, as the reference does per query
        # row ids -> slots of the in-memory store (the reference's dict lookup by row key)
        slot2 = np.searchsorted(uniq, r.reshape(-1)).reshape(r.shape).astype(np.int64)
        cnt = O.counts_from_rows(store, slot2, self.cols)
        return np.nonzero(cnt >= self.u)[0]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nq = max(1, min(8, args.steps + args.warmup))
    arm = CpuArm(args, nq)
    for i in range(args.warmup):
        arm.step(i % nq)
    t0 = time.perf_counter()
    for i in range(args.steps):
        hits = arm.step((args.warmup + i) % nq)
    dt = time.perf_counter() - t0
    assert len(hits) >= 3
    value = args.kmers * args.steps / dt
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port",
                         "sample": "every step = one full %d-k-mer query on one %d-column shard (rows it touches "
                                   "pre-generated in host RAM, %d distinct queries rotating)" % (args.kmers, args.cols, nq)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, n_gpus):
    return {
        "workload": "BASELINE configs[1]: synthetic m=%d h=%d k=%d, N=%d columns per GPU (x%d GPUs, column-sharded), "
                    "one %d-k-mer exact query (min_kmers=U) per step: canonical+murmur3 hash, fused gather-AND-popcount, "
                    "merge and threshold in ONE kernel launch" % (args.m, H, K, args.cols, n_gpus, args.kmers),
        "m": args.m, "h": H, "k": K, "cols_per_gpu": args.cols, "kmers_per_query": args.kmers,
        "distinct_queries": N_DISTINCT,
        "exchange": ("none (one shard)" if n_gpus == 1 else
                     "in-kernel, pipelined: query pushed to peers and hits all-gathered through NVLink peer memory by the query "
                     "kernel, no collective call; the kernel of query s completes the all-gather of query s-1, the last "
                     "query is drained inside the timed region"
                     if getattr(args, "exchange", "pipelined") == "pipelined" else
                     "in-kernel, lock-step: query pushed to peers and hits all-gathered through NVLink peer memory, no collective call"
                     if getattr(args, "exchange", "pipelined") == "fused" else "NCCL broadcast + all-gather per query"),
        "l2_policy": "inputs larger than L2: each step gathers %.1f MB of distinct rows, %d distinct queries rotate"
                     % (args.kmers * H * math.ceil(args.cols / 8) / 1e6, N_DISTINCT),
        "matrix_density": 0.5,
    }


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import bigsi_b200
    from bigsi_b200.sharded import DeviceShard, ShardedSearcher, unpack_hits

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU oracle")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    U, cols = args.kmers, args.cols
    pc, pt = planted_columns(world, cols)
    index = bigsi_b200.DeviceIndex(args.m, cols, col_offset=rank * cols, device=local_rank)
    t0 = time.perf_counter()
    index.fill_synthetic(0, 1, pc, pt)
    fill_s = time.perf_counter() - t0
    for kv in args.opt:
        key, val = kv.split("=")
        index.set_option(key, int(val))
    shard = DeviceShard(index, K, H, cap=HIT_CAP)
    fused = world > 1 and args.exchange in ("fused", "pipelined")
    piped = world > 1 and args.exchange == "pipelined"
    searcher = ShardedSearcher(shard, dist if world > 1 else None, world, rank, fused_max_kmers=U if fused else 0)

    queries = make_queries(N_DISTINCT, U)
    d_queries = torch.from_numpy(queries).to(dev)  # resident k-mer bytes (value arm)
    h_queries = torch.from_numpy(queries).pin_memory()  # pinned host copy (e2e arm)
    d_qoff = torch.tensor([0, U], dtype=torch.int64, device=dev)
    d_min = torch.tensor([U], dtype=torch.int32, device=dev)
    h_min = np.array([U], dtype=np.uint32)
    h_qoff = np.array([0, U], dtype=np.int64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dev_step(i):
        if fused:  # one kernel per rank and step, no collective: rank 0's k-mers are pushed by its kernel
            return searcher.search_one_fused(d_queries[i % N_DISTINCT] if rank == 0 else None, U, U, pipelined=piped)
        return searcher.search_step(d_queries[i % N_DISTINCT], d_qoff, d_min, 1, U)

    def dev_drain():
        """Pipelined exchange: complete the all-gather of the last query (a no-op otherwise)."""
        return searcher.drain_fused() if piped else None

    # ---- correctness gate on the first query (planted all-ones columns must be the exact hits)
    g = dev_step(0)
    if piped:
        g = dev_drain()
    torch.cuda.synchronize()
    n, hc, hv = unpack_hits(g.cpu().numpy(), 1, HIT_CAP)
    for r in range(world):
        got = sorted(hc[r, 0, : int(n[r, 0])].tolist())
        assert got == [0, 1, cols - 1], "rank %d: unexpected exact hits %r" % (r, got[:10])

    # ---- pre-warm (clocks), then W warm-up steps, then K timed steps: device-resident inputs
    for i in range(args.prewarm):
        dev_step(i)
    dev_drain()
    barrier()
    for i in range(args.warmup):
        dev_step(i)
    dev_drain()
    barrier()
    launches0 = index.info()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        dev_step(args.warmup + i)
    dev_drain()  # every query's hits have arrived on this rank before the clock stops
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = index.info()["kernel_launches"] - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * U * args.steps / (ms_max * 1e-3)

    if args.timeline:
        import ctypes

        from bigsi_b200 import _lib as _L
        index.set_option("debug_flags", 2)
        for i in range(50):
            dev_step(i)
        dev_drain()
        barrier()
        grid = index.info()["last_grid"]
        buf = np.zeros(grid * 16, dtype=np.uint64)
        # the drain kernel does not stamp: the buffer holds the last query kernel's stamps
        _L.check(_L.lib().bigsi_b200_index_debug_read(index.handle, ctypes.c_void_p(buf.ctypes.data), buf.size))
        ts = buf.reshape(grid, 16).astype(np.int64)
        t0 = ts[:, 0].min()
        names = ["entry", "prod_first_issue", "first_slot_landed", "last_slot_consumed", "flushed", "prod_last_issue",
                 "past_grid_barrier", "merge_done", "past_pdl_wait", "hash_done", "merge_loaded", "merge_in_smem",
                 "merge_expanded", "merge_stage_issue"]
        lines = ["rank %d timeline (us since the first CTA entered; min / median / max over CTAs; last CTA separately)" % rank]
        for j, nm in enumerate(names):
            col = ts[:-1, j]
            col = (col[col > 0] - t0) / 1e3
            if col.size:
                lines.append("  %-20s %7.2f %7.2f %7.2f   last CTA %7.2f" % (nm, col.min(), np.median(col), col.max(),
                                                                            (ts[-1, j] - t0) / 1e3 if ts[-1, j] > 0 else -1))
        sys.stderr.write("\n".join(lines) + "\n")
        index.set_option("debug_flags", 0)

    # ---- roofline pass: same K steps with the fused kernel bracketed by CUDA events on its stream
    index.set_option("timing", 1)
    for i in range(args.steps):
        dev_step(args.warmup + i)
    dev_drain()
    barrier()
    fused_ms, merge_ms, n_timed = index.timing_collect()
    index.set_option("timing", 0)
    info = index.info()
    fused_avg_ms = fused_ms / max(n_timed, 1)
    merge_avg_ms = merge_ms / max(n_timed, 1)
    algo_bytes = info["last_algorithmic_bytes"]
    # The timed region is `steps` back-to-back launches of ONE kernel, bracketed by a CUDA-event pair on its stream:
    # its average launch duration there is ms / steps (consecutive launches overlap by the programmatic-dependent-
    # launch prologue, so this is what a launch costs in the stream).  The second pass brackets EVERY launch with
    # its own event pair, which serialises the launches and adds the launch gap: reported as *_isolated.
    kernel_ms = ms_max / args.steps
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    achieved_isolated = algo_bytes / (fused_avg_ms * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "fused_query_dram_bytes.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- e2e arm: the C-ABI host call (N=1) / pinned host -> device -> exchange -> host (N>1)
    def e2e_step(i):
        q = i % N_DISTINCT
        if world == 1:
            # the reference-facing call: BIGSI.search takes a SEQUENCE (graph/bigsi.py:174); its filter stage is one
            # C-ABI call on host buffers (windows -> set of raw k-mers -> threshold -> hash -> gather-AND-count -> hits)
            return index.search_sequence(h_seqs[q], K, H, 1.0, cap=HIT_CAP)
        if piped:
            # pinned host k-mers -> H2D -> kernel of query i; its return value is the COMPLETE result of query i-1,
            # copied to pinned host memory behind the kernel; the host then waits for the copy enqueued one step
            # earlier, so the host->device->host legs of consecutive queries overlap (results arrive two steps late)
            d_k = h_queries[q].to(dev, non_blocking=True) if rank == 0 else None
            prev = searcher.search_one_fused(d_k, U, U, pipelined=True)
            if rank != 0:
                return None
            slot = i % 3
            if prev is not None:
                e2e_host[slot].copy_(prev, non_blocking=True)
            e2e_ev[slot].record()
            e2e_ev[(i - 1) % 3].synchronize()
            return e2e_host[(i - 1) % 3]
        if fused:
            d_k = h_queries[q].to(dev, non_blocking=True) if rank == 0 else None
            g = searcher.search_one_fused(d_k, U, U)
        else:
            d_k = h_queries[q].to(dev, non_blocking=True) if rank == 0 else d_queries[q]
            g = searcher.search_step(d_k, d_qoff, d_min, 1, U)
        return g.cpu() if rank == 0 else None

    if world == 1:
        # 64 distinct random sequences of U + K - 1 bases: U windows, all distinct (checked), so one call = U lookups
        h_seqs = []
        for q in range(N_DISTINCT):
            rng = np.random.default_rng(1000 + q)
            h_seqs.append(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=U + K - 1)].tobytes())
            c, v, nh, uq = index.search_sequence(h_seqs[q], K, H, 1.0, cap=HIT_CAP)
            assert uq == U and sorted(c.tolist()) == [0, 1, cols - 1] and set(v.tolist()) == {U}, (q, uq, c[:8], v[:8])
    if piped:
        e2e_host = [torch.empty((world, 2 + 2 * HIT_CAP), dtype=torch.int32).pin_memory() for _ in range(3)]
        e2e_ev = [torch.cuda.Event() for _ in range(3)]
        for ev in e2e_ev:
            ev.record()
    for i in range(max(args.warmup, 3)):
        e2e_step(i)
    dev_drain()
    barrier()
    e2e_steps = min(args.steps, 500)
    w0 = time.perf_counter()
    for i in range(e2e_steps):
        res = e2e_step(args.warmup + i)
    if piped:  # the last query's hits: drain, then to the host
        last = dev_drain()
        if rank == 0:
            res = last.cpu()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0  # host calls synchronise every step, so wall clock == device time + host overhead
    barrier()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * U * e2e_steps / float(te.item())
    n_hits = 3
    h2d = (U + K - 1) if world == 1 else U * K + 16 + 4
    d2h = 24 + n_hits * 8 if world == 1 else world * (2 + 2 * HIT_CAP) * 4
    e2e_kmers = None
    if world == 1:  # the same query handed over as U unique raw k-mers (31 bytes each) instead of the sequence
        for i in range(3):
            index.search_kmers_hits(h_queries[i].numpy(), K, H, h_min, q_offsets=h_qoff, cap=HIT_CAP)
        w0 = time.perf_counter()
        for i in range(e2e_steps):
            index.search_kmers_hits(h_queries[(args.warmup + i) % N_DISTINCT].numpy(), K, H, h_min, q_offsets=h_qoff, cap=HIT_CAP)
        dt = time.perf_counter() - w0
        e2e_kmers = {"value": U * e2e_steps / dt, "ms_per_step": 1e3 * dt / e2e_steps, "h2d_bytes_per_step": U * K + 16 + 4,
                     "path": "bigsi_b200_search_kmers_hits (C ABI, pinned host k-mers read zero-copy by the kernel)"}

    # ---- AND-mode (exact_filter) kernel, for context
    d_and = torch.empty((1, (cols + 7) // 8 + 16), dtype=torch.uint8, device=dev)
    index.set_option("timing", 1)
    for i in range(20):
        r = shard.hash(d_queries[i % N_DISTINCT])
        index.query_dev(1, r.data_ptr(), d_qoff.data_ptr(), 1, U, H, d_and.data_ptr(), d_and.shape[1],
                        torch.cuda.current_stream().cuda_stream, U)
    torch.cuda.synchronize()
    and_ms, and_merge_ms, and_n = index.timing_collect()
    index.set_option("timing", 0)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline(args)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitwise (LOP3) / u32 counts", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * float(te.item()) / e2e_steps,
                    "path": "bigsi_b200_search_sequence (C ABI, host sequence in, hit list out: front-end kernel + ONE query kernel, "
                            "no host round trip in between)" if world == 1 else
                            ("pinned host k-mers -> H2D on rank 0 -> ONE kernel per rank (k-mers pushed to the peers over NVLink in the prologue, "
                             "hash, gather-AND-count, merge, threshold, hits published to every rank's result blocks) -> D2H"
                             + ("; pipelined: query i's kernel completes the all-gather of query i-1, whose D2H the host awaits one step later" if piped else "")
                             if fused else
                             "pinned host k-mers -> H2D -> hash -> NCCL broadcast -> fused query -> threshold -> NCCL all-gather -> D2H")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "fused_query<COUNTS,h=3> (in-kernel hash prologue + gather-AND-popcount + grid barrier + merge/threshold phase)" if info["last_fused"] == 3 else "fused_query<COUNTS,h=3>", "kernel_ms": kernel_ms,
                         "timing": "CUDA events around the timed region of back-to-back launches, / launches",
                         "kernel_ms_isolated": fused_avg_ms, "achieved_isolated": achieved_isolated,
                         "frac_isolated": achieved_isolated / peak,
                         "timing_isolated": "one CUDA-event pair per launch (serialises the launches, includes the launch gap)",
                         "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                         "frac_of_8TBps_nominal": achieved / 8000.0, "merge_kernel_ms": merge_avg_ms,
                         "launches_timed": int(n_timed)},
            "and_mode": {"kernel_ms": and_ms / max(and_n, 1), "achieved_GBps": algo_bytes / (and_ms / max(and_n, 1) * 1e-3) / 1e9,
                         "merge_kernel_ms": and_merge_ms / max(and_n, 1)},
            "launch_geometry": {kk: info[kk] for kk in ("last_grid", "last_block", "last_smem_bytes", "last_tile_bytes",
                                                        "last_n_tiles", "last_kmers_per_stage", "last_n_stages",
                                                        "last_n_slices")},
            "index": {"matrix_bytes": info["matrix_bytes"], "row_pitch_bytes": info["row_pitch_bytes"], "fill_seconds": fill_s},
        }
        if e2e_kmers is not None:
            line["e2e_kmers_path"] = e2e_kmers
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    index.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_cpu_baseline(args):
    """Bounded sample of the same workload on the host cores (oracle port, all threads)."""
    nq = 4
    arm = CpuArm(args, nq)
    arm.step(0)
    t0 = time.perf_counter()
    done = 0
    while True:
        arm.step(done % nq)
        done += 1
        dt = time.perf_counter() - t0
        if dt >= args.cpu_seconds or done >= 2000:
            break
    return {"value": args.kmers * done / dt, "unit": UNIT, "cores": arm.cores, "kind": "port",
            "sample": "%d full %d-k-mer queries (%d distinct, rotating) on one %d-column shard in %.1f s; rows "
                      "pre-generated in host RAM; hash + AND + per-column count + threshold, OpenMP"
                      % (done, args.kmers, nq, args.cols, dt)}


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
