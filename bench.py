#!/usr/bin/env python
"""bench.py -- BIGSI search hot path on B200 (BASELINE.json configs[1], weak-scaled by column shard).

Workload (config.workload): synthetic m=25 000 000, h=3, k=31 index with 50 000 sample columns
PER GPU resident in HBM (156.8 GB/GPU); one STEP = a block of 64 distinct 10 000-k-mer queries, each
run through the hot path on its own: canonical+murmur3 hashing, gather-AND-popcount (the gather warps of
its kernel), merge + threshold at min_kmers = U (an exact query through the count path; done by the merge
team of the NEXT query's kernel, by a flush kernel behind the 64th), i.e. one kernel launch per query,
no batching across queries.  The 64 queries of a step are all different (187.5 MB of rows per query >
126 MB L2), every query's hit list is produced separately and is complete when the step ends.

Metric: k-mer row-AND lookups/s, one lookup = gather h rows of one 50 000-column shard and AND
them (18 750 algorithmic bytes).  At N GPUs every rank looks the same k-mers up in its own
column shard (rank 0's gather kernel pushes the k-mers to the peers, the merge teams all-gather
the hits), so the whole-job value is N * U * 64 * steps / time.

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
QUERIES_PER_STEP = 64  # one step = every distinct query once
HIT_CAP = 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40, help="timed steps; one step = 64 queries")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=25_000_000, help="Bloom filter size (rows); default = BASELINE")
    ap.add_argument("--cols", type=int, default=50_000, help="sample columns per GPU")
    ap.add_argument("--kmers", type=int, default=10_000, help="unique k-mers per query")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prewarm", type=int, default=5, help="untimed steps before the warm-up (clock ramp)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N>1: query broadcast + hit all-gather inside the query kernels over NVLink peer memory (default), "
                         "or as two NCCL collectives per query (nccl)")
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
    """The oracle port of the reference's algorithm on the same synthetic workload: `shards` column shards of
    `cols` columns (the N-GPU job's whole index).  Rows the queries touch are regenerated from the synthetic
    index's pure function BEFORE timing (the reference's in-memory store would hold them already)."""

    def __init__(self, args, n_queries, shards=1):
        from oracle import oracle as O

        self.O = O
        O.lib().oracle_set_num_threads(os.cpu_count() or 1)  # torchrun pins OMP_NUM_THREADS=1: use every host core
        self.cores = O.lib().oracle_num_threads()
        self.shards = shards
        pc, pt = planted_columns(shards, args.cols)
        self.spec = O.SynthSpec(0, 1, pc, pt)
        self.m, self.cols, self.u = args.m, args.cols, args.kmers
        self.queries = make_queries(n_queries, args.kmers)
        self.stores = []
        for q in range(n_queries):
            r = O.hash_kmers(self.queries[q], K, H, self.m)
            uniq, inv = np.unique(r.reshape(-1), return_inverse=True)
            per_shard = [self.spec.rows(uniq, g * self.cols, self.cols) for g in range(shards)]
            self.stores.append((per_shard, uniq))

    def step(self, q):
        """hash + per-k-mer AND + per-column count + threshold (graph/index.py:62-80, graph/bigsi.py:35-44,211-242),
        over every shard of the index."""
        O = self.O
        per_shard, uniq = self.stores[q]
        r = O.hash_kmers(self.queries[q], K, H, self.m)  # canonical + murmur3, as the reference does per query
        # row ids -> slots of the in-memory store (the reference's dict lookup by row key)
        slot = np.searchsorted(uniq, r.reshape(-1)).reshape(r.shape).astype(np.int64)
        hits = []
        for g, store in enumerate(per_shard):
            cnt = O.counts_from_rows(store, slot, self.cols)
            hits.append(np.nonzero(cnt >= self.u)[0] + g * self.cols)
        return np.concatenate(hits)


def run_reference(args):
    """--impl reference: the reference's algorithm (oracle port, OpenMP on every host core) on the GPU arm's
    config: the whole N-shard index, one full query per step.  Under torchrun rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    shards = max(1, args.gpus)
    nq = max(1, min(4, args.steps + args.warmup))
    arm = CpuArm(args, nq, shards)
    for i in range(args.warmup):
        arm.step(i % nq)
    t0 = time.perf_counter()
    for i in range(args.steps):
        hits = arm.step((args.warmup + i) % nq)
    dt = time.perf_counter() - t0
    assert len(hits) >= 3 * shards
    value = shards * args.kmers * args.steps / dt
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": workload_config(args, shards),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port",
                         "sample": "every step = ONE full %d-k-mer query over all %d column shards of %d columns (a bounded "
                                   "sample of the GPU arm's step of %d such queries; rows it touches pre-generated in host RAM, "
                                   "%d distinct queries rotating)" % (args.kmers, shards, args.cols, QUERIES_PER_STEP, nq)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, n_gpus):
    return {
        "workload": "BASELINE configs[1]: synthetic m=%d h=%d k=%d, N=%d columns per GPU (x%d GPUs, column-sharded), "
                    "%d-k-mer exact queries (min_kmers=U), each on its own: canonical+murmur3 hash + gather-AND-popcount "
                    "(one kernel per query), merge + threshold by the merge team of the next query's kernel (flush kernel behind "
                    "the last); one step = %d distinct queries back to back, all hit lists complete at its end"
                    % (args.m, H, K, args.cols, n_gpus, args.kmers, QUERIES_PER_STEP),
        "m": args.m, "h": H, "k": K, "cols_per_gpu": args.cols, "kmers_per_query": args.kmers,
        "distinct_queries": N_DISTINCT, "queries_per_step": QUERIES_PER_STEP,
        "exchange": ("none (one shard)" if n_gpus == 1 else
                     "in-kernel: rank 0's gather kernel pushes the query to the peers, every rank's merge team publishes its hits "
                     "to every rank and waits for the others' (while the next query's rows stream) -- NVLink peer memory, "
                     "no collective call"
                     if getattr(args, "exchange", "fused") == "fused" else "NCCL broadcast + all-gather per query"),
        "l2_policy": "inputs larger than L2: each query gathers %.1f MB of distinct rows, %d distinct queries rotate"
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

    U, cols, QPS = args.kmers, args.cols, QUERIES_PER_STEP
    pc, pt = planted_columns(world, cols)
    index = bigsi_b200.DeviceIndex(args.m, cols, col_offset=rank * cols, device=local_rank)
    t0 = time.perf_counter()
    index.fill_synthetic(0, 1, pc, pt)
    fill_s = time.perf_counter() - t0
    index.set_option("inputs_ready", 1)  # the k-mer buffers below are resident / host-written before every call
    for kv in args.opt:
        key, val = kv.split("=")
        index.set_option(key, int(val))
    shard = DeviceShard(index, K, H, cap=HIT_CAP)
    fused = world > 1 and args.exchange == "fused"
    searcher = ShardedSearcher(shard, dist if world > 1 else None, world, rank, fused_max_kmers=U if fused else 0)

    queries = make_queries(N_DISTINCT, U)
    d_queries = torch.from_numpy(queries).to(dev)  # resident k-mer bytes (value arm)
    h_queries = torch.from_numpy(queries).pin_memory()  # pinned host copy (e2e arm)
    d_qoff = torch.tensor([0, U], dtype=torch.int64, device=dev)
    d_min = torch.tensor([U], dtype=torch.int32, device=dev)
    h_min = np.array([U], dtype=np.uint32)
    h_qoff = np.array([0, U], dtype=np.int64)
    # the ClockSampler (NVML init, thread) exists BEFORE any barrier that precedes a timed region
    sampler = ClockSampler(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def dev_query(q, min_kmers=U):
        """One query, DEFERRED launch: ONE kernel per rank (gather of this query + merge/threshold/publication of the
        previous one by its merge team); the returned buffer is complete after the next dev_query or dev_flush."""
        if fused:  # no collective: rank 0's k-mers are pushed by its gather kernel, the merge teams all-gather the hits
            return searcher.search_one_fused(d_queries[q] if rank == 0 else None, U, min_kmers)
        if world == 1:
            return shard.search_kmers_hits_stream(d_queries[q], min_kmers)[None]
        if min_kmers != U:
            return searcher.search_step(d_queries[q], d_qoff, torch.tensor([min_kmers], dtype=torch.int32, device=dev), 1, U)
        return searcher.search_step(d_queries[q], d_qoff, d_min, 1, U)

    def dev_flush():
        if fused or world == 1:
            index.flush()  # stage 2 of the last query as a kernel of its own

    def dev_step(i):
        for j in range(QPS):
            g = dev_query((i * QPS + j) % N_DISTINCT)
        dev_flush()  # every hit list of the step is complete in stream order when the step ends
        return g

    # ---- correctness gate (the parity tests proper are tests/, this guards the bench's own wiring): query 0
    # exact -> exactly the planted all-ones columns on every shard; query 1 at score >= 0.4 -> those plus the graded
    # column of density 0.95, whose count must be the same through the batch path (generic kernel) of the same shard
    g = dev_query(0)
    dev_flush()
    torch.cuda.synchronize()
    n, hc, hv = unpack_hits(g.cpu().numpy(), 1, HIT_CAP)
    for r in range(world):
        got = sorted(hc[r, 0, : int(n[r, 0])].tolist())
        assert got == [0, 1, cols - 1], "rank %d: unexpected exact hits %r" % (r, got[:10])
    thr04 = int(math.ceil(U * 0.4))
    g = dev_query(1, thr04)
    dev_flush()
    torch.cuda.synchronize()
    n, hc, hv = unpack_hits(g.cpu().numpy(), 1, HIT_CAP)
    d_cnt = torch.zeros((1, cols + 8), dtype=torch.int32, device=dev)
    rows1 = shard.hash(d_queries[1])
    index.query_dev(0, rows1.data_ptr(), d_qoff.data_ptr(), 1, U, H, d_cnt.data_ptr(), d_cnt.shape[1],
                    torch.cuda.current_stream().cuda_stream, U)
    cnt_local = d_cnt[0, :cols].cpu().numpy().astype(np.int64)
    exp_local = np.nonzero(cnt_local >= thr04)[0]
    assert set(exp_local.tolist()) == {0, 1, cols - 1, cols // 2}, exp_local[:10]
    mine = 0 if world == 1 else rank
    order = np.argsort(hc[mine, 0, : int(n[mine, 0])])
    assert np.array_equal(hc[mine, 0, : int(n[mine, 0])][order], exp_local), "graded hits differ between the two paths"
    assert np.array_equal(hv[mine, 0, : int(n[mine, 0])][order].astype(np.int64), cnt_local[exp_local])
    for r in range(world):  # every shard reports its own graded column with a count near 0.95^3 * U
        got = dict(zip(hc[r, 0, : int(n[r, 0])].tolist(), hv[r, 0, : int(n[r, 0])].tolist()))
        assert sorted(got) == [0, 1, cols // 2, cols - 1], "rank %d: graded hits %r" % (r, sorted(got)[:10])
        assert abs(got[cols // 2] - 0.95 ** 3 * U) < 0.03 * U and got[0] == got[1] == got[cols - 1] == U

    # ---- pre-warm (clocks), W warm-up steps, aligned start, K timed steps: device-resident inputs
    for i in range(args.prewarm + args.warmup):
        dev_step(i)
    launches0 = index.info()["kernel_launches"]
    sampler.start()
    barrier()
    dev_query(0)  # device-side rendezvous: one untimed query through the exchange right before the clock starts
    dev_flush()
    barrier()
    if fused:
        searcher.fused.wait_ns()  # reset the wait counter
        barrier()
    launches0 = index.info()["kernel_launches"]  # (after the rendezvous query: only the timed region's launches count)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        dev_step(args.warmup + i)
    ev1.record()  # every query's hits have arrived on this rank when this event completes (stream order)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = index.info()["kernel_launches"] - launches0
    exchange_wait_us = None
    if fused:
        w_ns, _ = searcher.fused.wait_ns()
        exchange_wait_us = w_ns / 1e3 / (args.steps * QPS)
    t = torch.tensor([ms, exchange_wait_us or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        per_rank = torch.stack(gathered).cpu().numpy()
    else:
        per_rank = t.cpu().numpy()[None]
    ms_max = float(per_rank[:, 0].max())
    n_queries = args.steps * QPS
    value = world * U * n_queries / (ms_max * 1e-3)

    if args.timeline:
        dump_timeline(index, dev_query, barrier, rank)

    # ---- isolated pass: one CUDA-event pair per launch (serialises the launches, no overlap between queries)
    index.set_option("timing", 1)  # (no deferral under timing: gather kernel, then its flush kernel)
    for j in range(QPS):
        dev_query(j)
    barrier()
    fused_ms, merge_ms, n_timed = index.timing_collect()
    index.set_option("timing", 0)
    info = index.info()
    fused_avg_ms = fused_ms / max(n_timed, 1)
    merge_avg_ms = merge_ms / max(n_timed, 1)
    algo_bytes = info["last_algorithmic_bytes"]
    # The timed region is steps x 64 back-to-back queries, bracketed by ONE CUDA-event pair on their stream.  A query =
    # one gather_solo launch (all the row traffic; its merge team finishes the previous query in the shadow of the row
    # stream) and consecutive launches overlap CTA by CTA, so the kernel's average launch duration in the stream is
    # region / queries (the one flush kernel per step is inside the region too).
    kernel_ms = ms_max / n_queries
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    achieved_isolated = algo_bytes / (fused_avg_ms * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_gather_dram_bytes.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            traffic = None

    # ---- e2e arm: host buffers in, hit lists out, through the C ABI
    e2e = run_e2e(args, index, searcher, shard, queries, h_queries, barrier, rank, world, dev, fused)

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

    # ---- the same 64 queries as ONE batched launch (bulk_search, bigsi/__main__.py:261-314), for context: hash kernel +
    # one generic gather kernel over the whole batch with the merge behind its grid barrier; all hit lists at the end
    d_all = d_queries.reshape(N_DISTINCT * U, K)
    d_qoff_b = torch.arange(0, (N_DISTINCT + 1) * U, U, dtype=torch.int64, device=dev)
    d_min_b = torch.full((N_DISTINCT,), U, dtype=torch.int32, device=dev)
    for i in range(2):
        gb = shard.search_kmers_hits(d_all, d_qoff_b, N_DISTINCT, d_min_b, U)
    torch.cuda.synchronize()
    nb_, cb_, vb_ = unpack_hits(gb.cpu().numpy(), N_DISTINCT, HIT_CAP)
    assert all(sorted(cb_[0, q, : int(nb_[0, q])].tolist()) == [0, 1, cols - 1] for q in range(N_DISTINCT))
    eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eb0.record()
    for i in range(5):
        shard.search_kmers_hits(d_all, d_qoff_b, N_DISTINCT, d_min_b, U)
    eb1.record()
    torch.cuda.synchronize()
    batch_ms = eb0.elapsed_time(eb1) / 5

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline(args)

    if rank == 0:
        streamed = bool(info["last_fused"] & 8)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitwise (LOP3) / u32 counts", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks,
            "us_per_query": 1e3 * ms_max / n_queries,
            "per_rank_ms_per_step": [float(x) / args.steps for x in per_rank[:, 0]],
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "gather_solo<COUNTS,h=3,team> (in-kernel hash prologue + TMA row gather + AND + bit-sliced count; merge + "
                                   "threshold of the previous query by its merge team)" if streamed else "fused_query<COUNTS,h=3>",
                         "kernel_ms": kernel_ms,
                         "timing": "one CUDA-event pair around the timed region of back-to-back queries, / queries",
                         "kernel_ms_isolated": fused_avg_ms, "achieved_isolated": achieved_isolated,
                         "frac_isolated": achieved_isolated / peak,
                         "timing_isolated": "one CUDA-event pair per launch (serialises the launches: no overlap between queries, includes the "
                                            "launch gap; the query's flush kernel is timed separately as reduce_kernel_ms_isolated)",
                         "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                         "frac_of_8TBps_nominal": achieved / 8000.0, "reduce_kernel_ms_isolated": merge_avg_ms,
                         "launches_timed": int(n_timed)},
            "and_mode": {"kernel_ms": and_ms / max(and_n, 1), "achieved_GBps": algo_bytes / (and_ms / max(and_n, 1) * 1e-3) / 1e9,
                         "merge_kernel_ms": and_merge_ms / max(and_n, 1)},
            "batched_64_queries_one_launch": {"ms": batch_ms, "us_per_query": 1e3 * batch_ms / N_DISTINCT,
                                              "value": U * N_DISTINCT / (batch_ms * 1e-3),
                                              "achieved_GBps": algo_bytes * N_DISTINCT / (batch_ms * 1e-3) / 1e9,
                                              "note": "context, not the headline: the 64 queries of a step handed over as ONE batch "
                                                      "(hash kernel + one gather kernel + in-kernel merge), hit lists available at the end"},
            "launch_geometry": {kk: info[kk] for kk in ("last_grid", "last_block", "last_smem_bytes", "last_tile_bytes",
                                                        "last_n_tiles", "last_kmers_per_stage", "last_n_stages",
                                                        "last_n_slices", "last_reduce_grid")},
            "index": {"matrix_bytes": info["matrix_bytes"], "row_pitch_bytes": info["row_pitch_bytes"], "fill_seconds": fill_s},
        }
        if exchange_wait_us is not None:
            line["exchange_wait_us_per_query"] = [float(x) for x in per_rank[:, 1]]
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    index.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(args, index, searcher, shard, queries, h_queries, barrier, rank, world, dev, fused):
    """The same metric end to end: HOST buffers in, hit lists in HOST memory out, every step.  N = 1: the
    reference-facing C-ABI call bigsi_b200_search_sequence (BIGSI.search takes a SEQUENCE, graph/bigsi.py:174), 64
    calls per step.  N > 1: rank 0's pinned host k-mers are read by its gather kernel (zero copy over PCIe) and
    pushed to the peers; the all-gathered hits are copied to pinned host memory behind every query."""
    import torch

    U, cols, QPS = args.kmers, args.cols, QUERIES_PER_STEP
    e2e_steps = max(1, min(args.steps, 10))
    extra = {}
    if world == 1:
        # 64 distinct random sequences of U + K - 1 bases: U windows, all distinct (checked), so one call = U lookups
        h_seqs = []
        for q in range(N_DISTINCT):
            rng = np.random.default_rng(1000 + q)
            h_seqs.append(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=U + K - 1)].tobytes())
            c, v, nh, uq = index.search_sequence(h_seqs[q], K, H, 1.0, cap=HIT_CAP)
            assert uq == U and sorted(c.tolist()) == [0, 1, cols - 1] and set(v.tolist()) == {U}, (q, uq, c[:8], v[:8])

        def e2e_step(i):
            # one step = ONE bulk call with the step's 64 host sequences (bulk_search, bigsi/__main__.py:261-314): 64 hit
            # lists + 64 k-mer counts back in host memory
            res = index.search_sequences(h_seqs, K, H, 1.0, cap=HIT_CAP)
            assert len(res) == QPS and res[-1][3] == U and res[-1][2] == 3
            return res

        h2d, d2h = QPS * (U + K - 1), QPS * (24 + 3 * 8)
        path = ("bigsi_b200_search_sequences (C ABI bulk call: %d host sequences in, %d hit lists out; per sequence the unique "
                "windows are found inside the gather kernel, which reads the sequence out of pinned host memory; the merge team "
                "of the next sequence's kernel writes the hits into mapped host memory; up to 7 searches in flight)" % (QPS, QPS))
        # the same 64 sequences one synchronous call at a time (BIGSI.search's own call pattern)
        for q in range(3):
            index.search_sequence(h_seqs[q], K, H, 1.0, cap=HIT_CAP)
        w0 = time.perf_counter()
        for q in range(QPS):
            index.search_sequence(h_seqs[q], K, H, 1.0, cap=HIT_CAP)
        dt = time.perf_counter() - w0
        extra["single_call"] = {"value": U * QPS / dt, "us_per_query": 1e6 * dt / QPS,
                                "path": "bigsi_b200_search_sequence, one synchronous call per sequence (nothing in flight around it)"}
        # ... and through the Python drop-in class: BIGSI(config).search(seq, 1.0) -> list of result dicts
        from bigsi_b200 import bigsi as _bg

        store = _bg._Store(index, args.m, H, K)
        _bg.SampleMetadata(store.meta).add_samples(["s%d" % c for c in range(cols)])
        api_cfg = {"k": K, "m": args.m, "h": H, "storage-engine": "b200", "storage-config": {"filename": "bench-e2e", "device": index.info()["device"]}}
        _bg._STORES[_bg._store_key(api_cfg)] = store
        api = _bg.BIGSI(api_cfg)
        s_strs = [s.decode("ascii") for s in h_seqs]
        for q in range(3):
            r = api.search(s_strs[q], 1.0)
        assert [d["sample_name"] for d in r] == ["s0", "s1", "s%d" % (cols - 1)] and r[0]["num_kmers"] == U
        w0 = time.perf_counter()
        for q in range(QPS):
            api.search(s_strs[q], 1.0)
        dt = time.perf_counter() - w0
        del _bg._STORES[_bg._store_key(api_cfg)]
        extra["python_search"] = {"value": U * QPS / dt, "us_per_query": 1e6 * dt / QPS,
                                  "path": "bigsi_b200.BIGSI(config).search(seq, threshold=1.0): str in, list of result dicts out"}
    else:
        h_out = [torch.empty((world, 2 + 2 * HIT_CAP), dtype=torch.int32).pin_memory() for _ in range(4)]
        d_qoff = torch.tensor([0, U], dtype=torch.int64, device=dev)
        d_min = torch.tensor([U], dtype=torch.int32, device=dev)

        def e2e_step(i):
            if fused:
                # rank 0: the gather kernel reads the pinned host k-mers itself (zero copy) and pushes them to the peers; rank
                # 0's stage-2 code writes the all-gathered hit lists into mapped host memory, where rank 0 consumes them with
                # up to 6 searches in flight -- no copy operation in the stream, consecutive queries keep overlapping
                ex = searcher.fused
                ex.enable_host_results(rank == 0)
                seqs = []
                last = None
                for j in range(QPS):
                    q = (i * QPS + j) % N_DISTINCT
                    ex.search(h_queries[q].data_ptr() if rank == 0 else None, U, U)
                    seqs.append(ex.last_seq())
                    if rank == 0 and j >= 6:
                        last = ex.wait_host(seqs[j - 6])
                ex.flush()
                if rank == 0:
                    for s_ in seqs[max(0, QPS - 6):]:
                        last = ex.wait_host(s_)
                    assert int(last[0, 0]) == 3 and int(last[world - 1, 0]) == 3  # 3 exact hits per shard, low word of n_hits
                torch.cuda.synchronize()
                ex.enable_host_results(False)
                return last
            for j in range(QPS):
                q = (i * QPS + j) % N_DISTINCT
                d_k = h_queries[q].to(dev, non_blocking=True)
                g = searcher.search_step(d_k, d_qoff, d_min, 1, U)
                if rank == 0:
                    h_out[j % 4].copy_(g, non_blocking=True)
            torch.cuda.synchronize()
            return h_out[(QPS - 1) % 4]

        # fused: per query and shard the block header (16 B) + 3 hits x 8 B land in host memory, + the 8-byte completion word
        h2d, d2h = QPS * U * K, (QPS * (world * (16 + 3 * 8) + 8) if fused else QPS * world * (2 + 2 * HIT_CAP) * 4)
        path = ("rank 0: pinned host k-mers read zero-copy by the gather kernel and pushed to the peers over NVLink; 1 kernel "
                "per rank and query; the all-gathered hit lists are written into mapped host memory by the stage-2 code and read "
                "there by rank 0 with up to 6 searches in flight (FusedExchange.wait_host)"
                if fused else "pinned host k-mers -> H2D -> NCCL broadcast -> query -> NCCL all-gather -> D2H")
    for i in range(2):
        e2e_step(i)
    barrier()
    w0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(2 + i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0  # the host calls synchronise every step, so wall clock == device time + host overhead
    barrier()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    if world == 1:  # the same query handed over as U unique raw k-mers (31 bytes each) instead of the sequence
        h_min = np.array([U], dtype=np.uint32)
        h_qoff = np.array([0, U], dtype=np.int64)
        for i in range(3):
            index.search_kmers_hits(h_queries[i].numpy(), K, H, h_min, q_offsets=h_qoff, cap=HIT_CAP)
        w0 = time.perf_counter()
        nk = QPS * min(e2e_steps, 3)
        for i in range(nk):
            index.search_kmers_hits(h_queries[i % N_DISTINCT].numpy(), K, H, h_min, q_offsets=h_qoff, cap=HIT_CAP)
        dt = time.perf_counter() - w0
        extra["kmers_path"] = {"value": U * nk / dt, "us_per_query": 1e6 * dt / nk, "h2d_bytes_per_query": U * K + 16 + 4,
                               "path": "bigsi_b200_search_kmers_hits (C ABI, pinned host k-mers read zero-copy by the kernel)"}
    out = {"value": world * U * QPS * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "us_per_query": 1e6 * e2e_s / (e2e_steps * QPS), "path": path}
    out.update(extra)
    return out


def dump_timeline(index, dev_query, barrier, rank):
    """Diagnostics: per-CTA globaltimer stamps of the gather / reduce kernels of the last 8 queries of a short burst:
    a summary of the last query on stderr, the raw stamps in gpurun_out/timeline_rank<r>.npy."""
    import ctypes

    from bigsi_b200 import _lib as _L

    index.set_option("debug_flags", 2)
    for i in range(32):
        dev_query(i)
    index.flush()
    barrier()
    info = index.info()
    grid, rgrid = info["last_grid"], max(info["last_reduce_grid"], info["sm_count"])
    per = grid + rgrid
    buf = np.zeros(8 * per * 16, dtype=np.uint64)
    _L.check(_L.lib().bigsi_b200_index_debug_read(index.handle, ctypes.c_void_p(buf.ctypes.data), buf.size))
    allts = buf.reshape(8, per, 16).astype(np.int64)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.save(os.path.join(ROOT, "gpurun_out", "timeline_rank%d.npy" % rank), allts)
    except OSError:
        pass
    last = int(np.argmax(allts[:, :grid, 0].min(axis=1)))  # the region with the latest gather entry = the last query
    ts = allts[last]
    t0 = ts[:grid, 0].min()
    lines = ["rank %d timeline of the last query (us since its first gather CTA entered; min / median / max over CTAs)" % rank]
    for title, block, names in (("gather kernel", ts[:grid], {0: "entry", 8: "past_gate", 10: "front_end_done", 1: "prod_first_issue",
                                                             9: "hash_done", 2: "first_slot_landed", 5: "prod_last_issue",
                                                             3: "last_slot_consumed", 4: "flushed", 6: "team_push_done",
                                                             7: "team_done (merge of the previous query)"}),
                                ("stage 2 of this query (merge team of the next kernel / flush kernel)", ts[grid:],
                                 {0: "entry (flush kernel only)", 8: "past_dependency_wait", 13: "stage_issue",
                                                             10: "planes_loaded", 11: "counters_in_smem", 12: "expanded",
                                                             7: "items_done"})):
        lines.append(" " + title)
        for j, nm in names.items():
            col = block[:, j]
            col = (col[col > 0] - t0) / 1e3
            if col.size:
                lines.append("  %-22s %8.2f %8.2f %8.2f   (n=%d)" % (nm, col.min(), np.median(col), col.max(), col.size))
    firsts = np.sort(allts[:, :grid, 0].min(axis=1))
    lines.append(" first gather-CTA entry of the last 8 queries, differences (us): %s"
                 % np.round(np.diff(firsts) / 1e3, 2).tolist())
    sys.stderr.write("\n".join(lines) + "\n")
    index.set_option("debug_flags", 0)


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
    out = {"value": args.kmers * done / dt, "unit": UNIT, "cores": arm.cores, "kind": "port",
           "sample": "%d full %d-k-mer queries (%d distinct, rotating) on one %d-column shard in %.1f s; rows "
                     "pre-generated in host RAM; hash + AND + per-column count + threshold, OpenMP"
                     % (done, args.kmers, nq, args.cols, dt)}
    # the UNMODIFIED reference package's own BIGSI.search, when it is installed (baseline/_ref/): single thread (its search
    # is single-threaded, graph/bigsi.py:64-85), a reduced query (BASELINE.md section 3 line 1)
    try:
        from oracle import ref_harness, ref_timing

        if ref_harness.reference_available():
            u_ref = min(args.kmers, 2000)
            pc, pt = planted_columns(1, args.cols)
            r = ref_timing.time_reference_search(args.m, args.cols, K, H, u_ref, pc, pt)
            assert r[1.0][2] == ["s0", "s1", "s%d" % (args.cols - 1)], r[1.0][2]
            out["reference_python"] = {
                "value": r[1.0][1], "unit": UNIT, "cores": 1, "kind": "reference",
                "value_threshold_0.4": r[0.4][1],
                "sample": "the unmodified bigsi 0.3.8 package (baseline/_ref, imported through the mmh3 / bitarray stand-ins of "
                          "oracle/ref_shims; dict-backed BaseStorage holding the rows the query touches): BIGSI.search of ONE "
                          "%d-k-mer sequence on one %d-column shard, exact (%.3f s) and at threshold 0.4 (%.3f s); not "
                          "extrapolated: per-k-mer cost is flat in the query length" % (u_ref, args.cols, r[1.0][0], r[0.4][0])}
    except Exception as e:  # noqa: BLE001 -- a reported extra, never a reason to lose the bench line
        out["reference_python"] = {"unavailable": repr(e)[:200]}
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
