"""One rank of tests/test_gpu_parity.py::test_exchange_across_processes_ipc (run as a script, one process per
shard): builds its column shard of a small random index, wires the exchange through CUDA IPC (gloo carries the
64-byte handles, so that several ranks may share one GPU), runs a burst of queries back to back and checks its own
copy of every query's all-gathered hits against the CPU oracle.  Writes <outdir>/rank<r>.json."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    import bigsi_b200 as B
    from bigsi_b200.sharded import DeviceShard, FusedExchange, merge_shard_hits, unpack_hits
    from oracle import oracle as O

    outdir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(61)  # the same matrix and queries on every rank
    part = 1600
    m, N, k, h, cap = 15_013, part * world, 31, 3, 2048
    rows = rng.random((m, N)) < 0.9
    packed = np.packbits(rows, axis=1)
    oix = O.OracleIndex(k, m, h, N, rows=packed)
    ix = B.DeviceIndex(m, part, col_offset=rank * part, device=local)
    ix.upload_rows(0, packed, src_byte_offset=rank * part // 8)
    ix.set_option("inputs_ready", 1)
    shard = DeviceShard(ix, k, h, cap=cap)
    ex = FusedExchange(shard, world, rank, 8000, dist=dist)
    ex.enable_host_results()  # the kernels also write every query's gathered hits into mapped host memory
    sizes = [25, 3000, 1, 7000, 640, 2500, 6000, 77, 4000, 4000, 333, 8000, 5, 1234]
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    queries = [acgt[rng.integers(0, 4, size=(n, k))] for n in sizes]
    d_queries = [torch.from_numpy(q).to(shard.device) for q in queries] if rank == 0 else [None] * len(sizes)
    torch.cuda.synchronize()
    dist.barrier()
    ok, detail = True, ""
    try:
        copies, prev, seqs, host = [], None, [], {}
        for j, n in enumerate(sizes):  # back to back: no host synchronisation between the queries
            thr = int(math.ceil(n * (0.85 if j % 3 else 0.4)))
            view = ex.search(d_queries[j], n, thr)
            seqs.append(ex.last_seq())
            if prev is not None:  # deferred: the result of query j-1 is complete behind the launch of query j
                copies.append(prev.clone())
            prev = view
            if j >= 3 and rank % 2 == 0:  # host results consumed while later searches are in flight (some ranks only)
                host[j - 3] = np.array(ex.wait_host(seqs[j - 3]))
        ex.flush()  # the last one: stage 2 as a kernel of its own (every rank)
        copies.append(prev.clone())
        for j in range(max(0, len(sizes) - 6), len(sizes)):
            host[j] = np.array(ex.wait_host(seqs[j]))
        torch.cuda.synchronize()
        _lib = B._lib
        _lib.check(_lib.lib().bigsi_b200_index_status(ix.handle))
        offs = [g * part for g in range(world)]
        for j, n in enumerate(sizes):
            thr = int(math.ceil(n * (0.85 if j % 3 else 0.4)))
            cnt = oix.counts([bytes(r).decode() for r in queries[j]])
            exp = np.nonzero(cnt >= thr)[0]
            nh, cols, vals = unpack_hits(copies[j].cpu().numpy(), 1, cap)
            if j in host:  # the host copy carries the same hit counts and the same hits (up to what a block holds)
                hn, hc_, hv_ = unpack_hits(host[j], 1, cap)
                same = np.array_equal(hn, nh) and all(
                    np.array_equal(hc_[g, 0, : min(int(nh[g, 0]), cap)], cols[g, 0, : min(int(nh[g, 0]), cap)]) and
                    np.array_equal(hv_[g, 0, : min(int(nh[g, 0]), cap)], vals[g, 0, : min(int(nh[g, 0]), cap)]) for g in range(world))
                if not same:
                    ok, detail = False, "query %d: host result block differs from the device blocks on rank %d" % (j, rank)
                    break
            if (nh[:, 0] > cap).any():
                if int(nh[:, 0].sum()) != len(exp):
                    ok, detail = False, "query %d: %d hits, expected %d" % (j, int(nh[:, 0].sum()), len(exp))
                continue
            gc, gv = merge_shard_hits(nh[:, 0], cols[:, 0], vals[:, 0], offs)
            if not (np.array_equal(gc, exp) and np.array_equal(gv, cnt[exp])):
                ok, detail = False, "query %d differs from the oracle on rank %d" % (j, rank)
                break
    except Exception as e:  # noqa: BLE001
        ok, detail = False, repr(e)
    with open(os.path.join(outdir, "rank%d.json" % rank), "w") as f:
        json.dump({"ok": ok, "detail": detail, "queries": len(sizes), "world": world, "rank": rank}, f)
    dist.barrier()
    ex.close()
    ix.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
