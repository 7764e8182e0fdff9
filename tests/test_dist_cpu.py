"""N>1 host logic on CPU: world_size-2 gloo -- partition the rows, gather per-rank partial state tables on
rank 0 and merge them with the accumulators' merge semantics (what bench.py does over NCCL)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "datafusion-comet_b200")]
    import pyarrow as pa
    import torch.distributed as dist
    from comet_b200 import tpch
    from comet_b200.dist import gather_tables, partition_bounds
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 40_000
    cols = tpch.gen_lineitem(n, seed=77)
    lo, hi = partition_bounds(n, rank, world)
    keep = cols["l_shipdate"][lo:hi] <= tpch.Q1_CUTOFF
    gid = (cols["l_returnflag"][lo:hi].astype(np.int64) * 2 + cols["l_linestatus"][lo:hi])[keep]
    acc = oracle.SumDecimalGroups(6, 22)
    acc.update(oracle.dec_from_i64(cols["l_quantity"][lo:hi][keep]), None, gid)
    s, sv, e = acc.state()
    state = pa.table({"gid": np.arange(6), "sum_lo": s[:, 0].astype(np.uint64), "sum_hi": s[:, 1].astype(np.uint64), "sum_valid": sv, "is_empty": e})
    tables = gather_tables(state, dist, 0)
    if rank == 0:
        assert len(tables) == world
        fin = oracle.SumDecimalGroups(6, 22)
        for t in tables:
            ss = np.stack([t["sum_lo"].to_numpy(), t["sum_hi"].to_numpy()], axis=1).astype(np.uint64)
            fin.merge(ss, t["sum_valid"].to_numpy(), t["is_empty"].to_numpy(), t["gid"].to_numpy())
        out, outv = fin.evaluate()
        keep_all = cols["l_shipdate"] <= tpch.Q1_CUTOFF
        gid_all = cols["l_returnflag"].astype(np.int64) * 2 + cols["l_linestatus"]
        exp = [int(cols["l_quantity"][keep_all & (gid_all == k)].sum()) if (keep_all & (gid_all == k)).any() else None for k in range(6)]
        q.put(oracle.dec_to_ints(out, outv) == exp)
    else:
        assert tables is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_merge():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_partition_bounds_cover_everything():
    from comet_b200.dist import partition_bounds
    for n in (0, 1, 7, 1000, 600_037_902):
        for w in (1, 2, 4, 8):
            spans = [partition_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def _exchange_worker(rank, world, port, q):
    """The hash-repartition exchange on CPU: map side = oracle murmur3 / pmod / stable counting sort (what PartitionNode does on the
    device), counts all-gathered, receive layout from the LIBRARY's cb200_exchange_layout, payload by all_to_all -- then the Final
    merge, compared with a single-process group-by."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "datafusion-comet_b200")]
    import torch
    import torch.distributed as dist
    from comet_b200 import native, tpch
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 60_000
    cols = tpch.gen_lineitem(n, seed=91)
    lo, hi = n * rank // world, n * (rank + 1) // world
    keys, vals = cols["l_orderkey"][lo:hi], cols["l_extendedprice"][lo:hi]
    uk, inv = np.unique(keys, return_inverse=True)                       # Partial aggregate of this partition
    psum = np.zeros(len(uk), dtype=np.int64)
    np.add.at(psum, inv, vals)
    hashes = oracle.murmur3_column("i64", uk)
    pids, starts, row_idx = oracle.partition_rows(hashes, world)         # ShuffleWriter: stable by partition
    send_k, send_v = uk[row_idx], psum[row_idx]
    counts = torch.tensor([int(starts[p + 1] - starts[p]) for p in range(world)], dtype=torch.int64)
    allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, counts)
    matrix = torch.stack(allc).flatten().tolist()
    total, rc, ro = native.exchange_layout(matrix, world, rank)          # library: receive counts / offsets
    assert total == sum(rc) and ro == [sum(rc[:i]) for i in range(world)]
    rk, rv = np.zeros(total, dtype=np.int64), np.zeros(total, dtype=np.int64)

    def move(src, dst):  # grouped send / recv, the shape of cb200_exchange's ncclGroup (gloo has no all_to_all)
        ops, bufs = [], {}
        for p in range(world):
            seg = torch.from_numpy(src[starts[p]:starts[p + 1]].copy())
            if p == rank:
                dst[ro[p]:ro[p] + rc[p]] = seg.numpy()
                continue
            if seg.numel():
                ops.append(dist.P2POp(dist.isend, seg, p))
            if rc[p]:
                bufs[p] = torch.zeros(rc[p], dtype=torch.int64)
                ops.append(dist.P2POp(dist.irecv, bufs[p], p))
        for r in (dist.batch_isend_irecv(ops) if ops else []):
            r.wait()
        for p, b in bufs.items():
            dst[ro[p]:ro[p] + rc[p]] = b.numpy()
    move(send_k, rk)
    move(send_v, rv)
    assert rk.shape[0] == total
    own = oracle.murmur3_column("i64", rk)
    assert all(oracle.pmod(int(h), world) == rank for h in own[:2000])  # every received key belongs to this rank
    fk, finv = np.unique(rk, return_inverse=True)                        # Final merge
    fsum = np.zeros(len(fk), dtype=np.int64)
    np.add.at(fsum, finv, rv)
    gathered = [None] * world
    dist.all_gather_object(gathered, (fk.tolist(), fsum.tolist()))
    if rank == 0:
        got = {}
        for ks, vs in gathered:
            for k, v in zip(ks, vs):
                assert k not in got                                       # a key lives on exactly one rank
                got[k] = v
        ek, einv = np.unique(cols["l_orderkey"], return_inverse=True)
        es = np.zeros(len(ek), dtype=np.int64)
        np.add.at(es, einv, cols["l_extendedprice"])
        q.put(got == dict(zip(ek.tolist(), es.tolist())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_hash_repartition_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_exchange_layout_matches_a_simulated_all_to_all():
    from comet_b200 import native
    rng = np.random.default_rng(3)
    for world in (1, 2, 4, 8):
        counts = rng.integers(0, 50, (world, world))
        for me in range(world):
            total, rc, ro = native.exchange_layout(counts.flatten().tolist(), world, me)
            assert rc == counts[:, me].tolist() and total == int(counts[:, me].sum())
            assert ro == [int(counts[:s, me].sum()) for s in range(world)]


def test_oracle_groupby_matches_numpy():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "datafusion-comet_b200")]
    from comet_b200 import tpch
    from oracle import oracle
    oracle.build()
    cols = tpch.gen_lineitem(200_000, seed=17)
    k, s = oracle.groupby_sum_dec(cols["l_orderkey"], oracle.dec_from_i64(cols["l_extendedprice"]), 3, want_rows=True)
    ek, einv = np.unique(cols["l_orderkey"], return_inverse=True)
    es = np.zeros(len(ek), dtype=np.int64)
    np.add.at(es, einv, cols["l_extendedprice"])
    assert dict(zip(k.tolist(), s)) == dict(zip(ek.tolist(), es.tolist()))
