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
