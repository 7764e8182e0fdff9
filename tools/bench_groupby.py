#!/usr/bin/env python
"""Config 4 of BASELINE.json: high-cardinality GROUP BY l_orderkey SUM(l_extendedprice) with the hash
repartition of partial state across the GPUs of one box (NCCL all-to-all over NVLink).

Per rank: HashAggregate(Partial) [fused scan+hash-agg kernel] -> ShuffleWriter(HashPartition l_orderkey, N)
[murmur3/pmod + counting sort on device] -> all-to-all of state rows -> HashAggregate(Final).
Prints one JSON line on rank 0 (not the driver's bench contract: that is bench.py / Q1)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "datafusion-comet_b200")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=200_000_000)  # per GPU
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--variant", default="dec")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import torch, torch.distributed as dist
    from comet_b200 import native, proto as P
    from comet_b200.dist import exchange_partitions
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.rows
    g = torch.Generator(device=dev); g.manual_seed(100 + rank)
    # clustered keys: ~4 lines per order; rank r owns orders {r, r+world, ...} so partitions hold disjoint keys
    lines = torch.randint(1, 8, (n // 3 + 8,), generator=g, device=dev)
    order = torch.repeat_interleave(torch.arange(lines.shape[0], device=dev, dtype=torch.int64), lines)[:n].contiguous()
    keys = (order * world + rank).contiguous()
    cents = (torch.randint(1, 51, (n,), generator=g, device=dev) * torch.randint(90000, 210001, (n,), generator=g, device=dev)).contiguous()
    del lines, order
    dec = args.variant == "dec"
    if dec:
        val = torch.empty((n, 2), dtype=torch.int64, device=dev); val[:, 0] = cents; val[:, 1] = 0
        m, sdt, w = P.DECIMAL(12, 2), P.DECIMAL(22, 2), 16
    else:
        val = (cents.to(torch.float64) / 100.0).contiguous(); m, sdt, w = P.DOUBLE, P.DOUBLE, 8
    agg = P.hash_agg(P.scan([P.INT64, m]), [P.bound(0, P.INT64)], [P.agg_sum(P.bound(1, m), sdt)], P.PARTIAL)
    map_plan = P.shuffle_writer(agg, P.hash_partitioning([P.bound(0, P.INT64)], world))
    state_types = [P.INT64, sdt, P.BOOL] if dec else [P.INT64, sdt]
    final_plan = P.hash_agg(P.scan(state_types, source="shuffle"), [P.bound(0, P.INT64)], [P.agg_sum(P.unbound("c", m), sdt)], P.FINAL)
    widths = [8, 16, 1] if dec else [8, 8]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step():
        t = native.DeviceTable(n)
        t.add(P.INT64, keys.data_ptr(), 8, keep=keys)
        t.add(m, val.data_ptr(), w, keep=val)
        t0 = time.perf_counter()
        p = native.Plan(map_plan, [t], config={"spark.comet.b200.chunkRows": str(1 << 31), "spark.comet.b200.hashThreads": os.environ.get("CB200_HASH_THREADS", "512")}, device=local)
        rows, cols = p.execute_device()
        starts = p.partition_starts()
        st = p.stats()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        spec = []
        for i, c in enumerate(cols):
            vptr = c.bool_bytes if state_types[i].name == "BOOL" else c.values
            spec.append((vptr, widths[i], c.validity_bytes))
        n_recv, recvd = exchange_partitions(torch, dist if world > 1 else None, dev, spec, starts)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        p.release()
        rt = native.DeviceTable(n_recv)
        for i, (vt, vb) in enumerate(recvd):
            rt.add_bytes(state_types[i], vt.data_ptr(), widths[i], vb.data_ptr() if vb is not None else None, keep=(vt, vb))
        p2 = native.Plan(final_plan, [rt], config={"spark.comet.b200.chunkRows": str(1 << 31), "spark.comet.b200.hashThreads": os.environ.get("CB200_HASH_THREADS", "512")}, device=local)
        out = p2.execute_device()
        n_groups = out[0] if out else 0
        st2 = p2.stats()
        result = None
        if args.check:
            import ctypes as C
            rows_out, oc = out
            kk = torch.as_tensor(__import__("comet_b200.dist", fromlist=["_DevPtr"])._DevPtr(oc[0].values, rows_out * 8), device=dev).view(torch.int64).clone()
            vv = torch.as_tensor(__import__("comet_b200.dist", fromlist=["_DevPtr"])._DevPtr(oc[1].values, rows_out * (16 if dec else 8)), device=dev)
            vv = (vv.view(torch.int64).view(-1, 2)[:, 0] if dec else vv.view(torch.float64)).clone()
            result = (kk, vv)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        p2.release()
        return dict(partial=t1 - t0, exchange=t2 - t1, final=t3 - t2, rows_state=rows, n_recv=n_recv, n_groups=n_groups, st=st, st2=st2, result=result)

    for _ in range(args.warmup):
        step()
    sync()
    t0 = time.perf_counter()
    acc = dict(partial=0.0, exchange=0.0, final=0.0)
    last = None
    for _ in range(args.steps):
        last = step()
        for k in acc:
            acc[k] += last[k]
    sync()
    elapsed = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([elapsed], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); elapsed = float(t.item())
    ok = None
    if args.check:
        kk, vv = last["result"]
        # owner check: every key this rank ended up with hashes to this rank; sums checked against torch
        exp = torch.zeros(int(keys.max().item()) // world + 2, dtype=torch.int64 if dec else torch.float64, device=dev)
        if world == 1:
            exp.scatter_add_(0, keys // world, cents if dec else val)
            ok = bool((exp[kk // world] == vv).all().item()) if dec else bool(torch.allclose(exp[kk // world], vv, rtol=1e-12))
    gsum = torch.tensor([last["n_groups"]], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(gsum)
    if rank == 0:
        bytes_state = sum(widths) + (len(widths))  # values + validity bytes per state row
        print(json.dumps({"metric": "rows/sec GROUP BY l_orderkey SUM(l_extendedprice)", "value": world * n * args.steps / elapsed, "unit": "rows/s",
                          "n_gpus": world, "rows_per_gpu": n, "groups_total": int(gsum.item()), "ms_per_step": 1e3 * elapsed / args.steps,
                          "phase_ms": {k: 1e3 * v / args.steps for k, v in acc.items()},
                          "partial_kernel_ms": last["st"]["pipeline_ms"], "final_kernel_ms": last["st2"]["pipeline_ms"],
                          "state_rows_per_gpu": last["rows_state"], "exchange_GBps_per_gpu": last["rows_state"] * bytes_state / max(acc["exchange"] / args.steps, 1e-9) / 1e9,
                          "variant": args.variant, "checked": ok}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
