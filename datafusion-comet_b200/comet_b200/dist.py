"""Multi-GPU host logic: one process per GPU, partitions round-robin, partial aggregate state gathered on
rank 0 (low-cardinality / ungrouped aggregates need no all-to-all -- the N tiny state
tables are merged by the Final plan with the accumulators' merge_batch semantics, not by an allreduce).
Works over any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
import io

import pyarrow as pa


def table_to_bytes(tbl):
    sink = io.BytesIO()
    with pa.ipc.new_stream(sink, tbl.schema) as w:
        w.write_table(tbl)
    return sink.getvalue()


def table_from_bytes(b):
    return pa.ipc.open_stream(io.BytesIO(b)).read_all()


def gather_tables(tbl, dist=None, dst=0):
    """Gather each rank's (small) Arrow table on `dst`; returns the list there, None elsewhere."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [tbl]
    payload = table_to_bytes(tbl) if tbl is not None else b""
    objs = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(payload, objs, dst=dst)
    if dist.get_rank() != dst:
        return None
    return [table_from_bytes(o) for o in objs if o]


def partition_bounds(n_rows, rank, world):
    """Contiguous row range of `rank` when `n_rows` are split round-robin by row group of equal size."""
    per = (n_rows + world - 1) // world
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


# ---- the one real exchange step: hash-repartitioned partial aggregate state -------------------
class _DevPtr:
    """zero-copy torch view of a raw device pointer"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_bytes(torch, ptr, nbytes, device):
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevPtr(ptr, nbytes), device=device)


def exchange_partitions(torch, dist, device, columns, starts):
    """All-to-all of the per-destination row segments a ShuffleWriter plan left on the device.

    columns: list of (values_ptr, width_bytes, validity_bytes_ptr_or_None); rows of destination p are
    [starts[p], starts[p+1]).  Returns (n_recv, [(values_tensor, validity_bytes_tensor_or_None)]).
    NCCL all_to_all over NVLink/NVSwitch; the state rows, not the raw rows, are what moves -- exactly what the
    reference ships between Partial and Final (shuffle of state columns)."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    send = [int(starts[p + 1] - starts[p]) for p in range(world)]
    if world == 1:  # owned copies: the producing plan may be released before the consumer runs
        n = send[0]
        return n, [(device_bytes(torch, v, n * w, device).clone(), device_bytes(torch, vb, n, device).clone() if vb else None) for v, w, vb in columns]
    sc = torch.tensor(send, dtype=torch.int64, device=device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc)
    recv = [int(x) for x in rc.tolist()]
    n_recv, n_send = sum(recv), sum(send)
    out = []
    for v, w, vb in columns:
        src = device_bytes(torch, v, n_send * w, device)
        dst = torch.empty(n_recv * w, dtype=torch.uint8, device=device)
        dist.all_to_all_single(dst, src, output_split_sizes=[c * w for c in recv], input_split_sizes=[c * w for c in send])
        dvb = None
        if vb:
            srcv = device_bytes(torch, vb, n_send, device)
            dvb = torch.empty(n_recv, dtype=torch.uint8, device=device)
            dist.all_to_all_single(dvb, srcv, output_split_sizes=recv, input_split_sizes=send)
        out.append((dst, dvb))
    return n_recv, out
