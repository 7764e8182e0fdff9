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


# ---- helpers for callers that look at the library's device buffers through torch (bench.py).  The exchange step itself (hash-repartitioned
#      partial aggregate state) lives in the library: csrc/exchange.cpp, native.Comm.exchange --------------------------------------
class _DevPtr:
    """zero-copy torch view of a raw device pointer"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_bytes(torch, ptr, nbytes, device):
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevPtr(ptr, nbytes), device=device)
