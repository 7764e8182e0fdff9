"""Host-side trace of the Parquet e2e leg (CB200_TRACE=1 prints the library's spans): where a batch's time goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "datafusion-comet_b200")]
import numpy as np, torch, pyarrow as pa, pyarrow.parquet as pq
import bench
from comet_b200 import native, tpch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 << 20
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
comp = sys.argv[3] if len(sys.argv) > 3 else "NONE"
dev = torch.device("cuda", 0)
cols = bench.gen_device(torch, n, 42, dev)
money = bench.build_columns(torch, cols, "dec")
batches, host = bench.host_arrow_batches(torch, pa, tpch, "dec", money, cols, 1 << 22)
tbl = pa.Table.from_batches(batches)
sink = pa.BufferOutputStream()
pq.write_table(tbl, sink, row_group_size=1 << 20, compression=comp, use_dictionary=True, data_page_version="1.0", store_decimal_as_integer=True)
buf = sink.getvalue()
h = torch.empty(buf.size, dtype=torch.uint8, pin_memory=True)
h.numpy()[:] = np.frombuffer(buf, dtype=np.uint8)
f = native.register_memory_file("trace-li", h)
plan = tpch.q1_partial_plan("dec", scan=tpch.q1_native_scan("dec", [f]))
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with native.Plan(plan, [], config={"spark.comet.b200.chunkRows": str(chunk)}) as p:
        p.collect()
        st = p.stats()
    t1 = time.perf_counter()
    print(f"[py] iter {it}: {1e3 * (t1 - t0):.1f} ms, {n / (t1 - t0) / 1e9:.2f} G rows/s, file {buf.size / 1e6:.0f} MB, launches {st['kernel_launches']}, pipeline_ms {st.get('pipeline_ms')}", file=sys.stderr)
