"""Host-overhead trace of one bench step (CB200_TRACE=1 prints the library's spans)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "datafusion-comet_b200")]
import torch, pyarrow as pa
import bench
from comet_b200 import native, proto as P, tpch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
dev = torch.device("cuda", 0)
cols = bench.gen_device(torch, n, 42, dev)
money = bench.build_columns(torch, cols, "dec")
pp, fp = tpch.q1_partial_plan("dec"), tpch.q1_final_plan("dec")
for it in range(4):
    t0 = time.perf_counter()
    table = bench.bind_table(native, P, tpch, "dec", n, money, cols)
    t1 = time.perf_counter()
    p = native.Plan(pp, [table], config={"spark.comet.b200.chunkRows": str(1 << 30)})
    t2 = time.perf_counter()
    state = p.collect()
    t3 = time.perf_counter()
    p.release()
    t4 = time.perf_counter()
    p2 = native.Plan(fp, [state])
    t5 = time.perf_counter()
    res = p2.collect()
    t6 = time.perf_counter()
    p2.release()
    t7 = time.perf_counter()
    print(f"[py] iter {it}: bind {1e3*(t1-t0):.2f} create {1e3*(t2-t1):.2f} collect {1e3*(t3-t2):.2f} release {1e3*(t4-t3):.2f} | final create {1e3*(t5-t4):.2f} collect {1e3*(t6-t5):.2f} release {1e3*(t7-t6):.2f} | total {1e3*(t7-t0):.2f} ms", file=sys.stderr)
