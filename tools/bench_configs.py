"""Kernel-level measurements of the other BASELINE.json configs on device-resident columns:
Config 1 (filter + project, F64 and DEC), TPC-H Q6 (F64 and DEC) and TPC-H Q1 F64.  Prints one JSON line per case:
rows/s, algorithmic GB/s of the fused pipeline kernel (CUDA events recorded by the library) and the fraction of the
measured HBM copy peak.  Not the headline bench (bench.py is); numbers go to DESIGN.md / profiles/."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "datafusion-comet_b200")]
import torch
import bench
from comet_b200 import native, proto as P, tpch

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=600_037_902)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--device", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda", args.device)
torch.cuda.set_device(dev)
n = args.rows
peak, _ = bench.measured_peak()
cols = bench.gen_device(torch, n, 42, dev)
sel1 = float((cols["l_shipdate"] < tpch.DATE_1998_09_02).float().mean().item())


def table(variant, names, money):
    m = tpch.D12 if variant == "dec" else P.DOUBLE
    w = 16 if variant == "dec" else 8
    t = native.DeviceTable(n)
    for k in names:
        if k == "l_shipdate":
            t.add(P.DATE, cols[k].data_ptr(), 4, keep=cols[k])
        else:
            t.add(m, money[k].data_ptr(), w, keep=money[k])
    return t


def run(name, variant, plan, names, money, bytes_per_row, final=None):
    best = None
    for it in range(args.steps + 2):
        t = table(variant, names, money)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with native.Plan(plan, [t], config={"spark.comet.b200.chunkRows": str(1 << 30)}, device=args.device) as p:
            if final is None:
                rows_out, _cols = p.execute_device()   # results stay on the device (no D2H of the ~9 GB projection)
                out = None
            else:
                out, rows_out = p.collect(), None
            st = p.stats()
        if final is not None:
            with native.Plan(final, [out], device=args.device) as p2:
                p2.collect()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= 2:
            rec = (dt, st["pipeline_ms"], st["pipeline_launches"], rows_out)
            best = rec if best is None or rec[0] < best[0] else best
    dt, ms, launches, rows_out = best
    gbs = bytes_per_row * n / (ms * 1e-3) / 1e9   # algorithmic bytes over the summed duration of every pipeline launch of the plan
    print(json.dumps({"case": name, "rows": n, "rows_per_s_wall": n / dt, "pipeline_kernels_ms": ms, "launches": launches, "algorithmic_bytes_per_row": bytes_per_row,
                      "achieved_GBps": gbs, "frac_of_measured_peak": gbs / peak, "rows_out": rows_out}), flush=True)


for variant in ("f64", "dec"):
    money = bench.build_columns(torch, {k: cols[k] for k in ("l_quantity", "l_extendedprice", "l_discount", "l_tax")}, variant)
    w = 16 if variant == "dec" else 8
    run(f"config1_{variant}", variant, tpch.config1_plan(variant), ["l_quantity", "l_extendedprice", "l_shipdate"], money, 4 + 2 * w + w * sel1)
    run(f"q6_{variant}", variant, tpch.q6_partial_plan(variant), ["l_quantity", "l_extendedprice", "l_discount", "l_shipdate"], money, 4 + 3 * w, final=tpch.q6_final_plan(variant))
    del money
    torch.cuda.empty_cache()
