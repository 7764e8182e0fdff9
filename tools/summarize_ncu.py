"""Turn ncu outputs into the text summaries committed under profiles/.
  launch list : python tools/summarize_ncu.py launches gpurun_out/x.csv > profiles/r1_x_launches.txt
  full capture: python tools/summarize_ncu.py full gpurun_out/x.ncu-rep > profiles/r1_x_full.txt   (needs ncu on PATH)"""
import collections, csv, subprocess, sys

mode, path = sys.argv[1], sys.argv[2]
if mode == "launches":
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    total = 0.0
    seq = []
    mi = hdr.index("Metric Name") if "Metric Name" in hdr else -1
    for r in rows[1:]:
        if mi >= 0 and r[mi] != "gpu__time_duration.sum":
            continue                                     # the csv may carry more metrics per launch (DRAM bytes): this table is about time
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v     # -> us
        k = r[ki]
        ours = k.startswith("cb_") or "cb200::" in k or k.startswith("k_")
        a = agg.setdefault(k[:90], [0, 0.0, 0.0, ours])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v)
        total += v
        if ours:
            seq.append((k[:60], r[gi], r[bi], v))
    print(f"# {path}: {len(rows) - 1} kernel launches, {total / 1e3:.3f} ms of kernel time (ncu gpu__time_duration.sum, --clock-control none; cold-cache, serialised)")
    print(f"# {'kernel':90s} {'n':>5s} {'total ms':>10s} {'max ms':>9s} {'share':>7s}  ours")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {k:90s} {a[0]:5d} {a[1] / 1e3:10.3f} {a[2] / 1e3:9.3f} {100 * a[1] / total:6.1f}%  {'*' if a[3] else ''}")
    ours_total = sum(a[1] for a in agg.values() if a[3])
    print(f"# comet_b200 kernels: {ours_total / 1e3:.3f} ms = {100 * ours_total / total:.1f}% of all kernel time (the rest is torch's synthetic-data generation and checks)")
    print("# launch sequence of comet_b200 kernels (first 80):")
    for k, g, b, v in seq[:80]:
        print(f"  {k:60s} grid {g:>16s} block {b:>14s} {v:10.1f} us")
else:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
            "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__cycles_active.avg",
            "sm__cycles_elapsed.avg.per_second", "gpc__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second"]
    want += [n for n in h if n.startswith("smsp__pcsamp_warps_issue_stalled") and not n.endswith("_not_issued")]
    for r in rows[2:]:
        print(f"# ncu --set full capture: {path}")
        for n in want:
            if n in h:
                i = h.index(n)
                print(f"  {n:75s} {r[i]:>22s} {units[i]}")
        try:
            t = float(r[h.index('gpu__time_duration.sum')].replace(',', ''))
            tu = units[h.index('gpu__time_duration.sum')]
            t_s = t * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}.get(tu, 1e-6)
            def by(name):
                v = float(r[h.index(name)].replace(',', '')); u = units[h.index(name)]
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(u, 1)
            rd, wr = by('dram__bytes_read.sum'), by('dram__bytes_write.sum')
            if len(sys.argv) > 3:  # also record it for bench.py's roofline.traffic:  full <rep> <json-out> <rows> <variant>
                import json
                json.dump({"cb_pipeline_agg": {"dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr), "rows": int(sys.argv[4]), "variant": sys.argv[5],
                                               "duration_ms_under_ncu": t_s * 1e3, "source": path}}, open(sys.argv[3], "w"), indent=1)
            print(f"  => dram traffic {rd + wr:.0f} B per launch ({(rd + wr) / 1e9:.3f} GB; read {rd / 1e9:.3f} + write {wr / 1e9:.3f}), {(rd + wr) / t_s / 1e9:.1f} GB/s under the profiler")
        except Exception as e:
            print("  (no dram summary:", e, ")")
