"""Digest of an `ncu --set full` report (text, committed under profiles/): per kernel launch the duration, DRAM bytes and throughput,
L2 atomics, achieved occupancy, top warp-stall reasons and the source lines where stall samples pile up.
usage: python tools/ncu_digest.py gpurun_out/x.ncu-rep [max_launches] > profiles/r2_x.txt      (needs ncu on PATH)"""
import csv, io, subprocess, sys

rep = sys.argv[1]
maxl = int(sys.argv[2]) if len(sys.argv) > 2 else 99
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    i = col.get(name)
    if i is None:   # ncu prefixes some metrics with their section ("FBSP.TriageCompute.dram__throughput...")
        for h, j in col.items():
            if h.endswith("." + name):
                i = j
                break
    if i is None or r[i] == "":
        return None
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return None


def unit(name):
    return units[col[name]] if name in col else ""


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def to_ms(v, u):
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)


print(f"# {rep}: ncu --set full --clock-control none (cold-cache, serialised; times are not bench values)")
stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
# per kernel name keep the `maxl` LONGEST launches (a scan launches the same kernel for every column; the tiny ones say nothing)
by_name = {}
for k, r in enumerate(rows[2:]):
    by_name.setdefault(r[col["Kernel Name"]], []).append((to_ms(val(r, "gpu__time_duration.sum") or 0.0, unit("gpu__time_duration.sum")), k))
keep = set()
for name, lst in by_name.items():
    keep.update(k for _, k in sorted(lst, reverse=True)[:maxl])
    print(f"# {name[:100]}: {len(lst)} launches captured, {sum(t for t, _ in lst):.3f} ms in total, longest {max(t for t, _ in lst):.3f} ms")
for k, r in enumerate(rows[2:]):
    name = r[col["Kernel Name"]]
    if k not in keep:
        continue
    t = to_ms(val(r, "gpu__time_duration.sum"), unit("gpu__time_duration.sum"))
    rd = to_bytes(val(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum"))
    wr = to_bytes(val(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum"))
    print(f"\n== launch {k}: {name}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}  regs/thread {r[col['launch__registers_per_thread']]}")
    print(f"   duration {t:.3f} ms   DRAM read {rd / 1e9:.3f} GB + write {wr / 1e9:.3f} GB = {(rd + wr) / 1e9:.3f} GB -> {(rd + wr) / 1e9 / (t * 1e-3):.0f} GB/s"
          f"   (dram {val(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed') or 0:.1f}% of ncu peak, L2 {val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed') or 0:.1f}%, SM {val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed') or 0:.1f}%)")
    extra = []
    for m, label in (("lts__t_sectors_op_atom.sum", "L2 atom sectors"), ("lts__t_sectors_op_red.sum", "L2 red sectors"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
                     ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("smsp__inst_executed.sum", "warp instructions")):
        v = val(r, m)
        if v is not None:
            extra.append(f"{label} {v:,.0f}" if v > 1000 else f"{label} {v:.1f}")
    print("   " + "; ".join(extra))
    st = sorted(((val(r, h) or 0.0, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall_cols), reverse=True)[:5]
    print("   stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in st))
    # source hot spots
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
    srows = [x for x in csv.reader(io.StringIO(src)) if len(x) > 4]
    if len(srows) < 3:
        continue
    sh = srows[0] if "Source" in srows[0] else srows[1]
    try:
        i_src, i_st = sh.index("Source"), sh.index("Warp Stall Sampling (All Samples)")
    except ValueError:
        continue
    data = []
    for x in srows:
        try:
            data.append((int(x[i_st]), x[i_src].strip()))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1
    data = data[: len(data) // 2] if len(data) % 2 == 0 and data[: len(data) // 2] == data[len(data) // 2:] else data   # the listing comes twice
    tot = sum(d[0] for d in data) or 1
    top = sorted(range(len(data)), key=lambda i: -data[i][0])[:6]
    print("   hottest SASS (share of stall samples): " + " | ".join(f"{100 * data[i][0] / tot:.1f}% {data[i][1][:44]}" for i in sorted(top)))
