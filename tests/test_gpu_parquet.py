"""GPU parity: NativeScan -- Parquet pages decoded on the device -- feeding the fused Q1 / Config 1 pipelines."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RF = ["A", "N", "R"]
LS = ["F", "O"]


@pytest.fixture(scope="module")
def cb():
    import comet_b200
    return comet_b200


def unscaled(x):
    return None if x is None else int(x.scaleb(-x.as_tuple().exponent))


def run(cb, plan, inputs=(), chunk_rows=None):
    cfg = {"spark.comet.b200.chunkRows": str(chunk_rows)} if chunk_rows else None
    with cb.native.Plan(plan, list(inputs), config=cfg) as p:
        t = p.collect()
        st = p.stats()
    return t, st


@pytest.mark.parametrize("as_int,memory,chunk", [(True, False, None), (False, False, 100_000), (True, True, 40_000)])
def test_q1_dec_from_parquet(cb, oracle, tmp_path, as_int, memory, chunk):
    t = cb.tpch
    n = 300_000
    cols = t.gen_lineitem(n, seed=31)
    half = n // 2
    paths = []
    for i, (lo, hi) in enumerate([(0, half), (half, n)]):   # two files, several row groups each
        part = {k: v[lo:hi] for k, v in cols.items()}
        paths.append(t.write_lineitem_parquet(part, str(tmp_path / f"li{i}.parquet"), "dec", row_group_size=32_768, decimal_as_int=as_int))
    if memory:
        paths = [cb.native.register_memory_file(f"t{i}", np.fromfile(p, dtype=np.uint8)) for i, p in enumerate(paths)]
    plan = t.q1_partial_plan("dec", scan=t.q1_native_scan("dec", paths))
    state, st = run(cb, plan, chunk_rows=chunk)
    assert st["h2d_bytes"] > 0 and st["h2d_bytes"] < n * 45          # encoded pages, not 70-byte Arrow rows
    res, _ = run(cb, t.q1_final_plan("dec"), [state])
    d = oracle.dec_from_i64
    exp = oracle.q1_dec(d(cols["l_quantity"]), d(cols["l_extendedprice"]), d(cols["l_discount"]), d(cols["l_tax"]), cols["l_shipdate"],
                        cols["l_returnflag"], cols["l_linestatus"], 3, 2, t.Q1_CUTOFF, 1)
    got = {(r["col_0"], r["col_1"]): r for r in res.to_pylist()}
    for k, e in enumerate(exp):
        if e is None:
            continue
        g = got[(RF[k // 2], LS[k % 2])]
        for j, name in enumerate(["sum_qty", "sum_base", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc"]):
            assert unscaled(g[f"col_{2 + j}"]) == e[name], (k, name)
        assert g["col_9"] == e["count"]


def test_config1_f64_from_parquet(cb, oracle, tmp_path):
    t = cb.tpch
    P = cb.proto
    n = 200_000
    cols = t.gen_lineitem(n, seed=33)
    names = ["l_quantity", "l_extendedprice", "l_shipdate"]
    path = t.write_lineitem_parquet(cols, str(tmp_path / "c1.parquet"), "f64", row_group_size=50_000, columns=names)
    fields = list(zip(names, t.config1_scan_fields("f64"), [True] * 3))
    sc = P.native_scan(fields, fields, [path])
    ship = P.bound(2, P.DATE)
    plan = P.projection(P.filter_(sc, P.lt(ship, P.literal(t.DATE_1998_09_02, P.DATE))), [P.multiply(P.bound(0, P.DOUBLE), P.bound(1, P.DOUBLE), P.DOUBLE)])
    res, _ = run(cb, plan, chunk_rows=120_000)
    q, p = cols["l_quantity"].astype(np.float64) / 100.0, cols["l_extendedprice"].astype(np.float64) / 100.0
    exp = oracle.filter_project_f64(q, p, cols["l_shipdate"], t.DATE_1998_09_02, 1)
    got = res.column(0).to_numpy()
    assert got.shape == exp.shape and (got.view(np.uint64) == exp.view(np.uint64)).all()


def test_dictionary_encoded_numeric_column(cb, tmp_path):
    """pyarrow's default dictionary-encodes low-cardinality numerics too: RLE_DICTIONARY with a gathered fixed-width dictionary."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    P = cb.proto
    n = 150_000
    rng = np.random.default_rng(2)
    a = rng.integers(0, 50, n).astype(np.int64) * 100
    b = rng.integers(0, 11, n).astype(np.int32)
    path = str(tmp_path / "dict.parquet")
    pq.write_table(pa.table({"a": a, "b": pa.array(b, type=pa.int32())}), path, row_group_size=40_000, compression="NONE", use_dictionary=True,
                   data_page_version="1.0")
    fields = [("a", P.INT64, True), ("b", P.INT32, True)]
    sc = P.native_scan(fields, fields, [path])
    plan = P.hash_agg(sc, [], [P.agg_sum(P.bound(0, P.INT64), P.INT64), P.agg_sum(P.bound(1, P.INT32), P.INT64), P.agg_count([P.literal(1, P.INT32)])], P.PARTIAL)
    res, _ = run(cb, plan)
    r = res.to_pylist()[0]
    assert r["col_0"] == int(a.sum()) and r["col_1"] == int(b.sum()) and r["col_2"] == n


def _nullable_table(n, seed, null_frac=0.15):
    import decimal
    import pyarrow as pa
    rng = np.random.default_rng(seed)
    ctx = decimal.Context(prec=60)

    def mask():
        m = rng.random(n) < null_frac
        m[: n // 50] = True                      # a long NULL stretch -> RLE level runs, pages with few / no values
        m[n // 2: n // 2 + n // 40] = False      # and a long all-valid one
        return m
    i32 = rng.integers(-2**31, 2**31, n).astype(np.int32)
    i64 = rng.integers(-2**62, 2**62, n)
    f64 = rng.standard_normal(n)
    d12 = rng.integers(-10**11, 10**11, n)
    d30 = [int(a) * 10**12 + int(b) for a, b in zip(rng.integers(-10**17, 10**17, n), rng.integers(0, 10**12, n))]
    lowcard = rng.integers(0, 40, n).astype(np.int64) * 1000
    words = np.array(["AIR", "MAIL", "SHIP", "TRUCK", "RAIL", "REG AIR", "FOB", ""])[rng.integers(0, 8, n)]
    ms = [mask() for _ in range(7)]
    dec = lambda vals, p, s, m: pa.array([None if mm else decimal.Decimal(int(v)).scaleb(-s, context=ctx) for v, mm in zip(vals, m)], type=pa.decimal128(p, s))
    return pa.table({
        "i32": pa.array(i32, mask=ms[0]), "i64": pa.array(i64, mask=ms[1]), "f64": pa.array(f64, mask=ms[2]), "d12": dec(d12, 12, 2, ms[3]),
        "d30": dec(d30, 30, 4, ms[4]), "low": pa.array(lowcard, mask=ms[5]), "word": pa.array(words.tolist(), mask=ms[6]).cast(pa.string()),
        "req": pa.array(i64),                    # no NULLs: statistics say so -> verify-only fast path in the same file
    })


def _scan_all(cb, tbl, paths):
    P = cb.proto
    dts = {"i32": P.INT32, "i64": P.INT64, "f64": P.DOUBLE, "d12": P.DECIMAL(12, 2), "d30": P.DECIMAL(30, 4), "low": P.INT64, "word": P.STRING, "req": P.INT64}
    fields = [(name, dts[name], True) for name in tbl.column_names]
    sc = P.native_scan(fields, fields, paths)
    return P.projection(sc, [P.bound(i, dts[name]) for i, name in enumerate(tbl.column_names)])


@pytest.mark.parametrize("compression,version,dictionary,dec_as_int", [("NONE", "1.0", False, True), ("SNAPPY", "1.0", True, True), ("SNAPPY", "2.0", False, False),
                                                                       ("NONE", "2.0", True, False), ("SNAPPY", "1.0", ["word", "low"], False)])
def test_nulls_and_snappy_roundtrip(cb, tmp_path, compression, version, dictionary, dec_as_int):
    """Every value and every NULL of a pyarrow-written file comes back: definition levels -> validity + scatter, Snappy pages
    decompressed on the device, data page v1 / v2, PLAIN and dictionary encodings, INT64- and FLBA-backed decimals."""
    import pyarrow.parquet as pq
    n = 70_000
    tbl = _nullable_table(n, seed=5)
    path = str(tmp_path / "nulls.parquet")
    use_dict = dictionary if dictionary is not False else ["word"]    # strings are only decoded from dictionary pages
    pq.write_table(tbl, path, row_group_size=20_000, compression=compression, use_dictionary=use_dict, data_page_version=version, data_page_size=8192,
                   store_decimal_as_integer=dec_as_int)
    res, st = run(cb, _scan_all(cb, tbl, [path]), chunk_rows=45_000)
    assert res.num_rows == n
    for j, name in enumerate(tbl.column_names):
        got, want = res.column(j).combine_chunks(), tbl.column(name).combine_chunks()
        assert got.null_count == want.null_count, name
        if name == "f64":
            gv, wv = got.to_numpy(zero_copy_only=False), want.to_numpy(zero_copy_only=False)
            ok = ~np.isnan(wv)
            assert (np.isnan(gv) == np.isnan(wv)).all() and (gv[ok].view(np.uint64) == wv[ok].view(np.uint64)).all(), name
        else:
            assert got.cast(want.type).equals(want), name


@pytest.mark.parametrize("compression,version", [("ZSTD", "1.0"), ("ZSTD", "2.0"), ("LZ4", "1.0"), ("GZIP", "2.0")])
def test_host_decompressed_codecs_roundtrip(cb, tmp_path, compression, version):
    """ZSTD (the codec of the reference's benchmark data, benchmarks/results/0.16.0/comet-tpch.json:8), LZ4 and GZIP pages: the bytes are
    decompressed on the host (csrc/host_codecs.cpp), levels / dictionaries / values are decoded by the same device kernels."""
    import pyarrow.parquet as pq
    n = 70_000
    tbl = _nullable_table(n, seed=8)
    path = str(tmp_path / "codec.parquet")
    pq.write_table(tbl, path, row_group_size=25_000, compression=compression, use_dictionary=["word", "low"], data_page_version=version, data_page_size=16384,
                   store_decimal_as_integer=True)
    res, st = run(cb, _scan_all(cb, tbl, [path]), chunk_rows=40_000)
    assert res.num_rows == n
    for j, name in enumerate(tbl.column_names):
        got, want = res.column(j).combine_chunks(), tbl.column(name).combine_chunks()
        assert got.null_count == want.null_count, name
        if name == "f64":
            gv, wv = got.to_numpy(zero_copy_only=False), want.to_numpy(zero_copy_only=False)
            ok = ~np.isnan(wv)
            assert (np.isnan(gv) == np.isnan(wv)).all() and (gv[ok].view(np.uint64) == wv[ok].view(np.uint64)).all(), name
        else:
            assert got.cast(want.type).equals(want), name


@pytest.mark.parametrize("compression,version", [("NONE", "1.0"), ("SNAPPY", "2.0"), ("ZSTD", "1.0")])
def test_plain_encoded_string_pages(cb, tmp_path, compression, version):
    """A string column without (or after falling back from) dictionary encoding: PLAIN BYTE_ARRAY pages.  The host turns their values
    into codes of the plan-wide dictionary, the device scatters the codes like any PLAIN INT32 page; mixed with dictionary pages of the
    same column (second file) the codes agree."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    P = cb.proto
    n = 50_000
    rng = np.random.default_rng(4)
    words = np.array([f"w{i:05d}" for i in range(3000)] + ["", "ünï", "a" * 300])[rng.integers(0, 3003, n)]
    mask = rng.random(n) < 0.1
    tbl = pa.table({"word": pa.array(words.tolist(), mask=mask).cast(pa.string()), "req": pa.array(np.arange(n, dtype=np.int64))})
    p1, p2 = str(tmp_path / "plain.parquet"), str(tmp_path / "dict.parquet")
    pq.write_table(tbl, p1, row_group_size=20_000, compression=compression, use_dictionary=False, data_page_version=version, data_page_size=8192)
    pq.write_table(tbl, p2, row_group_size=20_000, compression=compression, use_dictionary=["word"], data_page_version=version)
    fields = [("word", P.STRING, True), ("req", P.INT64, True)]
    plan = P.projection(P.native_scan(fields, fields, [p1, p2]), [P.bound(0, P.STRING), P.bound(1, P.INT64)])
    res, _ = run(cb, plan, chunk_rows=30_000)
    assert res.num_rows == 2 * n
    want = tbl.column("word").to_pylist()
    assert res.column(0).to_pylist() == want + want
    assert res.column(1).to_pylist() == list(range(n)) * 2
    # and as a group key: COUNT(*) per word over the PLAIN file
    agg = P.hash_agg(P.native_scan(fields, fields, [p1]), [P.bound(0, P.STRING)], [P.agg_count([P.literal(1, P.INT32)])], P.PARTIAL)
    out, _ = run(cb, agg)
    exp = {}
    for w in want:
        exp[w] = exp.get(w, 0) + 1
    got = {}
    for r in out.to_pylist():
        got[r["col_0"]] = got.get(r["col_0"], 0) + r["col_1"]
    assert got == exp


def test_snappy_pages_of_a_megabyte_are_decoded_by_segments(cb, tmp_path):
    """1 MB data pages (pyarrow's default size) of Snappy: 16 segments of 64 KB per page, every kind of stream the Q1 columns produce
    (PLAIN INT64 decimals = a literal + copy pair per value, random doubles = 64 KB literals, bit-packed dictionary indices)."""
    import pyarrow.parquet as pq
    n = 1_500_000
    tbl = _nullable_table(n, seed=21, null_frac=0.02)
    path = str(tmp_path / "big_pages.parquet")
    pq.write_table(tbl, path, row_group_size=700_000, compression="SNAPPY", use_dictionary=["word", "low"], data_page_version="1.0", store_decimal_as_integer=True)
    res, st = run(cb, _scan_all(cb, tbl, [path]), chunk_rows=1_000_000)
    assert res.num_rows == n
    for j, name in enumerate(tbl.column_names):
        got, want = res.column(j).combine_chunks(), tbl.column(name).combine_chunks()
        assert got.null_count == want.null_count, name
        if name == "f64":
            gv, wv = got.to_numpy(zero_copy_only=False), want.to_numpy(zero_copy_only=False)
            ok = ~np.isnan(wv)
            assert (np.isnan(gv) == np.isnan(wv)).all() and (gv[ok].view(np.uint64) == wv[ok].view(np.uint64)).all(), name
        else:
            assert got.cast(want.type).equals(want), name


def test_aggregate_over_nullable_parquet_columns(cb, tmp_path):
    import pyarrow.compute as pc
    import pyarrow.parquet as pq
    P = cb.proto
    n = 120_000
    tbl = _nullable_table(n, seed=9, null_frac=0.3)
    path = str(tmp_path / "agg.parquet")
    pq.write_table(tbl, path, row_group_size=50_000, compression="SNAPPY", use_dictionary=["word", "low"], data_page_version="1.0", store_decimal_as_integer=True)
    fields = [("word", P.STRING, True), ("d12", P.DECIMAL(12, 2), True), ("i32", P.INT32, True)]
    sc = P.native_scan(fields, fields, [path])
    d = P.bound(1, P.DECIMAL(12, 2))
    part = P.hash_agg(sc, [P.bound(0, P.STRING)], [P.agg_sum(d, P.DECIMAL(22, 2)), P.agg_count([d]), P.agg_count([P.literal(1, P.INT32)]),
                                                    P.agg_sum(P.bound(2, P.INT32), P.INT64)], P.PARTIAL)
    res, _ = run(cb, part)
    got = {r["col_0"]: r for r in res.to_pylist()}
    words = tbl.column("word").to_pylist()
    d12 = tbl.column("d12").to_pylist()
    i32 = tbl.column("i32").to_pylist()
    exp = {}
    for w, dv, iv in zip(words, d12, i32):
        e = exp.setdefault(w, [None, 0, 0, None])
        if dv is not None:
            e[0] = dv if e[0] is None else e[0] + dv
            e[1] += 1
        e[2] += 1
        if iv is not None:
            e[3] = iv if e[3] is None else e[3] + iv
    assert set(got) == set(exp)                       # includes the NULL group
    for w, e in exp.items():
        g = got[w]
        assert g["col_1"] == e[0] and g["col_3"] == e[1] and g["col_4"] == e[2] and g["col_5"] == e[3], w


# ---- row-group pruning by statistics, file splits, annotations (round 2) ------------------------------------------------------
def _sorted_q6_files(cb, tmp_path, n, seed, variant):
    t = cb.tpch
    cols = t.gen_lineitem(n, seed=seed)
    order = np.argsort(cols["l_shipdate"], kind="stable")           # date-clustered, like a table written ORDER BY l_shipdate
    cols = {k: v[order] for k, v in cols.items()}
    path = t.write_lineitem_parquet(cols, str(tmp_path / f"q6_{variant}.parquet"), variant, row_group_size=16_384, columns=t.Q6_COLUMNS)
    return cols, path


@pytest.mark.parametrize("variant", ["dec", "f64"])
def test_q6_row_groups_pruned_by_min_max(cb, oracle, tmp_path, variant, monkeypatch):
    """parquet_exec.rs:143-196: row groups whose statistics rule the pushed predicate out are never read.  Q6 over a
    date-clustered file touches one year in seven; the result is the unpruned result bit for bit."""
    t = cb.tpch
    n = 400_000
    cols, path = _sorted_q6_files(cb, tmp_path, n, 41, variant)
    plan = t.q6_partial_plan(variant, scan=t.q6_native_scan(variant, [path]))
    state, st = run(cb, plan, chunk_rows=60_000)
    n_rg = (n + 16_383) // 16_384
    assert st["scan_pruned_row_groups"] >= n_rg * 0.8 and st["scan_pruned_rows"] >= n * 0.8
    monkeypatch.setenv("CB200_NO_PRUNE", "1")
    state_all, st_all = run(cb, plan, chunk_rows=60_000)
    assert st_all["scan_pruned_row_groups"] == 0
    assert st["h2d_bytes"] <= 0.2 * st_all["h2d_bytes"]              # <= 10-20 % of the file crosses PCIe
    assert state.equals(state_all)
    res, _ = run(cb, t.q6_final_plan(variant), [state])
    if variant == "dec":
        d = oracle.dec_from_i64
        exp = oracle.q6_dec(d(cols["l_quantity"]), d(cols["l_extendedprice"]), d(cols["l_discount"]), cols["l_shipdate"], t.DATE_1994_01_01, t.DATE_1995_01_01, 5, 7, 2400, 1)
        assert unscaled(res.column(0)[0].as_py()) == exp
    # the filter above the scan is pushed even when the JVM sends no data_filters (the conjuncts are the same expressions)
    plan2 = t.q6_partial_plan(variant, scan=t.q6_native_scan(variant, [path], push_filters=False))
    monkeypatch.delenv("CB200_NO_PRUNE")
    state2, st2 = run(cb, plan2, chunk_rows=60_000)
    assert st2["scan_pruned_row_groups"] == st["scan_pruned_row_groups"] and state2.equals(state)


def test_q1_prunes_nothing_it_should_not(cb, tmp_path):
    """Q1 keeps 98 % of the rows: only the row groups entirely after the cut-off may go, and the result must not move."""
    t = cb.tpch
    n = 200_000
    cols = t.gen_lineitem(n, seed=43)
    order = np.argsort(cols["l_shipdate"], kind="stable")
    cols = {k: v[order] for k, v in cols.items()}
    path = t.write_lineitem_parquet(cols, str(tmp_path / "q1s.parquet"), "dec", row_group_size=8192)
    plan = t.q1_partial_plan("dec", scan=t.q1_native_scan("dec", [path]))
    state, st = run(cb, plan)
    exp_pruned = sum(1 for g in range(0, n, 8192) if cols["l_shipdate"][g] > t.Q1_CUTOFF)
    assert st["scan_pruned_row_groups"] == exp_pruned > 0
    import os
    os.environ["CB200_NO_PRUNE"] = "1"
    try:
        state_all, _ = run(cb, plan)
    finally:
        del os.environ["CB200_NO_PRUNE"]
    key = lambda tb: sorted(tb.to_pylist(), key=lambda r: (r["col_0"], r["col_1"]))
    assert key(state) == key(state_all)


def test_file_splits_own_the_row_groups_that_start_inside_them(cb, tmp_path):
    """SparkPartitionedFile.start / length (operator.proto:103-109): two tasks over the two halves of one file see every row
    group exactly once."""
    import pyarrow.parquet as pq
    t = cb.tpch
    P = cb.proto
    n = 100_000
    cols = t.gen_lineitem(n, seed=45)
    path = t.write_lineitem_parquet(cols, str(tmp_path / "split.parquet"), "dec", row_group_size=10_000)
    import os
    size = os.path.getsize(path)
    md = pq.ParquetFile(path).metadata
    starts = [md.row_group(g).column(0).dictionary_page_offset or md.row_group(g).column(0).data_page_offset for g in range(md.num_row_groups)]
    cut = starts[4] + 1                                              # split point in the middle of a row group
    fields = list(zip(t.Q1_COLUMNS, t.q1_scan_fields("dec"), [True] * 7))
    counts = []
    for lo, ln in ((0, cut), (cut, size - cut)):
        sc = P.native_scan(fields, fields, [(path, lo, ln, size)])
        plan = P.hash_agg(sc, [], [P.agg_count([P.literal(1, P.INT32)]), P.agg_sum(P.bound(0, t.D12), P.DECIMAL(22, 2))], P.PARTIAL)
        res, _ = run(cb, plan)
        counts.append((res.column(0)[0].as_py(), unscaled(res.column(1)[0].as_py())))
    exp_first = sum(md.row_group(g).num_rows for g in range(md.num_row_groups) if starts[g] < cut)
    assert counts[0][0] == exp_first and counts[0][0] + counts[1][0] == n
    assert counts[0][1] + counts[1][1] == int(cols["l_quantity"].sum())


def test_annotations_that_change_the_meaning_of_the_bytes_are_refused(cb, tmp_path):
    """TIMESTAMP_MILLIS / NANOS read as microseconds or UINT32 sign-extended would be silently wrong data: Unsupported (the caller
    keeps its CPU path), never a guess.  Microsecond timestamps decode."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    P = cb.proto
    n = 1000
    ts = np.arange(n, dtype=np.int64) * 1_000_000 + 1_600_000_000_000_000
    for unit, ok in (("us", True), ("ms", False), ("ns", False)):
        path = str(tmp_path / f"ts_{unit}.parquet")
        arr = pa.array(ts if unit == "us" else (ts // 1000 if unit == "ms" else ts * 1000), type=pa.timestamp(unit))
        pq.write_table(pa.table({"t": arr}), path, compression="NONE", use_dictionary=False, coerce_timestamps=None, version="2.6")
        fields = [("t", P.TIMESTAMP, True)]
        plan = P.projection(P.native_scan(fields, fields, [path]), [P.bound(0, P.TIMESTAMP)])
        if ok:
            res, _ = run(cb, plan)
            assert res.column(0).cast(pa.int64()).to_numpy().tolist() == ts.tolist()
        else:
            with pytest.raises(cb.native.Unsupported):
                run(cb, plan)
    path = str(tmp_path / "u32.parquet")
    pq.write_table(pa.table({"u": pa.array([1, 2**31 + 5, 7], type=pa.uint32())}), path, compression="NONE", use_dictionary=False)
    fields = [("u", P.INT64, True)]
    with pytest.raises(cb.native.Unsupported):
        run(cb, P.projection(P.native_scan(fields, fields, [path]), [P.bound(0, P.INT64)]))


def test_scan_blocks_are_reused_across_plans(cb, tmp_path):
    """Ten plans over the same files in a row, alternating batch sizes: results identical every time (the slot blocks come back
    from the process-wide cache; a stale pointer or a missing stream dependency would show up as a wrong sum)."""
    t = cb.tpch
    P = cb.proto
    n = 250_000
    cols = t.gen_lineitem(n, seed=47)
    path = t.write_lineitem_parquet(cols, str(tmp_path / "reuse.parquet"), "dec", row_group_size=8192)
    fields = list(zip(t.Q1_COLUMNS, t.q1_scan_fields("dec"), [True] * 7))
    sc = P.native_scan(fields, fields, [path])
    plan = P.hash_agg(sc, [], [P.agg_sum(P.bound(1, t.D12), P.DECIMAL(22, 2)), P.agg_sum(P.bound(3, t.D12), P.DECIMAL(22, 2)), P.agg_count([P.literal(1, P.INT32)])], P.PARTIAL)
    exp = (int(cols["l_extendedprice"].sum()), int(cols["l_tax"].sum()), n)
    for i in range(10):
        res, _ = run(cb, plan, chunk_rows=[20_000, 64_000, 250_000][i % 3])
        r = res.to_pylist()[0]
        assert (unscaled(r["col_0"]), unscaled(r["col_2"]), r["col_4"]) == exp, i
