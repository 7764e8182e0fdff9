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
                        cols["l_returnflag"], cols["l_linestatus"], 3, 2, t.DATE_1998_09_02, 1)
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
