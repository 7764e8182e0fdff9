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
