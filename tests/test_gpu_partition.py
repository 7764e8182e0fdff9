"""GPU parity: hash repartitioning (murmur3 seed 42 -> pmod -> stable counting sort) vs the oracle, which is
pinned to the reference's murmur3/pmod known-answer tests."""
import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import comet_b200
    return comet_b200


def make_table(n, seed, nulls=True):
    rng = np.random.default_rng(seed)
    i32 = rng.integers(-2**31, 2**31, n).astype(np.int32)
    i64 = rng.integers(-2**63, 2**63, n).astype(np.int64)
    f64 = rng.standard_normal(n)
    f64[::17] = -0.0
    f64[::19] = 0.0
    small = rng.integers(-10**11, 10**11, n).astype(np.int64)
    large_hi = rng.integers(-10**6, 10**6, n).astype(np.int64)
    names = [f"k{i}" for i in range(37)] + ["", "😁", "天地"]
    codes = rng.integers(0, len(names), n).astype(np.int32)
    date = rng.integers(0, 20000, n).astype(np.int32)
    mask = (rng.random(n) < 0.1) if nulls else np.zeros(n, dtype=bool)
    import decimal
    large_py = [decimal.Decimal(int(h) * 10**19 + 12345).scaleb(-2) for h in large_hi]
    tbl = pa.table({
        "i32": pa.array(i32, mask=mask), "i64": pa.array(i64), "f64": pa.array(f64),
        "small": pa.array([decimal.Decimal(int(v)).scaleb(-2) for v in small], type=pa.decimal128(12, 2)),
        "large": pa.array(large_py, type=pa.decimal128(28, 2)),
        "s": pa.DictionaryArray.from_arrays(pa.array(codes), pa.array(names)),
        "date": pa.array(date, type=pa.date32()),
        "row": pa.array(np.arange(n, dtype=np.int64)),
    })
    raw = dict(i32=i32, i32_valid=~mask, i64=i64, f64=f64, small=small, large=[int(h) * 10**19 + 12345 for h in large_hi], codes=codes, names=names, date=date)
    return tbl, raw


def oracle_hashes(o, raw, keys):
    n = raw["i64"].shape[0]
    h = np.full(n, 42, dtype=np.uint32)
    for k in keys:
        if k == "i32":
            h = o.murmur3_column("i32", raw["i32"], valid=raw["i32_valid"], hashes=h)
        elif k == "i64":
            h = o.murmur3_column("i64", raw["i64"], hashes=h)
        elif k == "f64":
            h = o.murmur3_column("f64", raw["f64"], hashes=h)
        elif k == "small":
            h = o.murmur3_column("dec_small", o.dec_from_i64(raw["small"]), hashes=h)
        elif k == "large":
            h = o.murmur3_column("dec_large", o.dec_from_ints(raw["large"]), hashes=h)
        elif k == "s":
            h = o.murmur3_strings([raw["names"][c] for c in raw["codes"]], hashes=h)
        elif k == "date":
            h = o.murmur3_column("date32", raw["date"], hashes=h)
    return h


SCHEMA = ["i32", "i64", "f64", "small", "large", "s", "date", "row"]


def plan_for(cb, keys, n_parts):
    P = cb.proto
    types = [P.INT32, P.INT64, P.DOUBLE, P.DECIMAL(12, 2), P.DECIMAL(28, 2), P.STRING, P.DATE, P.INT64]
    sc = P.scan(types)
    exprs = [P.bound(SCHEMA.index(k), types[SCHEMA.index(k)]) for k in keys]
    return P.shuffle_writer(sc, P.hash_partitioning(exprs, n_parts))


@pytest.mark.parametrize("keys,n_parts", [(["i64"], 8), (["i32"], 200), (["f64"], 8), (["small"], 8), (["large"], 16), (["s"], 8), (["date"], 7),
                                          (["s", "i64"], 8), (["i32", "small", "date", "f64"], 200)])
def test_partition_matches_oracle(cb, oracle, keys, n_parts):
    n = 50_000
    tbl, raw = make_table(n, seed=hash(tuple(keys)) % 1000)
    with cb.native.Plan(plan_for(cb, keys, n_parts), [tbl.to_batches(max_chunksize=8192)]) as p:
        out = p.execute()
        starts = p.partition_starts()
        assert p.execute() is None
    h = oracle_hashes(oracle, raw, keys)
    pids, estarts, eidx = oracle.partition_rows(h, n_parts)
    assert starts == list(estarts)
    assert out.num_rows == n
    got_rows = out.column(7).to_numpy()
    assert (got_rows == eidx).all()                       # same partitions, stable order inside each
    # every column travelled with its row
    assert (out.column(1).to_numpy() == raw["i64"][eidx]).all()
    exp_i32 = pa.array(raw["i32"], mask=~raw["i32_valid"]).take(pa.array(eidx))
    assert out.column(0).equals(exp_i32)
    assert out.column(5).to_pylist()[:200] == [raw["names"][c] for c in raw["codes"][eidx][:200]]


def test_partition_of_partial_aggregate_state(cb, oracle):
    """ShuffleWriter(HashAggregate(Partial)) -- the map side of a grouped aggregation."""
    P = cb.proto
    t = cb.tpch
    n = 120_000
    cols = t.gen_lineitem(n, seed=21)
    tbl = pa.table({"k": pa.array(cols["l_orderkey"]), "v": t._dec_array(cols["l_extendedprice"])})
    agg = P.hash_agg(P.scan([P.INT64, P.DECIMAL(12, 2)]), [P.bound(0, P.INT64)], [P.agg_sum(P.bound(1, P.DECIMAL(12, 2)), P.DECIMAL(22, 2))], P.PARTIAL)
    plan = P.shuffle_writer(agg, P.hash_partitioning([P.bound(0, P.INT64)], 8))
    with cb.native.Plan(plan, [tbl.to_batches(max_chunksize=8192)]) as p:
        out = p.execute()
        starts = p.partition_starts()
    keys = out.column(0).to_numpy()
    h = oracle.murmur3_column("i64", keys)
    pids = np.array([oracle.pmod(int(x), 8) for x in h[:5000]])
    bounds = np.searchsorted(np.array(starts), np.arange(5000), side="right") - 1
    assert (pids == bounds).all()
    assert starts[-1] == len(np.unique(cols["l_orderkey"]))
