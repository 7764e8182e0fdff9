"""GPU parity: SUM(int) in the three Spark eval modes (sum_int.rs: Legacy wraps, TRY -> NULL, ANSI -> error), dense and hash
grouping, Partial -> Final state flow (TRY carries the two-column state (sum, has_all_nulls))."""
import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

LEGACY, TRY, ANSI = 0, 1, 2


@pytest.fixture(scope="module")
def cb():
    import comet_b200
    return comet_b200


def run(cb, plan, inputs, chunk_rows=None):
    cfg = {"spark.comet.b200.chunkRows": str(chunk_rows)} if chunk_rows else None
    with cb.native.Plan(plan, inputs, config=cfg) as p:
        return p.collect()


def plans(cb, key_dt, val_dt, mode):
    P = cb.proto
    aggs_p = [P.agg_sum(P.bound(1, val_dt), P.INT64, mode), P.agg_count([P.bound(1, val_dt)])]
    partial = P.hash_agg(P.scan([key_dt, val_dt]), [P.bound(0, key_dt)], aggs_p, P.PARTIAL)
    state = [key_dt, P.INT64] + ([P.BOOL] if mode == TRY else []) + [P.INT64]
    aggs_f = [P.agg_sum(P.unbound("s", val_dt), P.INT64, mode), P.agg_count([P.unbound("c", val_dt)])]
    final = P.hash_agg(P.scan(state, source="shuffle"), [P.bound(0, key_dt)], aggs_f, P.FINAL)
    return partial, final


def oracle_sums(oracle, keys, vals, valid, mode):
    uk, inv = np.unique(keys, return_inverse=True)
    acc = oracle.SumIntGroups(len(uk), mode)
    acc.update(vals, valid, inv)
    return uk, acc


@pytest.mark.parametrize("mode", [LEGACY, TRY, ANSI])
@pytest.mark.parametrize("dense", [True, False])
def test_sum_int_modes_no_overflow(cb, oracle, mode, dense):
    P = cb.proto
    n = 60_000
    rng = np.random.default_rng(mode * 2 + dense)
    vals = rng.integers(-2**40, 2**40, n)
    valid = rng.random(n) > 0.2
    if dense:
        codes = rng.integers(0, 5, n)
        valid[codes == 4] = False                       # a group whose inputs are all NULL
        names = ["a", "b", "c", "d", "e"]
        keys_arr = pa.DictionaryArray.from_arrays(pa.array(codes.astype(np.int8)), pa.array(names))
        key_dt, keys = P.STRING, codes
    else:
        keys = rng.integers(0, 7000, n) * 1_000_003
        valid[keys == keys[0]] = False
        keys_arr, key_dt = pa.array(keys), P.INT64
    tbl = pa.table({"k": keys_arr, "v": pa.array(vals, mask=~valid)})
    partial, final = plans(cb, key_dt, P.INT64, mode)
    state = run(cb, partial, [tbl.to_batches(max_chunksize=8192)], 25_000)
    assert state.num_columns == (4 if mode == TRY else 3)
    res = run(cb, final, [state])
    uk, acc = oracle_sums(oracle, keys, vals, valid, mode)
    got = {r["col_0"]: (r["col_1"], r["col_2"]) for r in res.to_pylist()}
    assert len(got) == len(uk)
    for i, k in enumerate(uk.tolist()):
        kk = names[k] if dense else k
        exp = int(acc.sums[i]) if acc.sums_valid[i] else None
        assert got[kk] == (exp, int((valid & (keys == k)).sum())), kk


def test_try_overflow_is_null_and_sticky_through_final(cb):
    P = cb.proto
    big = 2**62
    keys = pa.array([1, 1, 1, 2, 2, 3, 3], type=pa.int64())
    vals = pa.array([big, big, big, 5, 6, None, None], type=pa.int64())     # group 1 overflows in every order; group 3 is all NULL
    partial, final = plans(cb, P.INT64, P.INT64, TRY)
    state = run(cb, partial, [pa.table({"k": keys, "v": vals})])
    st = {r["col_0"]: (r["col_1"], r["col_2"]) for r in state.to_pylist()}
    assert st == {1: (None, False), 2: (11, False), 3: (0, True)}            # (sum, has_all_nulls) exactly as sum_int.rs:322-329
    more = run(cb, partial, [pa.table({"k": pa.array([1, 3], type=pa.int64()), "v": pa.array([1, 9], type=pa.int64())})])
    res = run(cb, final, [pa.concat_tables([state, more])])
    got = {r["col_0"]: r["col_1"] for r in res.to_pylist()}
    assert got == {1: None, 2: 11, 3: 9}                                     # overflow stays NULL after merging a healthy partial


def test_ansi_overflow_raises(cb):
    P = cb.proto
    partial, final = plans(cb, P.INT64, P.INT64, ANSI)
    tbl = pa.table({"k": pa.array([1, 1, 2], type=pa.int64()), "v": pa.array([2**63 - 1, 1, 4], type=pa.int64())})
    with pytest.raises(cb.native.CometB200Error) as ei:
        run(cb, partial, [tbl])
    assert ei.value.error_class == "ARITHMETIC_OVERFLOW"
    # two partials that each fit but whose merge overflows: the error comes from the Final stage (merge_batch -> update_batch, sum_int.rs:236-243)
    a = run(cb, partial, [pa.table({"k": pa.array([7], type=pa.int64()), "v": pa.array([2**63 - 1], type=pa.int64())})])
    b = run(cb, partial, [pa.table({"k": pa.array([7], type=pa.int64()), "v": pa.array([2**63 - 1], type=pa.int64())})])
    with pytest.raises(cb.native.CometB200Error):
        run(cb, final, [pa.concat_tables([a, b])])


def test_order_dependent_overflow_is_refused_not_guessed(cb):
    """max, max, -max, -max: row order decides whether add_checked overflows; the engine reports that instead of picking an order."""
    P = cb.proto
    m = 2**63 - 1
    partial, _ = plans(cb, P.INT64, P.INT64, TRY)
    tbl = pa.table({"k": pa.array([1, 1, 1, 1], type=pa.int64()), "v": pa.array([m, -m, m, -m], type=pa.int64())})
    with pytest.raises(cb.native.CometB200Error) as ei:
        run(cb, partial, [tbl])
    assert "order" in str(ei.value).lower()


def test_small_int_inputs_widen(cb, oracle):
    P = cb.proto
    n = 30_000
    rng = np.random.default_rng(4)
    vals = rng.integers(-2**31, 2**31, n).astype(np.int32)
    keys = rng.integers(0, 3, n)
    tbl = pa.table({"k": pa.DictionaryArray.from_arrays(pa.array(keys.astype(np.int8)), pa.array(["x", "y", "z"])), "v": pa.array(vals)})
    partial, final = plans(cb, P.STRING, P.INT32, ANSI)
    res = run(cb, final, [run(cb, partial, [tbl.to_batches(max_chunksize=4096)])])
    got = {r["col_0"]: r["col_1"] for r in res.to_pylist()}
    for i, name in enumerate(["x", "y", "z"]):
        assert got[name] == int(vals[keys == i].astype(np.int64).sum())


def test_count_star_alone(cb):
    """COUNT(*) reads no column; with and without a filter, ungrouped and through Partial -> Final."""
    P = cb.proto
    n = 70_001
    rng = np.random.default_rng(1)
    d = rng.integers(8000, 10000, n).astype(np.int32)
    tbl = pa.table({"s": pa.array(["x"] * n), "k": pa.array(rng.integers(0, 9, n)), "d": pa.array(d, type=pa.date32())})
    sc = P.scan([P.STRING, P.INT64, P.DATE])
    partial = P.hash_agg(sc, [], [P.agg_count([P.literal(1, P.INT32)])], P.PARTIAL)
    final = P.hash_agg(P.scan([P.INT64], source="shuffle"), [], [P.agg_count([P.unbound("c", P.INT32)])], P.FINAL)
    st = run(cb, partial, [tbl.to_batches(max_chunksize=8192)], 30_000)
    assert run(cb, final, [st]).column(0).to_pylist() == [n]
    filtered = P.hash_agg(P.filter_(sc, P.lt(P.bound(2, P.DATE), P.literal(9000, P.DATE))), [], [P.agg_count([P.literal(1, P.INT32)])], P.PARTIAL)
    assert run(cb, filtered, [tbl.to_batches(max_chunksize=8192)]).column(0).to_pylist() == [int((d < 9000).sum())]
