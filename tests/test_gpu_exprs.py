"""GPU parity for the expression layer: random columns with NULLs through fused
filter + projection plans vs the oracle-backed interpreter in tests/exprs.py."""
import decimal

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import comet_b200
    return comet_b200


@pytest.fixture(scope="module")
def E():
    import exprs
    return exprs


N = 20_000


def table(seed=0):
    """columns: 0 i32, 1 i64, 2 f64, 3 d(12,2), 4 d(12,2), 5 d(26,4), 6 d(38,10), 7 i8, 8 date -- all with NULLs"""
    from comet_b200 import proto as P
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    n = N
    valid = [rng.random(n) > 0.12 for _ in range(9)]
    i32 = rng.integers(-2**31, 2**31, n)
    i32[:50] = 2**31 - 1
    i64 = rng.integers(-2**62, 2**62, n) * rng.integers(0, 3, n)
    i64[:30] = 2**63 - 1
    f64 = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 6, n)
    f64[::101] = np.nan
    f64[::103] = -0.0
    f64[::107] = np.inf
    d12a = rng.integers(-10**11, 10**11, n)
    d12b = rng.integers(0, 2000, n)
    d26 = [int(a) * int(b) for a, b in zip(rng.integers(-10**12, 10**12, n), rng.integers(0, 10**12, n))]
    d38 = [int(a) * 10**20 + int(b) for a, b in zip(rng.integers(-10**17, 10**17, n), rng.integers(0, 10**18, n))]
    i8 = rng.integers(-128, 128, n)
    date = rng.integers(8000, 11000, n)
    dts = [P.INT32, P.INT64, P.DOUBLE, P.DECIMAL(12, 2), P.DECIMAL(12, 2), P.DECIMAL(26, 4), P.DECIMAL(38, 10), P.INT8, P.DATE]

    ctx = decimal.Context(prec=60)   # the default context (28 digits) would round the 37-digit values

    def dec_arr(vals, p, s, v):
        return pa.array([decimal.Decimal(int(x)).scaleb(-s, context=ctx) if ok else None for x, ok in zip(vals, v)], type=pa.decimal128(p, s))
    arrays = [pa.array(i32.astype(np.int32), mask=~valid[0]), pa.array(i64, mask=~valid[1]), pa.array(f64, mask=~valid[2]), dec_arr(d12a, 12, 2, valid[3]),
              dec_arr(d12b, 12, 2, valid[4]), dec_arr(d26, 26, 4, valid[5]), dec_arr(d38, 38, 10, valid[6]), pa.array(i8.astype(np.int8), mask=~valid[7]),
              pa.array(date.astype(np.int32), type=pa.date32(), mask=~valid[8])]
    tbl = pa.table(arrays, names=[f"c{i}" for i in range(9)])
    cols = [(i32.astype(np.int64), valid[0]), (i64.astype(np.int64), valid[1]), (f64, valid[2]), (O.dec_from_i64(d12a), valid[3]), (O.dec_from_i64(d12b), valid[4]),
            (O.dec_from_ints(d26), valid[5]), (O.dec_from_ints(d38), valid[6]), (i8.astype(np.int64), valid[7]), (date.astype(np.int64), valid[8])]
    return tbl, cols, dts


def gpu_eval(cb, dts, tbl, pred, outs, chunk=7000):
    P = cb.proto
    node = P.scan(dts)
    if pred is not None:
        node = P.filter_(node, pred.proto())
    node = P.projection(node, [o.proto() for o in outs])
    with cb.native.Plan(node, [tbl.to_batches(max_chunksize=4096)], config={"spark.comet.b200.chunkRows": str(chunk)}) as p:
        return p.collect()


def check(cb, E, pred, outs, seed=0):
    from oracle import oracle as O
    tbl, cols, dts = table(seed)
    res = gpu_eval(cb, dts, tbl, pred, outs)
    keep = np.ones(N, dtype=bool)
    if pred is not None:
        pv, pvalid = pred.eval(cols)
        keep = pv & pvalid                                     # FilterExec drops FALSE and NULL
    n_keep = int(keep.sum())
    assert (res.num_rows if res is not None else 0) == n_keep
    if n_keep == 0:
        return
    kept_cols = [(np.asarray(v)[keep], np.asarray(ok)[keep]) for v, ok in cols]   # ProjectionExec sees FilterExec's output only
    for j, o in enumerate(outs):
        v, valid = o.eval(kept_cols)
        valid = np.asarray(valid)
        col = res.column(j).combine_chunks()
        if pa.types.is_date32(col.type):
            col = col.cast(pa.int32())
        got_valid = np.array([x is not None for x in col.to_pylist()]) if col.null_count else np.ones(n_keep, dtype=bool)
        assert (got_valid == valid).all(), f"output {j}: validity differs"
        if o.dt.name == "DECIMAL":
            got = np.frombuffer(col.buffers()[1], dtype=np.uint64)[:2 * n_keep].reshape(-1, 2)
            bad = np.nonzero((got[valid] != v[valid]).any(axis=1))[0]
            assert bad.size == 0, f"output {j}: {bad.size} rows differ, first got {O.dec_to_ints(got[valid][bad[:1]])} want {O.dec_to_ints(v[valid][bad[:1]])}"
        elif o.dt.name in ("DOUBLE", "FLOAT"):
            got = np.frombuffer(col.buffers()[1], dtype=np.float64 if o.dt.name == "DOUBLE" else np.float32)[:n_keep]
            w = np.uint64 if o.dt.name == "DOUBLE" else np.uint32
            gb, eb = np.ascontiguousarray(got[valid]).view(w), np.ascontiguousarray(v[valid]).view(w)
            bad = np.nonzero(gb != eb)[0]
            assert bad.size == 0, f"output {j}: {bad.size} rows differ, first got {gb[bad[0]]:#x} want {eb[bad[0]]:#x}"  # bit-exact incl. NaN, -0.0
        elif o.dt.name == "BOOL":
            got = np.array(col.to_pylist(), dtype=object)
            assert (got[valid].astype(bool) == v[valid]).all(), f"output {j}"
        else:
            got = np.array(col.to_pylist(), dtype=object)
            assert (got[valid].astype(np.int64) == v[valid]).all(), f"output {j}"


def C(E, i):
    from comet_b200 import proto as P
    dts = [P.INT32, P.INT64, P.DOUBLE, P.DECIMAL(12, 2), P.DECIMAL(12, 2), P.DECIMAL(26, 4), P.DECIMAL(38, 10), P.INT8, P.DATE]
    return E.Col(i, dts[i])


def test_integer_arithmetic_modes(cb, E):
    P = cb.proto
    a, b, c8 = C(E, 0), C(E, 1), C(E, 7)
    check(cb, E, None, [E.Arith("add", a, a, P.INT32), E.Arith("multiply", a, a, P.INT32), E.Arith("subtract", b, b, P.INT64),
                       E.Arith("multiply", b, b, P.INT64), E.Arith("add", c8, c8, P.INT8), E.Arith("multiply", c8, c8, P.INT8)])
    check(cb, E, None, [E.Arith("add", a, a, P.INT32, E.TRY), E.Arith("multiply", b, b, P.INT64, E.TRY), E.Arith("add", c8, c8, P.INT8, E.TRY),
                       E.Neg(a), E.Neg(b)])


def test_integer_ansi_overflow_raises(cb, E):
    P = cb.proto
    tbl, cols, dts = table(0)
    with pytest.raises(cb.native.CometB200Error) as ei:
        gpu_eval(cb, dts, tbl, None, [E.Arith("add", C(E, 0), C(E, 0), P.INT32, E.ANSI)])
    assert ei.value.error_class == "ARITHMETIC_OVERFLOW"
    with pytest.raises(E.AnsiError):
        E.Arith("add", C(E, 0), C(E, 0), P.INT32, E.ANSI).eval(cols)


def test_float_arithmetic_and_total_order(cb, E):
    P = cb.proto
    f = C(E, 2)
    lit = E.Lit(0.0, P.DOUBLE)
    check(cb, E, E.Cmp("gt_eq", f, lit), [E.Arith("add", f, f, P.DOUBLE), E.Arith("multiply", f, E.Arith("subtract", f, E.Lit(1.5, P.DOUBLE), P.DOUBLE), P.DOUBLE),
                                         E.Cmp("eq", f, f), E.Cmp("lt", f, E.Lit(-0.0, P.DOUBLE)), E.Neg(f)])
    check(cb, E, E.Cmp("lt", f, E.Lit(float("nan"), P.DOUBLE)), [f])   # totalOrder: everything but +NaN is < +NaN


def test_decimal_plain_and_wide(cb, E):
    P = cb.proto
    a, b, w, x = C(E, 3), C(E, 4), C(E, 5), C(E, 6)
    outs = [
        E.CheckOverflow(E.Arith("multiply", a, b, P.DECIMAL(25, 4)), P.DECIMAL(25, 4)),                   # plain mul
        E.CheckOverflow(E.Arith("subtract", E.Lit(1, P.DECIMAL(1, 0)), b, P.DECIMAL(13, 2)), P.DECIMAL(13, 2)),
        E.CheckOverflow(E.Arith("multiply", w, a, P.DECIMAL(38, 6)), P.DECIMAL(38, 6)),                   # wide mul, natural scale == out
        E.CheckOverflow(E.Arith("multiply", x, a, P.DECIMAL(38, 6)), P.DECIMAL(38, 6)),                   # wide mul with HALF_UP rescale + overflow -> NULL
        E.CheckOverflow(E.Arith("add", x, w, P.DECIMAL(38, 9)), P.DECIMAL(38, 9)),                        # wide add, scale reduction
        E.CheckOverflow(E.Arith("subtract", x, x, P.DECIMAL(38, 10)), P.DECIMAL(38, 10)),
        E.CheckOverflow(E.Arith("add", a, b, P.DECIMAL(13, 2)), P.DECIMAL(5, 2)),                         # bound check that really fires -> NULL
        E.Neg(a),
    ]
    check(cb, E, None, outs)
    check(cb, E, E.Logic("and", E.Cmp("gt", a, E.Lit(0, P.DECIMAL(12, 2))), E.Cmp("lt_eq", b, E.Lit(1000, P.DECIMAL(12, 2)))), outs[:3], seed=3)


def test_decimal_ansi_errors(cb, E):
    P = cb.proto
    tbl, cols, dts = table(0)
    bad = E.CheckOverflow(E.Arith("add", C(E, 3), C(E, 4), P.DECIMAL(13, 2)), P.DECIMAL(5, 2), fail=True)
    with pytest.raises(cb.native.CometB200Error):
        gpu_eval(cb, dts, tbl, None, [bad])
    with pytest.raises(E.AnsiError):
        bad.eval(cols)
    wide_bad = E.Arith("multiply", C(E, 6), C(E, 6), P.DECIMAL(38, 6), E.ANSI)
    with pytest.raises(cb.native.CometB200Error):
        gpu_eval(cb, dts, tbl, None, [wide_bad])


def test_casts_and_rescale(cb, E):
    P = cb.proto
    a, x, i = C(E, 3), C(E, 6), C(E, 0)
    outs = [E.Cast(a, P.DECIMAL(14, 4)), E.Cast(x, P.DECIMAL(38, 2)), E.Cast(x, P.DECIMAL(20, 0)), E.Cast(i, P.DECIMAL(12, 2)), E.Cast(i, P.DECIMAL(8, 2)),
            E.Cast(i, P.INT64), E.Cast(i, P.DOUBLE), E.CheckOverflow(E.Cast(x, P.DECIMAL(30, 4)), P.DECIMAL(30, 4)), E.Cast(C(E, 7), P.INT32)]
    check(cb, E, None, outs)


def test_kleene_logic_null_semantics_if_in(cb, E):
    P = cb.proto
    a, i, d = C(E, 3), C(E, 0), C(E, 8)
    p1 = E.Cmp("gt", a, E.Lit(0, P.DECIMAL(12, 2)))
    p2 = E.Cmp("lt", i, E.Lit(0, P.INT32))
    p3 = E.Cmp("gt_eq", d, E.Lit(9000, P.DATE))
    outs = [E.Logic("and", p1, p2), E.Logic("or", p1, p2), E.Not(E.Logic("or", p1, E.Logic("and", p2, p3))), E.IsNull(a), E.IsNull(i, negate=True),
            E.If(p1, i, E.Arith("add", i, E.Lit(1, P.INT32), P.INT32)), E.If(p2, a, E.Neg(a)),
            E.In(C(E, 7), [E.Lit(1, P.INT8), E.Lit(5, P.INT8), E.Lit(-7, P.INT8)]), E.In(C(E, 7), [E.Lit(1, P.INT8), E.Lit(None, P.INT8)], negated=True),
            E.In(C(E, 4), [E.Lit(100, P.DECIMAL(12, 2)), E.Lit(1999, P.DECIMAL(12, 2))]),
            E.CaseWhen([p1, p2], [i, E.Arith("multiply", i, E.Lit(2, P.INT32), P.INT32)], E.Lit(7, P.INT32)), E.CaseWhen([p3, p1], [a, E.Neg(a)])]
    check(cb, E, None, outs)
    check(cb, E, E.Logic("or", E.Logic("and", p1, p2), E.IsNull(d)), outs[:4], seed=5)     # NULL predicate rows are dropped
    check(cb, E, E.Logic("and", E.IsNull(a, negate=True), E.Not(p1)), [a, d], seed=6)


def test_filter_selectivity_extremes_and_chunk_boundaries(cb, E):
    P = cb.proto
    d = C(E, 8)
    check(cb, E, E.Cmp("lt", d, E.Lit(0, P.DATE)), [d])            # nothing passes
    check(cb, E, E.Cmp("gt", d, E.Lit(0, P.DATE)), [d, C(E, 1)])   # everything non-null passes
    check(cb, E, E.Cmp("eq", d, E.Lit(9500, P.DATE)), [C(E, 1)])   # a handful of rows


def test_ansi_error_only_on_rows_that_take_the_branch(cb, E):
    """CaseExpr evaluates THEN under the selection: an ANSI overflow in a row that does not take the branch is not an error."""
    P = cb.proto
    i = C(E, 0)
    safe = E.Cmp("lt", i, E.Lit(1 << 30, P.INT32))
    guarded = E.If(E.Logic("and", safe, E.Cmp("gt", i, E.Lit(-(1 << 30), P.INT32))), E.Arith("add", i, i, P.INT32, E.ANSI), E.Lit(0, P.INT32))
    tbl, cols, dts = table(0)
    res = gpu_eval(cb, dts, tbl, None, [guarded])
    assert res.num_rows == N
    with pytest.raises(cb.native.CometB200Error):
        gpu_eval(cb, dts, tbl, None, [E.If(E.Not(safe), E.Arith("add", i, i, P.INT32, E.ANSI), E.Lit(0, P.INT32))])


# ---- division (round 2): decimal_div, checked float / integer division --------------------------------------------------------
def test_decimal_division(cb, E):
    """decimal_div (spark-expr/src/math_funcs/div.rs:75-190): HALF_UP on one extra digit, narrow and BigInt-sized operands.  The
    JVM wraps the divisor in nullIf(= 0) outside ANSI mode; the tests do the same with IF."""
    P = cb.proto
    a, b, w, x = C(E, 3), C(E, 4), C(E, 5), C(E, 6)
    nz = lambda e, dt: E.If(E.Cmp("eq", e, E.Lit(0, dt)), E.Lit(None, dt), e)
    outs = [
        E.CheckOverflow(E.Arith("divide", a, nz(b, P.DECIMAL(12, 2)), P.DECIMAL(27, 15)), P.DECIMAL(27, 15)),   # d(12,2) / d(12,2): Spark's result type
        E.CheckOverflow(E.Arith("divide", w, nz(a, P.DECIMAL(12, 2)), P.DECIMAL(38, 15)), P.DECIMAL(38, 15)),   # d(26,4) / d(12,2)
        E.CheckOverflow(E.Arith("divide", x, nz(w, P.DECIMAL(26, 4)), P.DECIMAL(38, 6)), P.DECIMAL(38, 6)),     # wide: scaled numerator beyond 38 digits
        E.CheckOverflow(E.Arith("divide", a, nz(x, P.DECIMAL(38, 10)), P.DECIMAL(38, 20)), P.DECIMAL(38, 20)),  # l_exp = 29: BigInt path of the reference
        E.Arith("divide", x, nz(a, P.DECIMAL(12, 2)), P.DECIMAL(38, 18)),                                       # quotients past i128 -> i128::MAX sentinel
    ]
    check(cb, E, None, outs[:4])
    check(cb, E, E.Cmp("gt", b, E.Lit(0, P.DECIMAL(12, 2))), outs[:2], seed=4)
    # the sentinel row values are what the reference stores before CheckOverflow turns them into NULL: compare through CheckOverflow
    check(cb, E, None, [E.CheckOverflow(outs[4], P.DECIMAL(38, 18))], seed=2)


def test_decimal_division_by_zero_in_ansi_mode_raises(cb, E):
    P = cb.proto
    tbl, cols, dts = table(0)
    e = E.Arith("divide", C(E, 3), C(E, 4), P.DECIMAL(27, 15), E.ANSI)     # column 4 holds zeros
    with pytest.raises(cb.native.CometB200Error) as ei:
        gpu_eval(cb, dts, tbl, None, [e])
    assert ei.value.error_class == "DIVIDE_BY_ZERO"
    with pytest.raises(E.AnsiError):
        e.eval(cols)
    # rows that do not reach the division (filtered / NULL divisor) do not raise
    check(cb, E, E.Cmp("gt", C(E, 4), E.Lit(0, P.DECIMAL(12, 2))), [e], seed=1)


def test_float_division_modes(cb, E):
    """Legacy: IEEE (x / 0 = +-Inf / NaN); TRY: NULL for a zero divisor; ANSI: DIVIDE_BY_ZERO (checked_arithmetic.rs:53-128, routed
    there by planner.rs:1094-1125)."""
    P = cb.proto
    f = C(E, 2)
    g = E.Arith("subtract", f, f, P.DOUBLE)                                # zeros, NaNs
    check(cb, E, None, [E.Arith("divide", f, g, P.DOUBLE), E.Arith("divide", f, E.Lit(3.0, P.DOUBLE), P.DOUBLE), E.Arith("divide", f, g, P.DOUBLE, E.TRY),
                       E.Arith("divide", g, f, P.DOUBLE, E.TRY)])
    tbl, cols, dts = table(0)
    with pytest.raises(cb.native.CometB200Error) as ei:
        gpu_eval(cb, dts, tbl, None, [E.Arith("divide", f, g, P.DOUBLE, E.ANSI)])
    assert ei.value.error_class == "DIVIDE_BY_ZERO"
    check(cb, E, None, [E.Arith("divide", g, E.Lit(2.0, P.DOUBLE), P.DOUBLE, E.ANSI)])


def test_integer_division_modes(cb, E):
    P = cb.proto
    a, b, c8 = C(E, 0), C(E, 1), C(E, 7)
    nz32 = E.If(E.Cmp("eq", a, E.Lit(0, P.INT32)), E.Lit(7, P.INT32), a)
    check(cb, E, None, [E.Arith("divide", a, E.Lit(7, P.INT32), P.INT32), E.Arith("divide", b, E.Lit(-3, P.INT64), P.INT64), E.Arith("divide", a, nz32, P.INT32),
                       E.Arith("divide", c8, c8, P.INT8, E.TRY), E.Arith("divide", b, E.Lit(-1, P.INT64), P.INT64, E.TRY), E.Arith("divide", a, a, P.INT32, E.TRY)])
    tbl, cols, dts = table(0)
    with pytest.raises(cb.native.CometB200Error) as ei:
        gpu_eval(cb, dts, tbl, None, [E.Arith("divide", c8, c8, P.INT8, E.ANSI)])          # column 7 holds zeros
    assert ei.value.error_class == "DIVIDE_BY_ZERO"
    with pytest.raises(cb.native.CometB200Error):
        gpu_eval(cb, dts, tbl, None, [E.Arith("divide", c8, c8, P.INT8)])                  # Legacy: arrow-arith fails the query
