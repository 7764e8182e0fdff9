"""GPU parity, round 2: the accumulator edge cases the reference pins with its own in-file tests, driven through the C ABI and
compared with the oracle restatement of the same Rust code.

* SUM / AVG(decimal) pushed past the result precision: NULL in Legacy / TRY, error in ANSI, sticky through Partial -> Final
  (spark-expr/src/agg_funcs/sum_decimal.rs:418-439 update_single, :540-607 merge_batch; avg_decimal.rs:483-495, :542-595)
* AggExpr.filter (FILTER (WHERE ...)) with FALSE and NULL filter rows (sum_decimal.rs:442-475 update_batch, KATs :732-801)
* the exact 128-bit escape of the dense accumulators (values >= 2^46), the TIGHT -> TYPE -> SAFE re-run of the
  range-specialised kernels, dense MIN / MAX, under-aligned Decimal128 input buffers (aligned_stream_reader.rs:25-32) and two
  plan handles driven from two threads (jni_api.rs:194-223).
"""
import decimal
import threading

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

LEGACY, TRY, ANSI = 0, 1, 2
decimal.getcontext().prec = 60


@pytest.fixture(scope="module")
def cb():
    import comet_b200
    return comet_b200


def run(cb, plan, inputs, chunk_rows=None, batch_size=8192):
    cfg = {"spark.comet.b200.chunkRows": str(chunk_rows)} if chunk_rows else None
    with cb.native.Plan(plan, inputs, config=cfg, batch_size=batch_size) as p:
        return p.collect()


def dec_arr(vals, p, s, mask=None):
    """python unscaled ints (None = NULL) -> pyarrow decimal128(p, s)"""
    out = [None if v is None else decimal.Decimal(int(v)).scaleb(-s) for v in vals]
    if mask is not None:
        out = [None if m else v for v, m in zip(out, mask)]
    return pa.array(out, type=pa.decimal128(p, s))


def unscaled(x, s):
    return None if x is None else int(x.scaleb(s))


def keys_arr(codes, names):
    return pa.DictionaryArray.from_arrays(pa.array(np.asarray(codes, dtype=np.int8)), pa.array(names))


def sum_plans(P, key_dt, in_dt, sum_dt, mode, with_filter=False):
    cols = [key_dt, in_dt] + ([P.BOOL] if with_filter else [])
    flt = P.bound(2, P.BOOL) if with_filter else None
    partial = P.hash_agg(P.scan(cols), [P.bound(0, key_dt)], [P.agg_sum(P.bound(1, in_dt), sum_dt, mode, filter_expr=flt)], P.PARTIAL)
    final = P.hash_agg(P.scan([key_dt, sum_dt, P.BOOL], source="shuffle"), [P.bound(0, key_dt)], [P.agg_sum(P.unbound("s", in_dt), sum_dt, mode)], P.FINAL)
    return partial, final


def avg_plans(P, key_dt, in_dt, sum_dt, res_dt, mode=LEGACY, with_filter=False):
    cols = [key_dt, in_dt] + ([P.BOOL] if with_filter else [])
    flt = P.bound(2, P.BOOL) if with_filter else None
    partial = P.hash_agg(P.scan(cols), [P.bound(0, key_dt)], [P.agg_avg(P.bound(1, in_dt), res_dt, sum_dt, mode, filter_expr=flt)], P.PARTIAL)
    final = P.hash_agg(P.scan([key_dt, sum_dt, P.INT64], source="shuffle"), [P.bound(0, key_dt)], [P.agg_avg(P.unbound("s", in_dt), res_dt, sum_dt, mode)], P.FINAL)
    return partial, final


# ---- (a) SUM / AVG(decimal) overflow ---------------------------------------------------------------------------------------
BIG = 2 * 10**37          # six of these leave decimal(38, 0) in every row order (all positive)
NAMES = ["ovf", "ok", "empty", "neg"]


def overflow_table():
    #       group 0: 6 x BIG (overflows)   group 1: healthy    group 2: only NULLs    group 3: 6 x -BIG (overflows downwards)
    codes = [0] * 6 + [1] * 3 + [2] * 2 + [3] * 6
    vals = [BIG] * 6 + [5, 7, -3] + [None, None] + [-BIG] * 6
    return codes, vals


@pytest.mark.parametrize("mode", [LEGACY, TRY])
@pytest.mark.parametrize("dense", [True, False])
def test_sum_decimal_overflow_is_null_and_sticky(cb, oracle, mode, dense):
    P = cb.proto
    codes, vals = overflow_table()
    in_dt = sum_dt = P.DECIMAL(38, 0)
    if dense:
        key_dt, kcol = P.STRING, keys_arr(codes, NAMES)
    else:
        key_dt, kcol = P.INT64, pa.array([c * 1_000_003 for c in codes], type=pa.int64())
    partial, final = sum_plans(P, key_dt, in_dt, sum_dt, mode)
    state = run(cb, partial, [pa.table({"k": kcol, "v": dec_arr(vals, 38, 0)})])
    # oracle: the accumulator the reference runs, row by row
    acc = oracle.SumDecimalGroups(4, 38, mode)
    v = oracle.dec_from_ints(vals)
    valid = np.array([x is not None for x in vals], dtype=np.uint8)
    acc.update(v, valid, np.array(codes))
    s, sv, emp = acc.state()
    name_of = (lambda k: k) if dense else (lambda k: NAMES[k // 1_000_003])
    got = {name_of(r["col_0"]): (unscaled(r["col_1"], 0), r["col_2"]) for r in state.to_pylist()}
    exp_sums = oracle.dec_to_ints(s, sv)
    for g, name in enumerate(NAMES):
        assert got[name] == (exp_sums[g], bool(emp[g])), name          # (sum NULL after overflow, is_empty) sum_decimal.rs:526-538
    assert got["ovf"][0] is None and got["neg"][0] is None and got["ok"] == (9, False) and got["empty"] == (0, True)
    # a second, healthy partial for every group: overflow must stay NULL through the Final merge
    more_codes, more_vals = [0, 1, 2, 3], [1, 1, 1, 1]
    kcol2 = keys_arr(more_codes, NAMES) if dense else pa.array([c * 1_000_003 for c in more_codes], type=pa.int64())
    more = run(cb, partial, [pa.table({"k": kcol2, "v": dec_arr(more_vals, 38, 0)})])
    res = run(cb, final, [pa.concat_tables([state, more])])
    fin = oracle.SumDecimalGroups(4, 38, mode)
    for tbl in (state, more):
        rows = tbl.to_pylist()
        gi = np.array([NAMES.index(name_of(r["col_0"])) for r in rows])
        ps = oracle.dec_from_ints([unscaled(r["col_1"], 0) for r in rows])
        pv = np.array([r["col_1"] is not None for r in rows], dtype=np.uint8)
        pe = np.array([r["col_2"] for r in rows], dtype=np.uint8)
        fin.merge(ps, pv, pe, gi)
    out, outv = fin.evaluate()
    exp = oracle.dec_to_ints(out, outv)
    gotf = {name_of(r["col_0"]): unscaled(r["col_1"], 0) for r in res.to_pylist()}
    assert gotf == {name: exp[g] for g, name in enumerate(NAMES)}
    assert gotf == {"ovf": None, "ok": 10, "empty": 1, "neg": None}


def test_sum_decimal_overflow_ansi_raises(cb):
    P = cb.proto
    codes, vals = overflow_table()
    partial, final = sum_plans(P, P.STRING, P.DECIMAL(38, 0), P.DECIMAL(38, 0), ANSI)
    with pytest.raises(cb.native.CometB200Error) as ei:
        run(cb, partial, [pa.table({"k": keys_arr(codes, NAMES), "v": dec_arr(vals, 38, 0)})])
    assert ei.value.error_class == "ARITHMETIC_OVERFLOW"
    # two partials that fit on their own, merged past the precision: the Final stage raises (sum_decimal.rs:575-590)
    half = [4 * 10**37] * 2
    a = run(cb, partial, [pa.table({"k": keys_arr([0, 0], NAMES), "v": dec_arr(half, 38, 0)})])
    b = run(cb, partial, [pa.table({"k": keys_arr([0, 0], NAMES), "v": dec_arr(half, 38, 0)})])
    assert unscaled(a.column(1)[0].as_py(), 0) == 8 * 10**37
    with pytest.raises(cb.native.CometB200Error) as ei:
        run(cb, final, [pa.concat_tables([a, b])])
    assert ei.value.error_class == "ARITHMETIC_OVERFLOW"


def test_sum_decimal_ungrouped_overflow(cb, oracle):
    """SumDecimalAccumulator (sum_decimal.rs:176-369): the ungrouped kernel variant keeps its totals in registers."""
    P = cb.proto
    dt = P.DECIMAL(38, 2)
    vals = [BIG] * 6
    partial = P.hash_agg(P.scan([dt]), [], [P.agg_sum(P.bound(0, dt), dt, LEGACY)], P.PARTIAL)
    final = P.hash_agg(P.scan([dt, P.BOOL], source="shuffle"), [], [P.agg_sum(P.unbound("s", dt), dt, LEGACY)], P.FINAL)
    st = run(cb, partial, [pa.table({"v": dec_arr(vals, 38, 2)})])
    acc = oracle.SumDecimalAcc(38, LEGACY)
    acc.update(oracle.dec_from_ints(vals))
    s, sv, emp = acc.state()
    assert st.to_pylist() == [{"col_0": None, "col_1": False}] and not sv[0] and not emp[0]
    assert run(cb, final, [st]).column(0).to_pylist() == [None]


@pytest.mark.parametrize("dense", [True, False])
def test_avg_decimal_overflow_is_null(cb, oracle, dense):
    """avg_decimal.rs:483-495: the sum state overflows its decimal(38, s) -> is_not_null = false -> NULL average."""
    P = cb.proto
    codes, vals = overflow_table()
    in_dt, sum_dt, res_dt = P.DECIMAL(38, 0), P.DECIMAL(38, 0), P.DECIMAL(38, 4)
    if dense:
        key_dt, kcol = P.STRING, keys_arr(codes, NAMES)
    else:
        key_dt, kcol = P.INT64, pa.array([c * 7 for c in codes], type=pa.int64())
    name_of = (lambda k: k) if dense else (lambda k: NAMES[k // 7])
    partial, final = avg_plans(P, key_dt, in_dt, sum_dt, res_dt)
    state = run(cb, partial, [pa.table({"k": kcol, "v": dec_arr(vals, 38, 0)})])
    acc = oracle.AvgDecimalGroups(4, 38, 0, 38, 4)
    valid = np.array([x is not None for x in vals], dtype=np.uint8)
    acc.update(oracle.dec_from_ints(vals), valid, np.array(codes))
    sums, counts, nn = acc.state()
    got = {name_of(r["col_0"]): (unscaled(r["col_1"], 0), r["col_2"]) for r in state.to_pylist()}
    exp_s = oracle.dec_to_ints(sums, nn)
    for g, name in enumerate(NAMES):
        assert got[name] == ((exp_s[g], int(counts[g])) if nn[g] else (None, None)), name    # sums and counts share is_not_null (avg_decimal.rs:640-656)
    res = run(cb, final, [state])
    out, outv = acc.evaluate()
    exp = oracle.dec_to_ints(out, outv)
    gotf = {name_of(r["col_0"]): unscaled(r["col_1"], 4) for r in res.to_pylist()}
    assert gotf == {name: exp[g] for g, name in enumerate(NAMES)}
    assert gotf["ovf"] is None and gotf["neg"] is None and gotf["empty"] is None and gotf["ok"] == 30000


# ---- (b) AggExpr.filter ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dense", [True, False])
def test_aggregate_filter_clause(cb, oracle, dense):
    """FILTER (WHERE f): rows whose filter is FALSE *or NULL* do not reach the accumulator (sum_decimal.rs:452-458; the
    reference's KATs :732-801 cover the same three cases: all pass, some filtered, NULL filter)."""
    P = cb.proto
    n = 50_000
    rng = np.random.default_rng(11)
    codes = rng.integers(0, 4, n)
    vals = rng.integers(-10**11, 10**11, n)
    valid = rng.random(n) > 0.1
    f = rng.random(n) > 0.4
    fnull = rng.random(n) < 0.15
    codes[:50] = 3
    f[:50] = False                                   # ...
    f[codes == 3] = False                            # a group with every row filtered out: is_empty stays true
    in_dt, sum_dt, res_dt = P.DECIMAL(12, 2), P.DECIMAL(22, 2), P.DECIMAL(16, 6)
    names = ["a", "b", "c", "d"]
    if dense:
        key_dt, kcol = P.STRING, keys_arr(codes, names)
    else:
        key_dt, kcol = P.INT64, pa.array(codes * 1_000_003, type=pa.int64())
    name_of = (lambda k: k) if dense else (lambda k: names[k // 1_000_003])
    tbl = pa.table({"k": kcol, "v": dec_arr(vals.tolist(), 12, 2, mask=~valid), "f": pa.array(f, mask=fnull)})
    flt = P.bound(2, P.BOOL)
    aggs = [P.agg_sum(P.bound(1, in_dt), sum_dt, LEGACY, filter_expr=flt), P.agg_avg(P.bound(1, in_dt), res_dt, sum_dt, LEGACY, filter_expr=flt),
            P.agg_count([P.bound(1, in_dt)], filter_expr=flt), P.agg_sum(P.bound(1, in_dt), sum_dt, LEGACY), P.agg_count([P.literal(1, P.INT32)])]
    partial = P.hash_agg(P.scan([key_dt, in_dt, P.BOOL]), [P.bound(0, key_dt)], aggs, P.PARTIAL)
    state = run(cb, partial, [tbl.to_batches(max_chunksize=8192)], 20_000)
    fb = (f & ~fnull).astype(np.uint8)               # NULL filter = not selected
    v = oracle.dec_from_i64(vals)
    s_f = oracle.SumDecimalGroups(4, 22, LEGACY); s_f.update(v, valid.astype(np.uint8), codes, fb)
    a_f = oracle.AvgDecimalGroups(4, 22, 2, 16, 6); a_f.update(v, valid.astype(np.uint8), codes, fb)
    s_all = oracle.SumDecimalGroups(4, 22, LEGACY); s_all.update(v, valid.astype(np.uint8), codes)
    cnt_f = oracle.count_groups(n, valid.astype(np.uint8), codes, 4, fb)
    ss, ssv, se = s_f.state()
    asum, acnt, ann = a_f.state()
    alls, allv, alle = s_all.state()
    got = {name_of(r["col_0"]): r for r in state.to_pylist()}
    assert len(got) == 4
    for g, name in enumerate(names):
        r = got[name]
        assert (unscaled(r["col_1"], 2), r["col_2"]) == (oracle.dec_to_ints(ss, ssv)[g], bool(se[g])), name
        assert (unscaled(r["col_3"], 2), r["col_4"]) == (oracle.dec_to_ints(asum)[g], int(acnt[g])), name
        assert r["col_5"] == int(cnt_f[g]), name
        assert (unscaled(r["col_6"], 2), r["col_7"]) == (oracle.dec_to_ints(alls, allv)[g], bool(alle[g])), name
        assert r["col_8"] == int((codes == g).sum()), name
    assert got["d"]["col_2"] is True and got["d"]["col_5"] == 0       # every row of group d was filtered out


# ---- (c) exact 128-bit escape of the dense accumulators ---------------------------------------------------------------------
@pytest.mark.parametrize("ungrouped", [False, True])
def test_dense_sum_with_values_beyond_2_46_takes_the_exact_escape(cb, ungrouped):
    """Thread-private partial sums are 64-bit; |v| >= 2^46 goes through Acc::spill128 (cb_kernels.cuh).  Values up to ~2^66 with
    both signs and a chunk boundary in the middle: totals must be exact."""
    P = cb.proto
    n = 40_000
    rng = np.random.default_rng(3)
    mag = rng.integers(40, 67, n)                                      # bit lengths 40..66: both sides of the 2^46 threshold
    vals = [int(rng.integers(1 << 30, 1 << 31)) << int(m - 31) for m in mag]
    sign = rng.random(n) < 0.4
    vals = [-v if s else v for v, s in zip(vals, sign)]
    codes = rng.integers(0, 3, n)
    in_dt, sum_dt = P.DECIMAL(21, 2), P.DECIMAL(31, 2)
    tbl = pa.table({"k": keys_arr(codes, ["x", "y", "z"]), "v": dec_arr(vals, 21, 2)})
    if ungrouped:
        partial = P.hash_agg(P.scan([P.STRING, in_dt]), [], [P.agg_sum(P.bound(1, in_dt), sum_dt)], P.PARTIAL)
        st = run(cb, partial, [tbl.to_batches(max_chunksize=4096)], 15_000)
        assert unscaled(st.column(0)[0].as_py(), 2) == sum(vals)
        return
    partial, final = sum_plans(P, P.STRING, in_dt, sum_dt, LEGACY)
    st = run(cb, partial, [tbl.to_batches(max_chunksize=4096)], 15_000)
    res = run(cb, final, [st])
    got = {r["col_0"]: unscaled(r["col_1"], 2) for r in res.to_pylist()}
    for i, name in enumerate(["x", "y", "z"]):
        assert got[name] == sum(v for v, c in zip(vals, codes) if c == i), name


# ---- (d) range re-run: TIGHT -> TYPE -> SAFE --------------------------------------------------------------------------------
def test_later_chunk_violates_the_sampled_range(cb, oracle):
    """The kernel of chunk k+1 is specialised to the value ranges seen so far (+2 bits).  Later chunks carry values far beyond
    them: the launch must be discarded and re-run at the declared precision -- Q1 results stay bit-exact."""
    t = cb.tpch
    n = 90_000
    cols = t.gen_lineitem(n, seed=5)
    cols["l_extendedprice"][:30_000] = cols["l_extendedprice"][:30_000] % 1000      # first chunk: tiny prices
    cols["l_extendedprice"][60_000:] = 999_999_999_999 - cols["l_extendedprice"][60_000:] % 1000   # last chunk: the top of decimal(12,2)
    cols["l_discount"][:30_000] = 0
    tbl = t.lineitem_table(cols, "dec", dictionary=True)
    state = run(cb, t.q1_partial_plan("dec"), [tbl.to_batches(max_chunksize=10_000)], 30_000)
    res = run(cb, t.q1_final_plan("dec"), [state])
    d = oracle.dec_from_i64
    exp = oracle.q1_dec(d(cols["l_quantity"]), d(cols["l_extendedprice"]), d(cols["l_discount"]), d(cols["l_tax"]), cols["l_shipdate"],
                        cols["l_returnflag"], cols["l_linestatus"], 3, 2, t.Q1_CUTOFF, 1)
    got = {(r["col_0"], r["col_1"]): r for r in res.to_pylist()}
    seen = 0
    for k, e in enumerate(exp):
        if e is None:
            continue
        seen += 1
        g = got[(t.RETURNFLAGS[k // 2], t.LINESTATUS[k % 2])]
        assert unscaled(g["col_2"], 2) == e["sum_qty"] and unscaled(g["col_3"], 2) == e["sum_base"]
        assert unscaled(g["col_4"], 4) == e["sum_disc_price"] and unscaled(g["col_5"], 6) == e["sum_charge"]
        assert unscaled(g["col_6"], 6) == e["avg_qty"] and unscaled(g["col_7"], 6) == e["avg_price"] and unscaled(g["col_8"], 6) == e["avg_disc"]
        assert g["col_9"] == e["count"]
    assert seen == len(got)


def test_values_beyond_the_declared_precision_fall_back_to_the_checked_kernel(cb):
    """A decimal(12,2) column that carries 10^15 (not a valid decimal(12,2)): the TYPE-level kernel's assumption fails its
    value-mask validation and the SAFE (fully checked) kernel answers.  arrow-arith / SumDecimal do not re-validate inputs, so
    the plain sum is the exact sum."""
    P = cb.proto
    n = 50_000
    rng = np.random.default_rng(8)
    vals = rng.integers(0, 10**11, n)
    vals[n - 7] = 10**15
    vals[n - 3] = -(10**15) - 1
    raw = np.empty((n, 2), dtype=np.int64)
    raw[:, 0] = vals
    raw[:, 1] = vals >> 63
    arr = pa.Array.from_buffers(pa.decimal128(12, 2), n, [None, pa.py_buffer(raw.tobytes())])
    codes = rng.integers(0, 3, n)
    tbl = pa.table({"k": keys_arr(codes, ["x", "y", "z"]), "v": arr})
    partial, final = sum_plans(P, P.STRING, P.DECIMAL(12, 2), P.DECIMAL(22, 2), LEGACY)
    res = run(cb, final, [run(cb, partial, [tbl.to_batches(max_chunksize=8192)], 20_000)])
    got = {r["col_0"]: unscaled(r["col_1"], 2) for r in res.to_pylist()}
    for i, name in enumerate(["x", "y", "z"]):
        assert got[name] == int(vals[codes == i].sum()), name


# ---- (e) dense MIN / MAX ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dense", [True, False])
def test_min_max_f64_date_decimal(cb, dense):
    P = cb.proto
    n = 70_000
    rng = np.random.default_rng(21)
    codes = rng.integers(0, 5, n)
    f = rng.normal(0, 1e6, n)
    f[rng.integers(0, n, 50)] = 0.0
    dte = rng.integers(-5000, 20000, n).astype(np.int32)
    dec = rng.integers(-10**17, 10**17, n)
    i16 = rng.integers(-2**15, 2**15, n).astype(np.int16)
    vf, vd, vc = rng.random(n) > 0.2, rng.random(n) > 0.2, rng.random(n) > 0.2
    vf[codes == 4] = False                                         # MIN / MAX over only NULLs -> NULL
    names = ["a", "b", "c", "d", "e"]
    if dense:
        key_dt, kcol = P.STRING, keys_arr(codes, names)
    else:
        key_dt, kcol = P.INT64, pa.array(codes * 1_000_003, type=pa.int64())
    name_of = (lambda k: k) if dense else (lambda k: names[k // 1_000_003])
    D = P.DECIMAL(18, 3)
    tbl = pa.table({"k": kcol, "f": pa.array(f, mask=~vf), "d": pa.array(dte, type=pa.date32(), mask=~vd),
                    "c": dec_arr(dec.tolist(), 18, 3, mask=~vc), "s": pa.array(i16)})
    ins = [key_dt, P.DOUBLE, P.DATE, D, P.INT16]
    aggs, faggs, state_t = [], [], [key_dt]
    for i, dt in ((1, P.DOUBLE), (2, P.DATE), (3, D), (4, P.INT16)):
        aggs += [P.agg_min(P.bound(i, dt), dt), P.agg_max(P.bound(i, dt), dt)]
        faggs += [P.agg_min(P.unbound("x", dt), dt), P.agg_max(P.unbound("x", dt), dt)]
        state_t += [dt, dt]
    partial = P.hash_agg(P.scan(ins), [P.bound(0, key_dt)], aggs, P.PARTIAL)
    final = P.hash_agg(P.scan(state_t, source="shuffle"), [P.bound(0, key_dt)], faggs, P.FINAL)
    res = run(cb, final, [run(cb, partial, [tbl.to_batches(max_chunksize=8192)], 25_000)])
    got = {name_of(r["col_0"]): r for r in res.to_pylist()}
    import datetime
    epoch = datetime.date(1970, 1, 1)
    for g, name in enumerate(names):
        m = codes == g
        r = got[name]
        for col, arr, valid, conv in ((1, f, vf, float), (3, dte, vd, lambda x: epoch + datetime.timedelta(days=int(x))),
                                      (5, dec, vc, lambda x: decimal.Decimal(int(x)).scaleb(-3)), (7, i16, np.ones(n, bool), int)):
            sel = arr[m & valid]
            exp = (None, None) if sel.size == 0 else (conv(sel.min()), conv(sel.max()))
            assert (r[f"col_{col}"], r[f"col_{col + 1}"]) == exp, (name, col)


# ---- (f) under-aligned Decimal128 buffers -------------------------------------------------------------------------------------
def test_decimal128_buffers_that_are_only_8_byte_aligned(cb, oracle):
    """The JVM hands over Decimal128 buffers with 8-byte alignment (aligned_stream_reader.rs:25-32 re-aligns them)."""
    t = cb.tpch
    n = 33_333
    cols = t.gen_lineitem(n, seed=13)
    tbl = t.lineitem_table(cols, "dec", dictionary=True)
    keep = []

    def misalign(cents):
        raw = np.empty(2 * n + 3, dtype=np.int64)
        base = raw.ctypes.data
        off = 1 if (base % 16) == 0 else 0                    # start on an address that is 8 mod 16
        view = raw[off:off + 2 * n].reshape(n, 2)
        view[:, 0] = cents
        view[:, 1] = cents >> 63
        assert view.ctypes.data % 16 == 8
        keep.append(raw)
        buf = pa.foreign_buffer(view.ctypes.data, 16 * n, base=raw)
        return pa.Array.from_buffers(pa.decimal128(12, 2), n, [None, buf])
    for i, name in enumerate(["l_quantity", "l_extendedprice", "l_discount", "l_tax"]):
        tbl = tbl.set_column(i, name, misalign(cols[name]))
    state = run(cb, t.q1_partial_plan("dec"), [tbl.to_batches(max_chunksize=5000)], 12_000)
    res = run(cb, t.q1_final_plan("dec"), [state])
    d = oracle.dec_from_i64
    exp = oracle.q1_dec(d(cols["l_quantity"]), d(cols["l_extendedprice"]), d(cols["l_discount"]), d(cols["l_tax"]), cols["l_shipdate"],
                        cols["l_returnflag"], cols["l_linestatus"], 3, 2, t.Q1_CUTOFF, 1)
    got = {(r["col_0"], r["col_1"]): r for r in res.to_pylist()}
    for k, e in enumerate(exp):
        if e is None:
            continue
        g = got[(t.RETURNFLAGS[k // 2], t.LINESTATUS[k % 2])]
        assert unscaled(g["col_5"], 6) == e["sum_charge"] and g["col_9"] == e["count"]


# ---- (g) concurrent plan handles ---------------------------------------------------------------------------------------------
def test_two_plan_handles_from_two_threads(cb, oracle):
    """One handle per task thread, many tasks at once (jni_api.rs:194-223): two Q1 partial+final pipelines over different data,
    interleaved batch by batch from two Python threads (ctypes releases the GIL inside the library)."""
    t = cb.tpch
    results, errors = {}, []

    def work(seed, n):
        try:
            cols = t.gen_lineitem(n, seed=seed)
            tbl = t.lineitem_table(cols, "dec", dictionary=True)
            for rep in range(3):
                state = run(cb, t.q1_partial_plan("dec"), [tbl.to_batches(max_chunksize=4096)], 16_384)
                res = run(cb, t.q1_final_plan("dec"), [state])
            results[seed] = (cols, res)
        except Exception as e:  # noqa: BLE001
            errors.append(e)
    th = [threading.Thread(target=work, args=(s, n)) for s, n in ((101, 150_000), (202, 90_000), (303, 120_000))]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors
    d = oracle.dec_from_i64
    for seed, (cols, res) in results.items():
        exp = oracle.q1_dec(d(cols["l_quantity"]), d(cols["l_extendedprice"]), d(cols["l_discount"]), d(cols["l_tax"]), cols["l_shipdate"],
                            cols["l_returnflag"], cols["l_linestatus"], 3, 2, t.Q1_CUTOFF, 1)
        got = {(r["col_0"], r["col_1"]): r for r in res.to_pylist()}
        for k, e in enumerate(exp):
            if e is None:
                continue
            g = got[(t.RETURNFLAGS[k // 2], t.LINESTATUS[k % 2])]
            assert unscaled(g["col_5"], 6) == e["sum_charge"] and unscaled(g["col_8"], 6) == e["avg_disc"] and g["col_9"] == e["count"], seed


# ---- (i) spark.comet.batchSize on export (CometConf.scala:539-544, prepare_output jni_api.rs:674-742) ---------------------------------
@pytest.mark.parametrize("batch_size", [1000, 1001, 4096])
def test_output_batches_respect_the_batch_size(cb, batch_size):
    """A result larger than spark.comet.batchSize leaves cb200_execute in zero-offset slices of at most that many rows -- values, NULLs
    (bitmap slices that do not start on a byte), booleans and dictionary strings all line up with the unsliced result."""
    P = cb.proto
    rng = np.random.default_rng(7)
    n = 10_007
    a = rng.integers(-1000, 1000, n)
    mask = rng.random(n) < 0.2
    tbl = pa.table({"a": pa.array(a, mask=mask), "w": pa.DictionaryArray.from_arrays(pa.array(rng.integers(0, 4, n).astype(np.int32)), pa.array(["x", "yy", "zzz", ""])),
                    "d": dec_arr([int(v) * 7 for v in a], 12, 2, mask=np.roll(mask, 3))})
    plan = P.projection(P.scan([P.INT64, P.STRING, P.DECIMAL(12, 2)]),
                        [P.bound(0, P.INT64), P.gt(P.bound(0, P.INT64), P.literal(0, P.INT64)), P.bound(1, P.STRING), P.bound(2, P.DECIMAL(12, 2))])
    sizes = []
    with cb.native.Plan(plan, [tbl.to_batches(max_chunksize=3000)], batch_size=batch_size) as p:
        batches = []
        while True:
            b = p.execute()
            if b is None:
                break
            sizes.append(b.num_rows)
            batches.append(b)
    assert sum(sizes) == n and max(sizes) <= batch_size and all(s == batch_size for s in sizes[:-1])
    got = pa.Table.from_batches(batches)
    assert got.column(0).combine_chunks().equals(tbl.column("a").combine_chunks())
    exp_gt = pa.array([None if m else bool(v > 0) for v, m in zip(a.tolist(), mask.tolist())])
    assert got.column(1).combine_chunks().equals(exp_gt)
    assert got.column(2).to_pylist() == tbl.column("w").to_pylist()
    assert got.column(3).combine_chunks().equals(tbl.column("d").combine_chunks())
