"""ctypes binding for the CPU oracle (oracle/libcomet_oracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/comet_oracle.h.  Imported by tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke() as the checker; never by the
product package.

Decimal128 columns are numpy arrays of shape (n, 2) uint64 (little-endian lo, hi limbs), which is
bit-identical to Arrow's Decimal128 value buffer.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LEGACY, TRY, ANSI = 0, 1, 2


def build(force=False):
    so = os.path.join(_HERE, "libcomet_oracle.so")
    src = os.path.join(_HERE, "comet_oracle.c")
    hdr = os.path.join(_HERE, "comet_oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libcomet_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.co_murmur3_bytes.restype = C.c_uint32
        _LIB.co_murmur3_bytes.argtypes = [C.c_void_p, C.c_int64, C.c_uint32]
        _LIB.co_pmod.restype = C.c_uint32
        _LIB.co_pmod.argtypes = [C.c_uint32, C.c_uint32]
        _LIB.co_filter_project_dec.restype = C.c_int64
        _LIB.co_filter_project_f64.restype = C.c_int64
    return _LIB


# ---- decimal <-> python int helpers -------------------------------------------------------------
def dec_from_ints(vals):
    """list of python ints (None -> 0) -> (n,2) uint64"""
    out = np.zeros((len(vals), 2), dtype=np.uint64)
    for i, v in enumerate(vals):
        v = 0 if v is None else int(v)
        u = v & ((1 << 128) - 1)
        out[i, 0] = u & ((1 << 64) - 1)
        out[i, 1] = u >> 64
    return out


def dec_to_ints(arr, valid=None):
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 2)
    out = []
    for i in range(arr.shape[0]):
        if valid is not None and not valid[i]:
            out.append(None)
            continue
        u = int(arr[i, 0]) | (int(arr[i, 1]) << 64)
        if u >= 1 << 127:
            u -= 1 << 128
        out.append(u)
    return out


def dec_from_i64(a):
    """int64 numpy array -> sign-extended (n,2) uint64"""
    a = np.ascontiguousarray(a, dtype=np.int64)
    out = np.empty((a.shape[0], 2), dtype=np.uint64)
    out[:, 0] = a.view(np.uint64)
    out[:, 1] = (a >> 63).view(np.uint64)
    return out


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _valid(v, n):
    if v is None:
        return None
    v = np.ascontiguousarray(v, dtype=np.uint8)
    assert v.shape[0] == n
    return v


class OracleError(Exception):
    def __init__(self, code):
        super().__init__({1: "ARITHMETIC_OVERFLOW", 2: "DIVIDE_BY_ZERO", 3: "INVALID"}.get(code, str(code)))
        self.code = code


def _chk(rc):
    if rc != 0:
        raise OracleError(rc)


# ---- decimal elementwise ------------------------------------------------------------------------
def wide_decimal(op, l, lv, s1, r, rv, s2, p_out, s_out, eval_mode=LEGACY):
    n = l.shape[0]
    out = np.zeros((n, 2), dtype=np.uint64)
    outv = np.zeros(n, dtype=np.uint8)
    lv, rv = _valid(lv, n), _valid(rv, n)
    _chk(lib().co_wide_decimal(C.c_int(op), C.c_int64(n), _p(l), _p(lv), C.c_int(s1), _p(r), _p(rv), C.c_int(s2),
                               C.c_int(p_out), C.c_int(s_out), C.c_int(eval_mode), _p(out), _p(outv)))
    return out, outv


def plain_decimal(op, l, lv, p1, s1, r, rv, p2, s2):
    n = l.shape[0]
    out = np.zeros((n, 2), dtype=np.uint64)
    outv = np.zeros(n, dtype=np.uint8)
    pr, sr = C.c_int(0), C.c_int(0)
    lv, rv = _valid(lv, n), _valid(rv, n)
    _chk(lib().co_plain_decimal(C.c_int(op), C.c_int64(n), _p(l), _p(lv), C.c_int(p1), C.c_int(s1), _p(r), _p(rv),
                                C.c_int(p2), C.c_int(s2), _p(out), _p(outv), C.byref(pr), C.byref(sr)))
    return out, outv, pr.value, sr.value


def check_overflow(a, av, precision, fail_on_error=False):
    n = a.shape[0]
    out = np.zeros((n, 2), dtype=np.uint64)
    outv = np.zeros(n, dtype=np.uint8)
    av = _valid(av, n)
    _chk(lib().co_check_overflow(C.c_int64(n), _p(a), _p(av), C.c_int(precision), C.c_int(int(fail_on_error)),
                                 _p(out), _p(outv)))
    return out, outv


def decimal_rescale_check(a, av, s_in, p_out, s_out, fail_on_error=False):
    n = a.shape[0]
    out = np.zeros((n, 2), dtype=np.uint64)
    outv = np.zeros(n, dtype=np.uint8)
    av = _valid(av, n)
    _chk(lib().co_decimal_rescale_check(C.c_int64(n), _p(a), _p(av), C.c_int(s_in), C.c_int(p_out), C.c_int(s_out),
                                        C.c_int(int(fail_on_error)), _p(out), _p(outv)))
    return out, outv


def decimal_div(l, lv, s1, r, rv, s2, s3, integral=False, eval_mode=LEGACY, check_divide_overflow=False):
    """spark-expr/src/math_funcs/div.rs:75-190 (spark_decimal_div_internal), restated with Python integers -- the reference
    itself falls back to BigInt for wide operands.  Rows where either side is NULL are not evaluated (try_binary).  Returns
    ((n,2) uint64, valid); raises OracleError(2) for a zero divisor in ANSI mode, OracleError(1) for the integral overflow."""
    ls, rs = dec_to_ints(l), dec_to_ints(r)
    n = len(ls)
    lv = np.ones(n, dtype=bool) if lv is None else np.asarray(lv, dtype=bool)
    rv = np.ones(n, dtype=bool) if rv is None else np.asarray(rv, dtype=bool)
    l_exp, r_exp = max(0, s2 + s3 + 1 - s1), max(0, s1 - (s2 + s3 + 1))
    out, valid = [0] * n, lv & rv
    for i in range(n):
        if not valid[i]:
            continue
        a, b = ls[i] * 10 ** l_exp, rs[i] * 10 ** r_exp
        if b == 0:
            if eval_mode == ANSI:
                raise OracleError(2)
            div = 0
        else:
            div = abs(a) // abs(b)          # BigInt `/` truncates toward zero
            if (a < 0) != (b < 0):
                div = -div
        if integral:
            res = div
        else:
            t = div - 5 if div < 0 else div + 5
            res = -((-t) // 10) if t < 0 else t // 10
        if not (-(1 << 127) <= res < (1 << 127)):
            res = (1 << 127) - 1            # to_i128().unwrap_or(i128::MAX)
        if check_divide_overflow and eval_mode == ANSI and not (-(1 << 63) <= res < (1 << 63)):
            raise OracleError(1)
        out[i] = res
    return dec_from_ints(out), valid.astype(np.uint8)


def int_arith(op, width, l, lv, r, rv, eval_mode=LEGACY):
    l = np.ascontiguousarray(l, dtype=np.int64)
    r = np.ascontiguousarray(r, dtype=np.int64)
    n = l.shape[0]
    out = np.zeros(n, dtype=np.int64)
    outv = np.zeros(n, dtype=np.uint8)
    lv, rv = _valid(lv, n), _valid(rv, n)
    _chk(lib().co_int_arith(C.c_int(op), C.c_int(width), C.c_int64(n), _p(l), _p(lv), _p(r), _p(rv),
                            C.c_int(eval_mode), _p(out), _p(outv)))
    return out, outv


# ---- hashing ------------------------------------------------------------------------------------
KIND = {"bool": 0, "i8": 1, "i16": 2, "i32": 3, "i64": 4, "f32": 5, "f64": 6, "date32": 7, "timestamp": 8,
        "dec_small": 9, "dec_large": 10}
_KIND_DT = {0: np.uint8, 1: np.int8, 2: np.int16, 3: np.int32, 4: np.int64, 5: np.float32, 6: np.float64,
            7: np.int32, 8: np.int64}


def murmur3_bytes(b, seed=42):
    buf = (C.c_uint8 * len(b)).from_buffer_copy(b) if len(b) else None
    return lib().co_murmur3_bytes(buf, len(b), seed)


def murmur3_column(kind, values, valid=None, hashes=None, seed=42):
    k = KIND[kind] if isinstance(kind, str) else kind
    if k in _KIND_DT:
        values = np.ascontiguousarray(values, dtype=_KIND_DT[k])
        n = values.shape[0]
    else:
        values = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1, 2)
        n = values.shape[0]
    if hashes is None:
        hashes = np.full(n, seed, dtype=np.uint32)
    valid = _valid(valid, n)
    _chk(lib().co_murmur3_column(C.c_int(k), C.c_int64(n), _p(values), _p(valid), _p(hashes)))
    return hashes


def murmur3_strings(strings, hashes=None, seed=42):
    """strings: list of bytes/str/None"""
    n = len(strings)
    valid = np.array([s is not None for s in strings], dtype=np.uint8)
    bs = [(s.encode() if isinstance(s, str) else (s or b"")) for s in strings]
    offsets = np.zeros(n + 1, dtype=np.int32)
    offsets[1:] = np.cumsum([len(b) for b in bs])
    data = np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8).copy()
    if hashes is None:
        hashes = np.full(n, seed, dtype=np.uint32)
    lib().co_murmur3_strings(C.c_int64(n), _p(offsets), _p(data), _p(valid), _p(hashes))
    return hashes


def pmod(h, n):
    return lib().co_pmod(int(h), int(n))


def partition_rows(hashes, n_parts):
    hashes = np.ascontiguousarray(hashes, dtype=np.uint32)
    n = hashes.shape[0]
    pids = np.zeros(n, dtype=np.uint32)
    starts = np.zeros(n_parts + 1, dtype=np.int64)
    row_idx = np.zeros(n, dtype=np.int64)
    lib().co_partition_rows(C.c_int64(n), _p(hashes), C.c_uint32(n_parts), _p(pids), _p(starts), _p(row_idx))
    return pids, starts, row_idx


# ---- accumulators -------------------------------------------------------------------------------
class SumDecimalGroups:
    """spark-expr/src/agg_funcs/sum_decimal.rs:371-611"""

    def __init__(self, n_groups, precision, eval_mode=LEGACY):
        self.ng, self.p, self.mode = n_groups, precision, eval_mode
        self.sum = np.zeros((n_groups, 2), dtype=np.uint64)
        self.sum_valid = np.ones(n_groups, dtype=np.uint8)
        self.is_empty = np.ones(n_groups, dtype=np.uint8)

    def update(self, v, valid, group_idx, filt=None):
        n = v.shape[0]
        g = None if group_idx is None else np.ascontiguousarray(group_idx, dtype=np.int64)
        _chk(lib().co_sum_decimal_update(C.c_int64(n), _p(v), _p(_valid(valid, n)), _p(_valid(filt, n)), _p(g),
                                         _p(self.sum), _p(self.sum_valid), _p(self.is_empty), C.c_int(self.p),
                                         C.c_int(self.mode)))

    def merge(self, that_sum, that_sum_valid, that_is_empty, group_idx):
        n = that_sum.shape[0]
        g = None if group_idx is None else np.ascontiguousarray(group_idx, dtype=np.int64)
        te = np.ascontiguousarray(that_is_empty, dtype=np.uint8)
        _chk(lib().co_sum_decimal_merge(C.c_int64(n), _p(that_sum), _p(_valid(that_sum_valid, n)), _p(te), _p(g),
                                        _p(self.sum), _p(self.sum_valid), _p(self.is_empty), C.c_int(self.p),
                                        C.c_int(self.mode)))

    def state(self):
        return self.sum.copy(), self.sum_valid.copy(), self.is_empty.copy()

    def evaluate(self):
        out = np.zeros((self.ng, 2), dtype=np.uint64)
        outv = np.zeros(self.ng, dtype=np.uint8)
        lib().co_sum_decimal_evaluate(C.c_int64(self.ng), _p(self.sum), _p(self.sum_valid), _p(self.is_empty),
                                      C.c_int(self.p), _p(out), _p(outv))
        return out, outv


class SumDecimalAcc(SumDecimalGroups):
    """ungrouped accumulator, sum_decimal.rs:176-369"""

    def __init__(self, precision, eval_mode=LEGACY):
        super().__init__(1, precision, eval_mode)

    def update(self, v, valid=None):
        n = v.shape[0]
        _chk(lib().co_sum_decimal_acc_update(C.c_int64(n), _p(v), _p(_valid(valid, n)), _p(self.sum),
                                             _p(self.sum_valid), _p(self.is_empty), C.c_int(self.p), C.c_int(self.mode)))


class AvgDecimalGroups:
    """spark-expr/src/agg_funcs/avg_decimal.rs:410-668"""

    def __init__(self, n_groups, sum_precision, sum_scale, target_precision, target_scale, eval_mode=LEGACY):
        self.ng, self.sp, self.ss, self.tp, self.ts, self.mode = n_groups, sum_precision, sum_scale, target_precision, target_scale, eval_mode
        self.sums = np.zeros((n_groups, 2), dtype=np.uint64)
        self.counts = np.zeros(n_groups, dtype=np.int64)
        self.is_not_null = np.ones(n_groups, dtype=np.uint8)

    def update(self, v, valid, group_idx, filt=None):
        n = v.shape[0]
        g = None if group_idx is None else np.ascontiguousarray(group_idx, dtype=np.int64)
        _chk(lib().co_avg_decimal_update(C.c_int64(n), _p(v), _p(_valid(valid, n)), _p(_valid(filt, n)), _p(g),
                                         _p(self.sums), _p(self.counts), _p(self.is_not_null), C.c_int(self.sp)))

    def merge(self, psum, psum_valid, pcount, pcount_valid, group_idx):
        n = psum.shape[0]
        g = None if group_idx is None else np.ascontiguousarray(group_idx, dtype=np.int64)
        pc = np.ascontiguousarray(pcount, dtype=np.int64)
        _chk(lib().co_avg_decimal_merge(C.c_int64(n), _p(psum), _p(_valid(psum_valid, n)), _p(pc),
                                        _p(_valid(pcount_valid, n)), _p(g), _p(self.sums), _p(self.counts),
                                        _p(self.is_not_null), C.c_int(self.sp), C.c_int(self.mode)))

    def state(self):
        # sums and counts share the is_not_null buffer as validity (avg_decimal.rs:640-656)
        return self.sums.copy(), self.counts.copy(), self.is_not_null.copy()

    def evaluate(self):
        out = np.zeros((self.ng, 2), dtype=np.uint64)
        outv = np.zeros(self.ng, dtype=np.uint8)
        _chk(lib().co_avg_decimal_evaluate(C.c_int64(self.ng), _p(self.sums), _p(self.counts), _p(self.is_not_null),
                                           C.c_int(self.ss), C.c_int(self.tp), C.c_int(self.ts), C.c_int(self.mode),
                                           _p(out), _p(outv)))
        return out, outv


class AvgF64Groups:
    """spark-expr/src/agg_funcs/avg.rs:201-347"""

    def __init__(self, n_groups):
        self.ng = n_groups
        self.sums = np.zeros(n_groups, dtype=np.float64)
        self.counts = np.zeros(n_groups, dtype=np.int64)

    def update(self, v, valid, group_idx, filt=None):
        v = np.ascontiguousarray(v, dtype=np.float64)
        n = v.shape[0]
        g = None if group_idx is None else np.ascontiguousarray(group_idx, dtype=np.int64)
        lib().co_avg_f64_update(C.c_int64(n), _p(v), _p(_valid(valid, n)), _p(_valid(filt, n)), _p(g), _p(self.sums),
                                _p(self.counts))

    def merge(self, psum, pcount, group_idx):
        psum = np.ascontiguousarray(psum, dtype=np.float64)
        pcount = np.ascontiguousarray(pcount, dtype=np.int64)
        g = None if group_idx is None else np.ascontiguousarray(group_idx, dtype=np.int64)
        lib().co_avg_f64_merge(C.c_int64(psum.shape[0]), _p(psum), _p(pcount), _p(g), _p(self.sums), _p(self.counts))

    def evaluate(self):
        out = np.zeros(self.ng, dtype=np.float64)
        outv = np.zeros(self.ng, dtype=np.uint8)
        lib().co_avg_f64_evaluate(C.c_int64(self.ng), _p(self.sums), _p(self.counts), _p(out), _p(outv))
        return out, outv


class SumIntGroups:
    """spark-expr/src/agg_funcs/sum_int.rs:393-530 (+Ansi/Try)"""

    def __init__(self, n_groups, eval_mode=LEGACY):
        self.ng, self.mode = n_groups, eval_mode
        self.sums = np.zeros(n_groups, dtype=np.int64)
        self.sums_valid = np.zeros(n_groups, dtype=np.uint8)
        self.overflowed = np.zeros(n_groups, dtype=np.uint8)

    def update(self, v, valid, group_idx, filt=None):
        v = np.ascontiguousarray(v, dtype=np.int64)
        n = v.shape[0]
        g = None if group_idx is None else np.ascontiguousarray(group_idx, dtype=np.int64)
        _chk(lib().co_sum_int_update(C.c_int64(n), _p(v), _p(_valid(valid, n)), _p(_valid(filt, n)), _p(g),
                                     _p(self.sums), _p(self.sums_valid), _p(self.overflowed), C.c_int(self.mode)))


def sum_f64_groups(v, valid, group_idx, n_groups, filt=None):
    v = np.ascontiguousarray(v, dtype=np.float64)
    n = v.shape[0]
    sums = np.zeros(n_groups, dtype=np.float64)
    sv = np.zeros(n_groups, dtype=np.uint8)
    g = None if group_idx is None else np.ascontiguousarray(group_idx, dtype=np.int64)
    lib().co_sum_f64_update(C.c_int64(n), _p(v), _p(_valid(valid, n)), _p(_valid(filt, n)), _p(g), _p(sums), _p(sv))
    return sums, sv


def sum_f64_exact(v, valid, group_idx, n_groups, filt=None):
    v = np.ascontiguousarray(v, dtype=np.float64)
    n = v.shape[0]
    out = np.zeros(n_groups, dtype=np.float64)
    g = None if group_idx is None else np.ascontiguousarray(group_idx, dtype=np.int64)
    lib().co_sum_f64_exact(C.c_int64(n), _p(v), _p(_valid(valid, n)), _p(_valid(filt, n)), _p(g), C.c_int64(n_groups),
                           _p(out))
    return out


def count_groups(n, valid, group_idx, n_groups, filt=None):
    counts = np.zeros(n_groups, dtype=np.int64)
    g = None if group_idx is None else np.ascontiguousarray(group_idx, dtype=np.int64)
    lib().co_count_update(C.c_int64(n), _p(_valid(valid, n)), _p(_valid(filt, n)), _p(g), _p(counts))
    return counts


# ---- whole-pipeline baselines -------------------------------------------------------------------
class Q1DecRow(C.Structure):
    _fields_ = [(k, C.c_uint64 * 2) for k in
                ("sum_qty", "sum_base", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc")] + \
               [("count", C.c_int64)] + \
               [(k, C.c_uint8) for k in ("v_sum_qty", "v_sum_base", "v_sum_disc_price", "v_sum_charge", "v_avg_qty",
                                         "v_avg_price", "v_avg_disc", "present")]


class Q1F64Row(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("sum_qty", "sum_base", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc")] + \
               [("count", C.c_int64), ("present", C.c_uint8)]


def _i128_of(field):
    u = int(field[0]) | (int(field[1]) << 64)
    return u - (1 << 128) if u >= 1 << 127 else u


def q1_dec(qty, price, disc, tax, shipdate, rf, ls, n_rf, n_ls, cutoff, n_threads=0):
    n = shipdate.shape[0]
    out = (Q1DecRow * (n_rf * n_ls))()
    _chk(lib().co_q1_dec(C.c_int64(n), _p(qty), _p(price), _p(disc), _p(tax), _p(shipdate), _p(rf), _p(ls),
                         C.c_int(n_rf), C.c_int(n_ls), C.c_int32(cutoff), C.c_int(n_threads), out))
    rows = []
    for k in range(n_rf * n_ls):
        o = out[k]
        if not o.present:
            rows.append(None)
            continue
        rows.append({
            "sum_qty": _i128_of(o.sum_qty) if o.v_sum_qty else None,
            "sum_base": _i128_of(o.sum_base) if o.v_sum_base else None,
            "sum_disc_price": _i128_of(o.sum_disc_price) if o.v_sum_disc_price else None,
            "sum_charge": _i128_of(o.sum_charge) if o.v_sum_charge else None,
            "avg_qty": _i128_of(o.avg_qty) if o.v_avg_qty else None,
            "avg_price": _i128_of(o.avg_price) if o.v_avg_price else None,
            "avg_disc": _i128_of(o.avg_disc) if o.v_avg_disc else None,
            "count": o.count,
        })
    return rows


def q1_f64(qty, price, disc, tax, shipdate, rf, ls, n_rf, n_ls, cutoff, n_threads=0):
    n = shipdate.shape[0]
    out = (Q1F64Row * (n_rf * n_ls))()
    _chk(lib().co_q1_f64(C.c_int64(n), _p(qty), _p(price), _p(disc), _p(tax), _p(shipdate), _p(rf), _p(ls),
                         C.c_int(n_rf), C.c_int(n_ls), C.c_int32(cutoff), C.c_int(n_threads), out))
    rows = []
    for k in range(n_rf * n_ls):
        o = out[k]
        rows.append(None if not o.present else {f: getattr(o, f) for f, _ in Q1F64Row._fields_ if f != "present"})
    return rows


def _i128_arg(v):
    u = int(v) & ((1 << 128) - 1)
    arr = (C.c_uint64 * 2)(u & ((1 << 64) - 1), u >> 64)
    return arr


class _I128(C.Structure):
    _fields_ = [("lo", C.c_uint64), ("hi", C.c_uint64)]


def _i128_byval(v):
    u = int(v) & ((1 << 128) - 1)
    return _I128(u & ((1 << 64) - 1), u >> 64)


def q6_dec(qty, price, disc, shipdate, date_lo, date_hi, disc_lo, disc_hi, qty_max, n_threads=0):
    n = shipdate.shape[0]
    out = np.zeros((1, 2), dtype=np.uint64)
    outv = np.zeros(1, dtype=np.uint8)
    _chk(lib().co_q6_dec(C.c_int64(n), _p(qty), _p(price), _p(disc), _p(shipdate), C.c_int32(date_lo),
                         C.c_int32(date_hi), _p(dec_from_ints([disc_lo])), _p(dec_from_ints([disc_hi])), _p(dec_from_ints([qty_max])),
                         C.c_int(n_threads), _p(out), _p(outv)))
    return dec_to_ints(out, outv)[0]


def q6_f64(qty, price, disc, shipdate, date_lo, date_hi, disc_lo, disc_hi, qty_max, n_threads=0):
    n = shipdate.shape[0]
    out = C.c_double(0)
    outv = C.c_uint8(0)
    _chk(lib().co_q6_f64(C.c_int64(n), _p(qty), _p(price), _p(disc), _p(shipdate), C.c_int32(date_lo),
                         C.c_int32(date_hi), C.c_double(disc_lo), C.c_double(disc_hi), C.c_double(qty_max),
                         C.c_int(n_threads), C.byref(out), C.byref(outv)))
    return out.value if outv.value else None


def filter_project_dec(qty, price, shipdate, cutoff, n_threads=0):
    n = shipdate.shape[0]
    out = np.zeros((n, 2), dtype=np.uint64)
    outv = np.zeros(n, dtype=np.uint8)
    m = lib().co_filter_project_dec(C.c_int64(n), _p(qty), _p(price), _p(shipdate), C.c_int32(cutoff),
                                    C.c_int(n_threads), _p(out), _p(outv))
    return out[:m], outv[:m]


def filter_project_f64(qty, price, shipdate, cutoff, n_threads=0):
    n = shipdate.shape[0]
    out = np.zeros(n, dtype=np.float64)
    m = lib().co_filter_project_f64(C.c_int64(n), _p(qty), _p(price), _p(shipdate), C.c_int32(cutoff),
                                    C.c_int(n_threads), _p(out))
    return out[:m]


def groupby_sum_dec(keys, vals, n_threads=0, precision=22, want_rows=False):
    """GROUP BY key SUM(value): number of groups (and, with want_rows, (keys, sums as python ints or None))."""
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    n = keys.shape[0]
    f = lib().co_groupby_sum_dec
    f.restype = C.c_int64
    if not want_rows:
        r = f(C.c_int64(n), _p(keys), _p(vals), C.c_int(precision), C.c_int(n_threads), None, None, None)
        if r < 0:
            raise MemoryError("co_groupby_sum_dec")
        return int(r)
    ok, os_, ov = np.zeros(n, dtype=np.int64), np.zeros((n, 2), dtype=np.uint64), np.zeros(n, dtype=np.uint8)
    r = f(C.c_int64(n), _p(keys), _p(vals), C.c_int(precision), C.c_int(n_threads), _p(ok), _p(os_), _p(ov))
    if r < 0:
        raise MemoryError("co_groupby_sum_dec")
    return ok[:r], dec_to_ints(os_[:r], ov[:r])


def numa_spread(a, n_threads=0):
    """A copy of `a` whose pages are first touched by the OpenMP threads that will scan them (co_parallel_copy): on a multi-socket
    host a numpy array filled by one thread lives on one socket and every other socket reads it remotely."""
    a = np.ascontiguousarray(a)
    out = np.empty_like(a)
    n = a.shape[0]
    if n == 0:
        return out
    lib().co_parallel_copy(_p(out), _p(a), C.c_int64(n), C.c_int(a.nbytes // n), C.c_int(n_threads))
    return out
