/*
 * comet_oracle.c -- CPU restatement of the reference's hot-path semantics.  TEST INFRASTRUCTURE
 * ONLY (see comet_oracle.h).  Every function cites the reference file:line it restates
 * (paths relative to the apache/datafusion-comet tree, native/ prefix omitted where obvious).
 */
#include "comet_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef co_i128 i128;

/* =========================================================================================
 * 256-bit two's-complement integer (arrow-buffer i256 as used by wide_decimal_binary_expr.rs)
 * ========================================================================================= */
typedef struct { uint64_t w[4]; } i256; /* little-endian limbs */

static i256 i256_from_i128(i128 v) {
    i256 r;
    r.w[0] = (uint64_t)(u128)v;
    r.w[1] = (uint64_t)((u128)v >> 64);
    r.w[2] = r.w[3] = v < 0 ? ~0ULL : 0ULL;
    return r;
}
static int i256_is_neg(i256 a) { return (int)(a.w[3] >> 63); }
static i256 i256_add(i256 a, i256 b) {
    i256 r; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.w[i] + b.w[i]; r.w[i] = (uint64_t)c; c >>= 64; }
    return r;
}
static i256 i256_not(i256 a) { for (int i = 0; i < 4; i++) a.w[i] = ~a.w[i]; return a; }
static i256 i256_neg(i256 a) { i256 one = {{1, 0, 0, 0}}; return i256_add(i256_not(a), one); }
static i256 i256_sub(i256 a, i256 b) { return i256_add(a, i256_neg(b)); }
static i256 i256_mul(i256 a, i256 b) { /* wrapping */
    i256 r = {{0, 0, 0, 0}};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; i + j < 4; j++) {
            c += (u128)a.w[i] * b.w[j] + r.w[i + j];
            r.w[i + j] = (uint64_t)c; c >>= 64;
        }
    }
    return r;
}
static int i256_ucmp(i256 a, i256 b) {
    for (int i = 3; i >= 0; i--) { if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1; }
    return 0;
}
static int i256_cmp(i256 a, i256 b) {
    int na = i256_is_neg(a), nb = i256_is_neg(b);
    if (na != nb) return na ? -1 : 1;
    return i256_ucmp(a, b);
}
static i256 i256_abs(i256 a) { return i256_is_neg(a) ? i256_neg(a) : a; }
/* unsigned divide by a u64, returns remainder */
static uint64_t i256_udiv_small(i256 *a, uint64_t d) {
    u128 rem = 0;
    for (int i = 3; i >= 0; i--) {
        u128 cur = (rem << 64) | a->w[i];
        a->w[i] = (uint64_t)(cur / d); rem = cur % d;
    }
    return (uint64_t)rem;
}
static i256 i256_pow10(unsigned e) { /* wide_decimal_binary_expr.rs:150-158 */
    i256 r = {{1, 0, 0, 0}}, ten = {{10, 0, 0, 0}};
    for (unsigned i = 0; i < e; i++) r = i256_mul(r, ten);
    return r;
}
/* unsigned 256/256 division for divisor = 10^e (e <= 76): chunked by 10^19 */
static i256 i256_udiv_pow10(i256 a, unsigned e) {
    while (e >= 19) { i256_udiv_small(&a, 10000000000000000000ULL); e -= 19; }
    if (e) { uint64_t d = 1; for (unsigned i = 0; i < e; i++) d *= 10; i256_udiv_small(&a, d); }
    return a;
}
/* wide_decimal_binary_expr.rs:121-144 div_round_half_up (divisor = 10^e > 0): truncated quotient,
 * round away from zero when |rem|*2 >= |divisor|. */
static i256 i256_div_pow10_half_up(i256 value, unsigned e) {
    i256 divisor = i256_pow10(e);
    int neg = i256_is_neg(value);
    i256 av = i256_abs(value);
    i256 q = i256_udiv_pow10(av, e);
    i256 rem = i256_sub(av, i256_mul(q, divisor));
    i256 two = {{2, 0, 0, 0}}, one = {{1, 0, 0, 0}};
    if (i256_ucmp(i256_mul(rem, two), divisor) >= 0) q = i256_add(q, one);
    return neg ? i256_neg(q) : q;
}
static i128 i256_to_i128(i256 a) { return (i128)(((u128)a.w[1] << 64) | a.w[0]); }

/* =========================================================================================
 * decimal helpers
 * ========================================================================================= */
static i128 pow10_i128(int e) { i128 r = 1; for (int i = 0; i < e; i++) r *= 10; return r; }
#define I128_MAX ((i128)(((u128)1 << 127) - 1))
#define I128_MIN (-I128_MAX - 1)

/* 10^p table (arrow MAX_DECIMAL128_FOR_EACH_PRECISION is 10^p - 1): built once */
static i128 POW10_TAB[39];
static int pow10_ready = 0;
static void pow10_init(void) {
    if (pow10_ready) return;
    i128 r = 1;
    for (int i = 0; i <= 38; i++) { POW10_TAB[i] = r; if (i < 38) r *= 10; }
    pow10_ready = 1;
}
__attribute__((constructor)) static void co_init(void) { pow10_init(); }
static inline int valid_p(i128 v, int precision) { /* spark-expr/src/utils.rs:332-336 */
    i128 b = POW10_TAB[precision];
    return v < b && v > -b;
}
int co_is_valid_decimal_precision(i128 v, int precision) {
    if (precision > 38 || precision < 0) return 0;
    if (precision == 0) return v == 0; /* arrow table entry 0 is 0 */
    return valid_p(v, precision);
}
static int rowvalid(const uint8_t *v, int64_t i) { return v == NULL || v[i]; }

/* wide_decimal_binary_expr.rs:179-291 + check_overflow_and_convert :335-350 +
 * null_if_overflow_precision (non-ANSI). try_binary applies the op only where both sides are valid. */
int co_wide_decimal(int op, int64_t n, const i128 *l, const uint8_t *lv, int s1, const i128 *r,
                    const uint8_t *rv, int s2, int p_out, int s_out, int eval_mode, i128 *out,
                    uint8_t *outv) {
    i256 bound = i256_sub(i256_pow10((unsigned)p_out), (i256){{1, 0, 0, 0}});
    i256 neg_bound = i256_neg(bound);
    int scale_diff;
    i256 l_up = {{1, 0, 0, 0}}, r_up = {{1, 0, 0, 0}};
    if (op == 2) {
        scale_diff = (s1 + s2) - s_out;
    } else {
        int max_scale = s1 > s2 ? s1 : s2;
        l_up = i256_pow10((unsigned)(max_scale - s1));
        r_up = i256_pow10((unsigned)(max_scale - s2));
        scale_diff = max_scale - s_out;
    }
    for (int64_t i = 0; i < n; i++) {
        if (!rowvalid(lv, i) || !rowvalid(rv, i)) { out[i] = 0; outv[i] = 0; continue; }
        i256 raw;
        if (op == 2) raw = i256_mul(i256_from_i128(l[i]), i256_from_i128(r[i]));
        else {
            i256 a = i256_mul(i256_from_i128(l[i]), l_up), b = i256_mul(i256_from_i128(r[i]), r_up);
            raw = op == 0 ? i256_add(a, b) : i256_sub(a, b);
        }
        i256 res = raw;
        if (scale_diff > 0) res = i256_div_pow10_half_up(raw, (unsigned)scale_diff);
        else if (scale_diff < 0) res = i256_mul(raw, i256_pow10((unsigned)(-scale_diff)));
        if (i256_cmp(res, bound) > 0 || i256_cmp(res, neg_bound) < 0) {
            if (eval_mode == CO_ANSI) return CO_ERR_ARITHMETIC_OVERFLOW;
            out[i] = 0; outv[i] = 0; /* i128::MAX sentinel -> null_if_overflow_precision */
        } else {
            i128 v = i256_to_i128(res);
            /* non-ANSI: null_if_overflow_precision(p_out) -- already inside bound */
            out[i] = v; outv[i] = 1;
        }
    }
    return CO_OK;
}

/* arrow-arith 58.4.0 `decimal_op` (3P; restated from the Arrow decimal arithmetic rules):
 *  add/sub: result scale = max(s1,s2), precision = min(max(s1,s2)+max(p1-s1,p2-s2)+1, 38);
 *           value = l*10^(rs-s1) +/- r*10^(rs-s2), checked
 *  mul:     result scale = s1+s2, precision = min(p1+p2+1,38); value = l*r checked           */
int co_plain_decimal(int op, int64_t n, const i128 *l, const uint8_t *lv, int p1, int s1,
                     const i128 *r, const uint8_t *rv, int p2, int s2, i128 *out, uint8_t *outv,
                     int *p_res, int *s_res) {
    int rs, rp;
    if (op == 2) { rs = s1 + s2; rp = p1 + p2 + 1; }
    else {
        rs = s1 > s2 ? s1 : s2;
        int a = p1 - s1, b = p2 - s2;
        rp = rs + (a > b ? a : b) + 1;
    }
    if (rp > 38) rp = 38;
    if (p_res) *p_res = rp;
    if (s_res) *s_res = rs;
    i128 lm = op == 2 ? 1 : pow10_i128(rs - s1), rm = op == 2 ? 1 : pow10_i128(rs - s2);
    for (int64_t i = 0; i < n; i++) {
        if (!rowvalid(lv, i) || !rowvalid(rv, i)) { out[i] = 0; outv[i] = 0; continue; }
        i128 res;
        if (op == 2) {
            if (__builtin_mul_overflow(l[i], r[i], &res)) return CO_ERR_ARITHMETIC_OVERFLOW;
        } else {
            i128 a, b;
            if (__builtin_mul_overflow(l[i], lm, &a)) return CO_ERR_ARITHMETIC_OVERFLOW;
            if (__builtin_mul_overflow(r[i], rm, &b)) return CO_ERR_ARITHMETIC_OVERFLOW;
            if (op == 0 ? __builtin_add_overflow(a, b, &res) : __builtin_sub_overflow(a, b, &res))
                return CO_ERR_ARITHMETIC_OVERFLOW;
        }
        out[i] = res; outv[i] = 1;
    }
    return CO_OK;
}

/* checkoverflow.rs:105-200: bound check only, no rescale; overflow -> NULL (non-ANSI) / error. */
int co_check_overflow(int64_t n, const i128 *in, const uint8_t *inv, int precision,
                      int fail_on_error, i128 *out, uint8_t *outv) {
    for (int64_t i = 0; i < n; i++) {
        if (!rowvalid(inv, i)) { out[i] = 0; outv[i] = 0; continue; }
        if (co_is_valid_decimal_precision(in[i], precision)) { out[i] = in[i]; outv[i] = 1; }
        else {
            if (fail_on_error) return CO_ERR_ARITHMETIC_OVERFLOW;
            out[i] = 0; outv[i] = 0;
        }
    }
    return CO_OK;
}

/* decimal_rescale_check.rs:111-150 rescale_and_check */
int co_decimal_rescale_check(int64_t n, const i128 *in, const uint8_t *inv, int s_in, int p_out,
                             int s_out, int fail_on_error, i128 *out, uint8_t *outv) {
    int delta = s_out - s_in;
    int ad = delta < 0 ? -delta : delta;
    if (ad > 38) return CO_ERR_INVALID;
    i128 factor = pow10_i128(ad), bound = pow10_i128(p_out) - 1;
    for (int64_t i = 0; i < n; i++) {
        if (!rowvalid(inv, i)) { out[i] = 0; outv[i] = 0; continue; }
        i128 v = in[i], res;
        int ovf = 0;
        if (delta > 0) { if (__builtin_mul_overflow(v, factor, &res)) ovf = 1; }
        else if (delta < 0) {
            i128 half = factor / 2, sign = (v > 0) - (v < 0);
            res = (v + sign * half) / factor;
        } else res = v;
        if (!ovf) { i128 a = res < 0 ? -res : res; if (a > bound) ovf = 1; }
        if (ovf) { if (fail_on_error) return CO_ERR_ARITHMETIC_OVERFLOW; out[i] = 0; outv[i] = 0; }
        else { out[i] = res; outv[i] = 1; }
    }
    return CO_OK;
}

/* checked_arithmetic.rs:53-128 (TRY -> NULL on overflow, ANSI -> error);
 * LEGACY: DataFusion BinaryExpr -> arrow-arith wrapping ops (planner.rs:1126). */
int co_int_arith(int op, int width, int64_t n, const int64_t *l, const uint8_t *lv,
                 const int64_t *r, const uint8_t *rv, int eval_mode, int64_t *out, uint8_t *outv) {
    int64_t lo = width == 64 ? INT64_MIN : -((int64_t)1 << (width - 1));
    int64_t hi = width == 64 ? INT64_MAX : ((int64_t)1 << (width - 1)) - 1;
    for (int64_t i = 0; i < n; i++) {
        if (!rowvalid(lv, i) || !rowvalid(rv, i)) { out[i] = 0; outv[i] = 0; continue; }
        i128 x = op == 0 ? (i128)l[i] + r[i] : op == 1 ? (i128)l[i] - r[i] : (i128)l[i] * r[i];
        int ovf = x < lo || x > hi;
        if (ovf && eval_mode == CO_ANSI) return CO_ERR_ARITHMETIC_OVERFLOW;
        if (ovf && eval_mode == CO_TRY) { out[i] = 0; outv[i] = 0; continue; }
        /* wrap to width */
        uint64_t u = (uint64_t)(u128)x;
        if (width < 64) {
            u &= ((uint64_t)1 << width) - 1;
            if (u >> (width - 1)) u |= ~(((uint64_t)1 << width) - 1);
        }
        out[i] = (int64_t)u; outv[i] = 1;
    }
    return CO_OK;
}

/* =========================================================================================
 * murmur3 (Spark variant)  -- spark-expr/src/hash_funcs/murmur3.rs:73-137
 * ========================================================================================= */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint32_t mix_k1(uint32_t k1) { k1 *= 0xcc9e2d51u; k1 = rotl32(k1, 15); k1 *= 0x1b873593u; return k1; }
static inline uint32_t mix_h1(uint32_t h1, uint32_t k1) { h1 ^= k1; h1 = rotl32(h1, 13); return h1 * 5u + 0xe6546b64u; }
static inline uint32_t fmix(uint32_t h1, uint32_t len) {
    h1 ^= len; h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
    return h1;
}
uint32_t co_murmur3_bytes(const uint8_t *data, int64_t len, uint32_t seed) {
    int64_t aligned = len - len % 4;
    uint32_t h1 = seed;
    for (int64_t i = 0; i < aligned; i += 4) {
        uint32_t w; memcpy(&w, data + i, 4); /* little-endian host */
        h1 = mix_h1(h1, mix_k1(w));
    }
    for (int64_t i = aligned; i < len; i++) { /* tail bytes are sign-extended (murmur3.rs:131) */
        uint32_t w = (uint32_t)(int32_t)(int8_t)data[i];
        h1 = mix_h1(h1, mix_k1(w));
    }
    return fmix(h1, (uint32_t)len);
}

int co_murmur3_column(int kind, int64_t n, const void *values, const uint8_t *valid,
                      uint32_t *hashes) {
    for (int64_t i = 0; i < n; i++) {
        if (!rowvalid(valid, i)) continue; /* NULL leaves the running hash unchanged (utils.rs:38-42) */
        uint8_t buf[16];
        int len;
        switch (kind) {
        case 0: { int32_t x = ((const uint8_t *)values)[i] ? 1 : 0; memcpy(buf, &x, 4); len = 4; break; }
        case 1: { int32_t x = ((const int8_t *)values)[i]; memcpy(buf, &x, 4); len = 4; break; }
        case 2: { int32_t x = ((const int16_t *)values)[i]; memcpy(buf, &x, 4); len = 4; break; }
        case 3: case 7: { int32_t x = ((const int32_t *)values)[i]; memcpy(buf, &x, 4); len = 4; break; }
        case 4: case 8: { int64_t x = ((const int64_t *)values)[i]; memcpy(buf, &x, 8); len = 8; break; }
        case 5: { float f = ((const float *)values)[i];
                  if (f == 0.0f && signbit(f)) { int32_t z = 0; memcpy(buf, &z, 4); } else memcpy(buf, &f, 4);
                  len = 4; break; }
        case 6: { double d = ((const double *)values)[i];
                  if (d == 0.0 && signbit(d)) { int64_t z = 0; memcpy(buf, &z, 8); } else memcpy(buf, &d, 8);
                  len = 8; break; }
        case 9: { i128 v = ((const i128 *)values)[i]; /* utils.rs:159-196: i64::try_from, error if it does not fit */
                  if (v > INT64_MAX || v < INT64_MIN) return CO_ERR_INVALID;
                  int64_t x = (int64_t)v; memcpy(buf, &x, 8); len = 8; break; }
        case 10: { i128 v = ((const i128 *)values)[i]; memcpy(buf, &v, 16); len = 16; break; } /* utils.rs:199-226 */
        default: return CO_ERR_INVALID;
        }
        hashes[i] = co_murmur3_bytes(buf, len, hashes[i]);
    }
    return CO_OK;
}

void co_murmur3_strings(int64_t n, const int32_t *offsets, const uint8_t *data,
                        const uint8_t *valid, uint32_t *hashes) {
    for (int64_t i = 0; i < n; i++) {
        if (!rowvalid(valid, i)) continue;
        hashes[i] = co_murmur3_bytes(data + offsets[i], offsets[i + 1] - offsets[i], hashes[i]);
    }
}

uint32_t co_pmod(uint32_t hash, uint32_t n) { /* comet_partitioning.rs:51-57 */
    int32_t h = (int32_t)hash, m = (int32_t)n;
    int32_t r = h % m;
    return (uint32_t)(r < 0 ? (r + m) % m : r);
}

void co_partition_rows(int64_t n, const uint32_t *hashes, uint32_t n_parts, uint32_t *pids,
                       int64_t *starts, int64_t *row_idx) { /* multi_partition.rs:54-99,298-310 */
    memset(starts, 0, sizeof(int64_t) * (n_parts + 1));
    for (int64_t i = 0; i < n; i++) { pids[i] = co_pmod(hashes[i], n_parts); starts[pids[i] + 1]++; }
    for (uint32_t p = 0; p < n_parts; p++) starts[p + 1] += starts[p];
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * n_parts);
    memcpy(cur, starts, sizeof(int64_t) * n_parts);
    for (int64_t i = 0; i < n; i++) row_idx[cur[pids[i]]++] = i;
    free(cur);
}

/* =========================================================================================
 * accumulators
 * ========================================================================================= */
static int keep(const uint8_t *valid, const uint8_t *filter, int64_t i) {
    if (filter && !filter[i]) return 0;
    return rowvalid(valid, i);
}

/* sum_decimal.rs:418-439 update_single */
static inline int sum_decimal_update_single(i128 value, i128 *sum, uint8_t *sum_valid, uint8_t *is_empty,
                                            int precision, int eval_mode) {
    if (!*is_empty && !*sum_valid) return CO_OK; /* sticky overflow */
    i128 running = *sum_valid ? *sum : 0, ns;
    int ovf = __builtin_add_overflow(running, value, &ns);
    if (ovf || !valid_p(ns, precision)) {
        if (eval_mode == CO_ANSI) return CO_ERR_ARITHMETIC_OVERFLOW;
        *sum_valid = 0; *sum = 0;
    } else { *sum = ns; *sum_valid = 1; }
    *is_empty = 0;
    return CO_OK;
}
/* avg_decimal.rs:483-495 update_single */
static inline void avg_decimal_update_single(i128 v, i128 *sum, int64_t *count, uint8_t *is_not_null, int sum_precision) {
    i128 ns;
    int ovf = __builtin_add_overflow(*sum, v, &ns);
    if (ovf) ns = (i128)((u128)*sum + (u128)v);
    *count += 1; *sum = ns;
    if (ovf || !valid_p(ns, sum_precision)) *is_not_null = 0;
}
int co_sum_decimal_update(int64_t n, const i128 *v, const uint8_t *valid, const uint8_t *filter,
                          const int64_t *g, i128 *sum, uint8_t *sum_valid, uint8_t *is_empty,
                          int precision, int eval_mode) {
    for (int64_t i = 0; i < n; i++) {
        if (!keep(valid, filter, i)) continue;
        int64_t k = g ? g[i] : 0;
        int e = sum_decimal_update_single(v[i], &sum[k], &sum_valid[k], &is_empty[k], precision, eval_mode);
        if (e) return e;
    }
    return CO_OK;
}
int co_sum_decimal_acc_update(int64_t n, const i128 *v, const uint8_t *valid, i128 *sum,
                              uint8_t *sum_valid, uint8_t *is_empty, int precision, int eval_mode) {
    /* sum_decimal.rs:238-269 update_batch */
    if (!*is_empty && !*sum_valid) return CO_OK;
    int64_t nulls = 0;
    for (int64_t i = 0; i < n; i++) nulls += !rowvalid(valid, i);
    *is_empty = *is_empty && (n == nulls);
    if (*is_empty) return CO_OK;
    for (int64_t i = 0; i < n; i++) {
        if (!rowvalid(valid, i)) continue;
        /* :200-225 -- same rule as the grouped update_single */
        uint8_t e = *is_empty;
        int r = sum_decimal_update_single(v[i], sum, sum_valid, &e, precision, eval_mode);
        *is_empty = e;
        if (r) return r;
    }
    return CO_OK;
}
int co_sum_decimal_merge(int64_t n, const i128 *ts, const uint8_t *tsv, const uint8_t *te,
                         const int64_t *g, i128 *sum, uint8_t *sum_valid, uint8_t *is_empty,
                         int precision, int eval_mode) { /* sum_decimal.rs:540-607 */
    for (int64_t i = 0; i < n; i++) {
        int64_t k = g ? g[i] : 0;
        int that_valid = rowvalid(tsv, i), that_empty = te[i];
        int that_ovf = !that_empty && !that_valid, this_ovf = !is_empty[k] && !sum_valid[k];
        if (that_ovf || this_ovf) { sum_valid[k] = 0; sum[k] = 0; is_empty[k] = 0; continue; }
        if (that_empty) continue;
        if (is_empty[k]) { sum[k] = that_valid ? ts[i] : 0; sum_valid[k] = (uint8_t)that_valid; is_empty[k] = 0; continue; }
        i128 ns;
        int ovf = __builtin_add_overflow(sum[k], ts[i], &ns);
        if (ovf || !co_is_valid_decimal_precision(ns, precision)) {
            if (eval_mode == CO_ANSI) return CO_ERR_ARITHMETIC_OVERFLOW;
            sum_valid[k] = 0; sum[k] = 0; is_empty[k] = 0;
        } else sum[k] = ns;
    }
    return CO_OK;
}
void co_sum_decimal_evaluate(int64_t ng, const i128 *sum, const uint8_t *sum_valid,
                             const uint8_t *is_empty, int precision, i128 *out, uint8_t *outv) {
    for (int64_t k = 0; k < ng; k++) { /* sum_decimal.rs:477-500 */
        if (is_empty[k] || !sum_valid[k] || !co_is_valid_decimal_precision(sum[k], precision)) { out[k] = 0; outv[k] = 0; }
        else { out[k] = sum[k]; outv[k] = 1; }
    }
}

int co_avg_decimal_update(int64_t n, const i128 *v, const uint8_t *valid, const uint8_t *filter,
                          const int64_t *g, i128 *sums, int64_t *counts, uint8_t *is_not_null,
                          int sum_precision) { /* avg_decimal.rs:483-495 */
    for (int64_t i = 0; i < n; i++) {
        if (!keep(valid, filter, i)) continue;
        int64_t k = g ? g[i] : 0;
        avg_decimal_update_single(v[i], &sums[k], &counts[k], &is_not_null[k], sum_precision); /* overflowing_add keeps the wrapped value */
    }
    return CO_OK;
}
int co_avg_decimal_merge(int64_t n, const i128 *ps, const uint8_t *psv, const int64_t *pc,
                         const uint8_t *pcv, const int64_t *g, i128 *sums, int64_t *counts,
                         uint8_t *is_not_null, int sum_precision, int eval_mode) { /* :542-595 */
    for (int64_t i = 0; i < n; i++) { int64_t k = g ? g[i] : 0; counts[k] += pc[i]; }
    for (int64_t i = 0; i < n; i++) {
        int64_t k = g ? g[i] : 0;
        if (!rowvalid(psv, i)) { is_not_null[k] = 0; continue; }
        i128 ns;
        int ovf = __builtin_add_overflow(sums[k], ps[i], &ns);
        if (ovf || !co_is_valid_decimal_precision(ns, sum_precision)) {
            if (eval_mode == CO_ANSI) return CO_ERR_ARITHMETIC_OVERFLOW;
            is_not_null[k] = 0;
        } else sums[k] = ns;
    }
    if (pcv) for (int64_t i = 0; i < n; i++) if (!pcv[i]) is_not_null[g ? g[i] : 0] = 0;
    return CO_OK;
}
/* avg_decimal.rs:670-689 */
static int avg_fn(i128 sum, i128 count, i128 tmin, i128 tmax, i128 scaler, i128 *out) {
    i128 value;
    if (__builtin_mul_overflow(sum, scaler, &value)) return 0;
    i128 div = value / count, rem = value % count;
    i128 half = count / 2 + (count % 2 != 0 && count > 0 ? 1 : 0); /* div_ceil(count, 2) */
    i128 half_neg = -half, nv = div;
    if (value >= 0) { if (rem >= half) nv = div + 1; }
    else { if (rem <= half_neg) nv = div - 1; }
    if (nv >= tmin && nv <= tmax) { *out = nv; return 1; }
    return 0;
}
int co_avg_decimal_evaluate(int64_t ng, const i128 *sums, const int64_t *counts,
                            const uint8_t *is_not_null, int sum_scale, int tp, int ts,
                            int eval_mode, i128 *out, uint8_t *outv) { /* :597-636 */
    int d = ts - sum_scale; if (d < 0) d = 0; /* saturating_sub */
    i128 scaler = pow10_i128(d), tmax = pow10_i128(tp) - 1, tmin = -tmax;
    for (int64_t k = 0; k < ng; k++) {
        if (!is_not_null[k] && counts[k] > 0 && eval_mode == CO_ANSI) return CO_ERR_ARITHMETIC_OVERFLOW;
        if (!is_not_null[k] || counts[k] == 0) { out[k] = 0; outv[k] = 0; continue; }
        i128 r;
        if (avg_fn(sums[k], (i128)counts[k], tmin, tmax, scaler, &r)) { out[k] = r; outv[k] = 1; }
        else { out[k] = 0; outv[k] = 0; }
    }
    return CO_OK;
}

void co_avg_f64_update(int64_t n, const double *v, const uint8_t *valid, const uint8_t *filter,
                       const int64_t *g, double *sums, int64_t *counts) { /* avg.rs:229-277 */
    for (int64_t i = 0; i < n; i++) {
        if (!keep(valid, filter, i)) continue;
        int64_t k = g ? g[i] : 0;
        sums[k] += v[i]; counts[k] += 1;
    }
}
void co_avg_f64_merge(int64_t n, const double *ps, const int64_t *pc, const int64_t *g,
                      double *sums, int64_t *counts) { /* avg.rs:279-309 */
    for (int64_t i = 0; i < n; i++) { int64_t k = g ? g[i] : 0; counts[k] += pc[i]; }
    for (int64_t i = 0; i < n; i++) { int64_t k = g ? g[i] : 0; sums[k] += ps[i]; }
}
void co_avg_f64_evaluate(int64_t ng, const double *sums, const int64_t *counts, double *out,
                         uint8_t *outv) { /* avg.rs:311-327 */
    for (int64_t k = 0; k < ng; k++) {
        if (counts[k] != 0) { out[k] = sums[k] / (double)counts[k]; outv[k] = 1; }
        else { out[k] = 0; outv[k] = 0; }
    }
}

int co_sum_int_update(int64_t n, const int64_t *v, const uint8_t *valid, const uint8_t *filter,
                      const int64_t *g, int64_t *sums, uint8_t *sums_valid, uint8_t *overflowed,
                      int eval_mode) { /* sum_int.rs:403-440 (Legacy), Ansi/Try analogues */
    for (int64_t i = 0; i < n; i++) {
        if (!keep(valid, filter, i)) continue;
        int64_t k = g ? g[i] : 0;
        if (eval_mode == CO_TRY && overflowed && overflowed[k]) continue;
        int64_t cur = sums_valid[k] ? sums[k] : 0, ns;
        int ovf = __builtin_add_overflow(cur, v[i], &ns);
        if (eval_mode == CO_LEGACY) { sums[k] = (int64_t)((uint64_t)cur + (uint64_t)v[i]); sums_valid[k] = 1; }
        else if (!ovf) { sums[k] = ns; sums_valid[k] = 1; }
        else if (eval_mode == CO_ANSI) return CO_ERR_ARITHMETIC_OVERFLOW;
        else { sums_valid[k] = 0; sums[k] = 0; if (overflowed) overflowed[k] = 1; }
    }
    return CO_OK;
}

void co_sum_f64_update(int64_t n, const double *v, const uint8_t *valid, const uint8_t *filter,
                       const int64_t *g, double *sums, uint8_t *sums_valid) {
    for (int64_t i = 0; i < n; i++) {
        if (!keep(valid, filter, i)) continue;
        int64_t k = g ? g[i] : 0;
        sums[k] = (sums_valid[k] ? sums[k] : 0.0) + v[i]; sums_valid[k] = 1;
    }
}
void co_count_update(int64_t n, const uint8_t *valid, const uint8_t *filter, const int64_t *g,
                     int64_t *counts) {
    for (int64_t i = 0; i < n; i++) if (keep(valid, filter, i)) counts[g ? g[i] : 0] += 1;
}

/* Exact sum of doubles via a 2176-bit fixed-point superaccumulator (one per group, built lazily).
 * Result is the correctly-rounded (round-to-nearest-even) double of the exact real sum of the
 * finite inputs.  Non-finite inputs fall back to plain addition semantics. */
#define SA_WORDS 70 /* 70*32 = 2240 bits >= 2098 needed for the double range + carries */
typedef struct { int64_t w[SA_WORDS]; int nonfinite; double nf; int64_t pending; } superacc;
static void sa_normalize(superacc *a) {
    int64_t c = 0;
    for (int i = 0; i < SA_WORDS; i++) {
        int64_t x = a->w[i] + c;
        c = x >> 32; /* arithmetic shift: floor division */
        a->w[i] = x - (c << 32);
    }
    /* top carry is the sign extension; fold back into the top word */
    a->w[SA_WORDS - 1] += c << 32;
    a->pending = 0;
}
static void sa_add(superacc *a, double x) {
    if (!isfinite(x)) { a->nonfinite = 1; a->nf += x; return; }
    if (x == 0.0) return;
    uint64_t bits; memcpy(&bits, &x, 8);
    int neg = (int)(bits >> 63);
    int e = (int)((bits >> 52) & 0x7ff);
    uint64_t m = bits & 0xfffffffffffffULL;
    if (e == 0) e = 1; else m |= 1ULL << 52;
    /* value = m * 2^(e-1075); bit position of LSB relative to 2^-1074 = e-1 */
    int pos = e - 1, word = pos >> 5, sh = pos & 31;
    u128 mm = (u128)m << sh; /* up to 53+31 = 84 bits -> spans 3 words */
    for (int k = 0; k < 3; k++) {
        int64_t part = (int64_t)(uint32_t)(mm >> (32 * k));
        a->w[word + k] += neg ? -part : part;
    }
    if (++a->pending >= (1LL << 29)) sa_normalize(a);
}
static double sa_round(superacc *a) {
    if (a->nonfinite) return a->nf;
    sa_normalize(a);
    /* sign */
    int neg = a->w[SA_WORDS - 1] < 0;
    int64_t w[SA_WORDS];
    memcpy(w, a->w, sizeof(w));
    if (neg) { /* negate: two's complement over base-2^32 digits */
        int64_t c = 0;
        for (int i = 0; i < SA_WORDS; i++) {
            int64_t x = -w[i] + c;
            c = x >> 32; w[i] = x - (c << 32);
        }
        w[SA_WORDS - 1] += c << 32;
    }
    int top = SA_WORDS - 1;
    while (top >= 0 && w[top] == 0) top--;
    if (top < 0) return 0.0;
    /* gather the top 128 bits starting at the highest set bit */
    int hb = 63 - __builtin_clzll((uint64_t)w[top]); /* highest set bit within word */
    int64_t msb = (int64_t)top * 32 + hb;            /* absolute bit index (LSB = 2^-1074) */
    /* extract 64 bits below and including msb, plus sticky */
    uint64_t mant = 0; int sticky = 0;
    for (int b = 0; b < 64; b++) {
        int64_t idx = msb - b;
        int bit = 0;
        if (idx >= 0) bit = (int)((w[idx >> 5] >> (idx & 31)) & 1);
        mant = (mant << 1) | (uint64_t)bit;
    }
    for (int64_t idx = msb - 64; idx >= 0 && !sticky; idx--)
        if ((w[idx >> 5] >> (idx & 31)) & 1) sticky = 1;
    /* mant has 64 significant bits (top bit set); value = mant * 2^(msb-63) * 2^-1074 */
    int64_t exp2 = msb - 63 - 1074; /* exponent of mant's LSB */
    /* round to 53 bits (or fewer for subnormals) */
    int drop = 11;
    int64_t lsb_exp = exp2 + drop;
    if (lsb_exp < -1074) { drop += (int)(-1074 - lsb_exp); lsb_exp = -1074; }
    double res;
    if (drop >= 64) { /* far below subnormal range: cannot happen since LSB >= 2^-1074 */
        res = 0.0;
    } else {
        uint64_t keepm = mant >> drop, rb = (mant >> (drop - 1)) & 1;
        uint64_t rest = mant & ((1ULL << (drop - 1)) - 1);
        if (rb && (rest || sticky || (keepm & 1))) keepm++;
        res = ldexp((double)keepm, (int)lsb_exp);
    }
    return neg ? -res : res;
}
void co_sum_f64_exact(int64_t n, const double *v, const uint8_t *valid, const uint8_t *filter,
                      const int64_t *g, int64_t ng, double *out) {
    superacc *acc = (superacc *)calloc((size_t)ng, sizeof(superacc));
    for (int64_t i = 0; i < n; i++) if (keep(valid, filter, i)) sa_add(&acc[g ? g[i] : 0], v[i]);
    for (int64_t k = 0; k < ng; k++) out[k] = sa_round(&acc[k]);
    free(acc);
}

/* =========================================================================================
 * whole-pipeline baselines.  Each OpenMP thread plays one Spark partition: Partial aggregate
 * over its contiguous row range with the GroupsAccumulator rules above, then one Final merge
 * (planner.rs:1262-1271 modes; shuffle in between carries the state columns).
 * ========================================================================================= */
#define MAXG 64
static int nthreads_or_default(int n) {
#ifdef _OPENMP
    return n > 0 ? n : omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}

/* Q1 expression tree on d(12,2) inputs: 1 - disc -> d(13,2) plain sub (+CheckOverflow);
 * price * (1-disc) -> d(26,4) plain mul (12+13 = 25 < 38) + CheckOverflow(26); * (1+tax) -> d(38,6)
 * WideDecimal mul (26+13 >= 38).  sums: qty d(22,2), price d(22,2), disc_price d(36,4), charge
 * d(38,6); avgs: sum state d(22,2), result d(16,6). */
static inline int q1_row_dec(i128 price, i128 disc, i128 tax, i128 *disc_price, int *dp_valid,
                             i128 *charge, int *ch_valid) {
    static const i128 one = 100;          /* Literal 1 (d(1,0)) brought to scale 2 by arrow-arith add/sub */
    i128 om = one - disc;                 /* d(13,2) plain sub + CheckOverflow(13) */
    int om_valid = valid_p(om, 13);
    i128 dp = 0; int dpv = 0;
    if (om_valid) { dp = price * om; dpv = valid_p(dp, 26); } /* |.| < 10^25 < 2^127 */
    i128 op = one + tax; int op_valid = valid_p(op, 13);
    *disc_price = dpv ? dp : 0; *dp_valid = dpv;
    if (dpv && op_valid) {
        /* WideDecimal mul d(26,4)*d(13,2)->d(38,6): natural scale == output scale, so the i256
         * product is only bound-checked.  An i128-overflowing product has |x| >= 2^127 > 10^38-1,
         * i.e. it is out of bound too, so the checked i128 multiply is an exact restatement. */
        i128 o;
        if (__builtin_mul_overflow(dp, op, &o) || !valid_p(o, 38)) { *charge = 0; *ch_valid = 0; }
        else { *charge = o; *ch_valid = 1; }
    } else { *charge = 0; *ch_valid = 0; }
    return 0;
}

typedef struct {
    i128 s_qty, s_base, s_dp, s_ch; uint8_t v_qty, v_base, v_dp, v_ch, e_qty, e_base, e_dp, e_ch;
    i128 a_qty, a_price, a_disc; int64_t c_qty, c_price, c_disc; uint8_t nn_qty, nn_price, nn_disc;
    int64_t count;
} q1_state;
static void q1_state_init(q1_state *s) {
    memset(s, 0, sizeof(*s));
    s->v_qty = s->v_base = s->v_dp = s->v_ch = 1; s->e_qty = s->e_base = s->e_dp = s->e_ch = 1;
    s->nn_qty = s->nn_price = s->nn_disc = 1;
}
int co_q1_dec(int64_t n, const i128 *qty, const i128 *price, const i128 *disc, const i128 *tax,
              const int32_t *shipdate, const uint8_t *rf, const uint8_t *ls, int n_rf, int n_ls,
              int32_t cutoff, int n_threads, co_q1_dec_row *out) {
    int ng = n_rf * n_ls;
    if (ng > MAXG) return CO_ERR_INVALID;
    int T = nthreads_or_default(n_threads);
    q1_state *part = (q1_state *)malloc(sizeof(q1_state) * (size_t)T * MAXG);
    for (int i = 0; i < T * MAXG; i++) q1_state_init(&part[i]);
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        int64_t lo = n * t / T, hi = n * (t + 1) / T;
        q1_state *st = part + (size_t)t * MAXG;
        for (int64_t i = lo; i < hi; i++) {
            if (!(shipdate[i] <= cutoff)) continue;
            q1_state *s = &st[rf[i] * n_ls + ls[i]];
            i128 dp, ch; int dpv, chv;
            q1_row_dec(price[i], disc[i], tax[i], &dp, &dpv, &ch, &chv);
            sum_decimal_update_single(qty[i], &s->s_qty, &s->v_qty, &s->e_qty, 22, CO_LEGACY);
            sum_decimal_update_single(price[i], &s->s_base, &s->v_base, &s->e_base, 22, CO_LEGACY);
            if (dpv) sum_decimal_update_single(dp, &s->s_dp, &s->v_dp, &s->e_dp, 36, CO_LEGACY);
            if (chv) sum_decimal_update_single(ch, &s->s_ch, &s->v_ch, &s->e_ch, 38, CO_LEGACY);
            avg_decimal_update_single(qty[i], &s->a_qty, &s->c_qty, &s->nn_qty, 22);
            avg_decimal_update_single(price[i], &s->a_price, &s->c_price, &s->nn_price, 22);
            avg_decimal_update_single(disc[i], &s->a_disc, &s->c_disc, &s->nn_disc, 22);
            s->count++;
        }
    }
    /* Final: merge partition states in partition order */
    for (int k = 0; k < ng; k++) {
        q1_state f; q1_state_init(&f);
        int64_t z = 0;
        for (int t = 0; t < T; t++) {
            q1_state *p = &part[(size_t)t * MAXG + k];
            if (p->count == 0) continue; /* group absent from this partition's output */
            co_sum_decimal_merge(1, &p->s_qty, &p->v_qty, &p->e_qty, &z, &f.s_qty, &f.v_qty, &f.e_qty, 22, CO_LEGACY);
            co_sum_decimal_merge(1, &p->s_base, &p->v_base, &p->e_base, &z, &f.s_base, &f.v_base, &f.e_base, 22, CO_LEGACY);
            co_sum_decimal_merge(1, &p->s_dp, &p->v_dp, &p->e_dp, &z, &f.s_dp, &f.v_dp, &f.e_dp, 36, CO_LEGACY);
            co_sum_decimal_merge(1, &p->s_ch, &p->v_ch, &p->e_ch, &z, &f.s_ch, &f.v_ch, &f.e_ch, 38, CO_LEGACY);
            /* avg state arrays share one null buffer = is_not_null (avg_decimal.rs:640-656) */
            co_avg_decimal_merge(1, &p->a_qty, &p->nn_qty, &p->c_qty, &p->nn_qty, &z, &f.a_qty, &f.c_qty, &f.nn_qty, 22, CO_LEGACY);
            co_avg_decimal_merge(1, &p->a_price, &p->nn_price, &p->c_price, &p->nn_price, &z, &f.a_price, &f.c_price, &f.nn_price, 22, CO_LEGACY);
            co_avg_decimal_merge(1, &p->a_disc, &p->nn_disc, &p->c_disc, &p->nn_disc, &z, &f.a_disc, &f.c_disc, &f.nn_disc, 22, CO_LEGACY);
            f.count += p->count;
        }
        co_q1_dec_row *o = &out[k];
        memset(o, 0, sizeof(*o));
        o->present = f.count > 0;
        co_sum_decimal_evaluate(1, &f.s_qty, &f.v_qty, &f.e_qty, 22, &o->sum_qty, &o->v_sum_qty);
        co_sum_decimal_evaluate(1, &f.s_base, &f.v_base, &f.e_base, 22, &o->sum_base, &o->v_sum_base);
        co_sum_decimal_evaluate(1, &f.s_dp, &f.v_dp, &f.e_dp, 36, &o->sum_disc_price, &o->v_sum_disc_price);
        co_sum_decimal_evaluate(1, &f.s_ch, &f.v_ch, &f.e_ch, 38, &o->sum_charge, &o->v_sum_charge);
        co_avg_decimal_evaluate(1, &f.a_qty, &f.c_qty, &f.nn_qty, 2, 16, 6, CO_LEGACY, &o->avg_qty, &o->v_avg_qty);
        co_avg_decimal_evaluate(1, &f.a_price, &f.c_price, &f.nn_price, 2, 16, 6, CO_LEGACY, &o->avg_price, &o->v_avg_price);
        co_avg_decimal_evaluate(1, &f.a_disc, &f.c_disc, &f.nn_disc, 2, 16, 6, CO_LEGACY, &o->avg_disc, &o->v_avg_disc);
        o->count = f.count;
    }
    free(part);
    return CO_OK;
}

typedef struct { double s[7]; int64_t c[4]; uint8_t sv[4]; } q1f_state; /* s: qty,base,dp,ch,aq,ap,ad */
int co_q1_f64(int64_t n, const double *qty, const double *price, const double *disc,
              const double *tax, const int32_t *shipdate, const uint8_t *rf, const uint8_t *ls,
              int n_rf, int n_ls, int32_t cutoff, int n_threads, co_q1_f64_row *out) {
    int ng = n_rf * n_ls;
    if (ng > MAXG) return CO_ERR_INVALID;
    int T = nthreads_or_default(n_threads);
    q1f_state *part = (q1f_state *)calloc((size_t)T * MAXG, sizeof(q1f_state));
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        int64_t lo = n * t / T, hi = n * (t + 1) / T;
        q1f_state *st = part + (size_t)t * MAXG;
        for (int64_t i = lo; i < hi; i++) {
            if (!(shipdate[i] <= cutoff)) continue;
            q1f_state *s = &st[rf[i] * n_ls + ls[i]];
            double dp = price[i] * (1.0 - disc[i]), ch = dp * (1.0 + tax[i]);
            s->s[0] += qty[i]; s->s[1] += price[i]; s->s[2] += dp; s->s[3] += ch;
            s->s[4] += qty[i]; s->s[5] += price[i]; s->s[6] += disc[i];
            s->c[0]++;
        }
    }
    for (int k = 0; k < ng; k++) {
        double s[7] = {0}; int64_t c = 0;
        for (int t = 0; t < T; t++) {
            q1f_state *p = &part[(size_t)t * MAXG + k];
            if (!p->c[0]) continue;
            for (int j = 0; j < 7; j++) s[j] += p->s[j];
            c += p->c[0];
        }
        co_q1_f64_row *o = &out[k];
        memset(o, 0, sizeof(*o));
        o->present = c > 0; o->count = c;
        o->sum_qty = s[0]; o->sum_base = s[1]; o->sum_disc_price = s[2]; o->sum_charge = s[3];
        if (c) { o->avg_qty = s[4] / (double)c; o->avg_price = s[5] / (double)c; o->avg_disc = s[6] / (double)c; }
    }
    free(part);
    return CO_OK;
}

int co_q6_dec(int64_t n, const i128 *qty, const i128 *price, const i128 *disc,
              const int32_t *shipdate, int32_t dlo, int32_t dhi, const i128 *p_disc_lo,
              const i128 *p_disc_hi, const i128 *p_qty_max, int n_threads, i128 *out, uint8_t *out_valid) {
    const i128 disc_lo = *p_disc_lo, disc_hi = *p_disc_hi, qty_max = *p_qty_max;
    int T = nthreads_or_default(n_threads);
    i128 *ps = (i128 *)calloc((size_t)T, sizeof(i128));
    uint8_t *pv = (uint8_t *)malloc((size_t)T), *pe = (uint8_t *)malloc((size_t)T);
    memset(pv, 1, (size_t)T); memset(pe, 1, (size_t)T);
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        int64_t lo = n * t / T, hi = n * (t + 1) / T;
        i128 s = 0; uint8_t v = 1, e = 1;
        for (int64_t i = lo; i < hi; i++) {
            if (!(shipdate[i] >= dlo && shipdate[i] < dhi && disc[i] >= disc_lo && disc[i] <= disc_hi && qty[i] < qty_max)) continue;
            i128 rev = price[i] * disc[i]; /* d(25,4) plain mul, CheckOverflow(25) */
            if (!valid_p(rev, 25)) continue; /* NULL input to sum is skipped */
            sum_decimal_update_single(rev, &s, &v, &e, 35, CO_LEGACY);
        }
        ps[t] = s; pv[t] = v; pe[t] = e;
    }
    i128 fs = 0; uint8_t fv = 1, fe = 1;
    co_sum_decimal_merge(T, ps, pv, pe, NULL, &fs, &fv, &fe, 35, CO_LEGACY);
    co_sum_decimal_evaluate(1, &fs, &fv, &fe, 35, out, out_valid);
    free(ps); free(pv); free(pe);
    return CO_OK;
}

int co_q6_f64(int64_t n, const double *qty, const double *price, const double *disc,
              const int32_t *shipdate, int32_t dlo, int32_t dhi, double disc_lo, double disc_hi,
              double qty_max, int n_threads, double *out, uint8_t *out_valid) {
    int T = nthreads_or_default(n_threads);
    double *ps = (double *)calloc((size_t)T, sizeof(double));
    uint8_t *pv = (uint8_t *)calloc((size_t)T, 1);
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        int64_t lo = n * t / T, hi = n * (t + 1) / T;
        double s = 0; uint8_t v = 0;
        for (int64_t i = lo; i < hi; i++) {
            if (!(shipdate[i] >= dlo && shipdate[i] < dhi && disc[i] >= disc_lo && disc[i] <= disc_hi && qty[i] < qty_max)) continue;
            s += price[i] * disc[i]; v = 1;
        }
        ps[t] = s; pv[t] = v;
    }
    double s = 0; uint8_t v = 0;
    for (int t = 0; t < T; t++) if (pv[t]) { s += ps[t]; v = 1; }
    *out = s; *out_valid = v;
    free(ps); free(pv);
    return CO_OK;
}

int64_t co_filter_project_dec(int64_t n, const i128 *qty, const i128 *price,
                              const int32_t *shipdate, int32_t cutoff, int n_threads, i128 *out,
                              uint8_t *outv) {
    int T = nthreads_or_default(n_threads);
    int64_t *cnt = (int64_t *)calloc((size_t)T + 1, sizeof(int64_t));
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        int64_t lo = n * t / T, hi = n * (t + 1) / T, c = 0;
        for (int64_t i = lo; i < hi; i++) c += shipdate[i] < cutoff;
        cnt[t + 1] = c;
#pragma omp barrier
#pragma omp single
        for (int k = 0; k < T; k++) cnt[k + 1] += cnt[k];
        int64_t o = cnt[t];
        for (int64_t i = lo; i < hi; i++) {
            if (!(shipdate[i] < cutoff)) continue;
            i128 v = qty[i] * price[i]; /* d(25,4) plain mul + CheckOverflow(25) */
            int ok = valid_p(v, 25);
            out[o] = ok ? v : 0; if (outv) outv[o] = (uint8_t)ok; o++;
        }
    }
    int64_t total = cnt[T];
    free(cnt);
    return total;
}

int64_t co_filter_project_f64(int64_t n, const double *qty, const double *price,
                              const int32_t *shipdate, int32_t cutoff, int n_threads, double *out) {
    int T = nthreads_or_default(n_threads);
    int64_t *cnt = (int64_t *)calloc((size_t)T + 1, sizeof(int64_t));
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        int64_t lo = n * t / T, hi = n * (t + 1) / T, c = 0;
        for (int64_t i = lo; i < hi; i++) c += shipdate[i] < cutoff;
        cnt[t + 1] = c;
#pragma omp barrier
#pragma omp single
        for (int k = 0; k < T; k++) cnt[k + 1] += cnt[k];
        int64_t o = cnt[t];
        for (int64_t i = lo; i < hi; i++) if (shipdate[i] < cutoff) out[o++] = qty[i] * price[i];
    }
    int64_t total = cnt[T];
    free(cnt);
    return total;
}

/* NUMA placement helper for the timed baseline (no reference equivalent: DataFusion reads each partition on the core that
 * produced it).  Copies `n` elements with the SAME static partition co_q1_dec uses, so every page of `dst` is first touched --
 * and therefore placed -- on the socket of the thread that will later scan it. */
void co_parallel_copy(void *dst, const void *src, int64_t n, int elem_bytes, int n_threads) {
    int T = nthreads_or_default(n_threads);
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        int64_t lo = n * t / T, hi = n * (t + 1) / T;
        memcpy((char *)dst + lo * elem_bytes, (const char *)src + lo * elem_bytes, (size_t)(hi - lo) * (size_t)elem_bytes);
    }
}

/* =========================================================================================
 * High-cardinality GROUP BY key SUM(value) (BASELINE configs[3]): Partial hash aggregate per partition
 * (= OpenMP thread, contiguous rows) with the SumDecimal rules, hash partitioning of the state rows by
 * pmod(murmur3(key, 42), T) exactly as the reference's shuffle assigns them (multi_partition.rs:298-310),
 * Final merge per partition.  Open addressing over (key, state); 3P AggregateExec's own table layout is
 * not restated -- only what it computes.
 * out_keys / out_sum / out_valid: one row per group, partition by partition (capacity n).  Returns the
 * number of groups, < 0 on error.
 * ========================================================================================= */
typedef struct { i128 sum; int64_t key; uint8_t used, valid, empty; } gb_slot;
typedef struct { gb_slot *tab; uint64_t mask; int64_t n; } gb_table;
static inline uint64_t gb_mix(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33; return h; }
static inline gb_slot *gb_probe(gb_slot *tab, uint64_t mask, int64_t key) {
    uint64_t s = gb_mix((uint64_t)key) & mask;
    while (tab[s].used && tab[s].key != key) s = (s + 1) & mask;
    return &tab[s];
}
static int gb_init(gb_table *t, uint64_t cap) { t->tab = (gb_slot *)calloc(cap, sizeof(gb_slot)); t->mask = cap - 1; t->n = 0; return t->tab != NULL; }
static int gb_grow(gb_table *t) { /* double at load 1/2 */
    gb_table b;
    if (!gb_init(&b, (t->mask + 1) * 2)) return 0;
    for (uint64_t k = 0; k <= t->mask; k++) if (t->tab[k].used) *gb_probe(b.tab, b.mask, t->tab[k].key) = t->tab[k];
    b.n = t->n;
    free(t->tab);
    *t = b;
    return 1;
}
static inline gb_slot *gb_upsert(gb_table *t, int64_t key) {
    gb_slot *s = gb_probe(t->tab, t->mask, key);
    if (s->used) return s;
    if ((uint64_t)(t->n + 1) * 2 > t->mask + 1) { if (!gb_grow(t)) return NULL; s = gb_probe(t->tab, t->mask, key); }
    s->used = 1; s->key = key; s->sum = 0; s->valid = 1; s->empty = 1; t->n++;
    return s;
}
typedef struct { int64_t key; i128 sum; uint8_t valid, empty; } gb_row; /* one state row of the shuffle */
int64_t co_groupby_sum_dec(int64_t n, const int64_t *keys, const i128 *vals, int precision, int n_threads,
                           int64_t *out_keys, i128 *out_sum, uint8_t *out_valid) {
    int T = nthreads_or_default(n_threads);
    gb_row **rows = (gb_row **)calloc((size_t)T, sizeof(gb_row *));         /* map output of task t, ordered by destination */
    int64_t *starts = (int64_t *)calloc((size_t)T * (size_t)(T + 1), sizeof(int64_t));
    gb_table *ftabs = (gb_table *)calloc((size_t)T, sizeof(gb_table));
    int64_t *gcount = (int64_t *)calloc((size_t)T + 1, sizeof(int64_t));
    int failed = 0;
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        /* ---- Partial over this partition's rows ---- */
        int64_t lo = n * t / T, hi = n * (t + 1) / T;
        gb_table tb;
        int ok = gb_init(&tb, 1 << 16);
        for (int64_t i = lo; ok && i < hi; i++) {
            gb_slot *s = gb_upsert(&tb, keys[i]);
            if (!s) { ok = 0; break; }
            sum_decimal_update_single(vals[i], &s->sum, &s->valid, &s->empty, precision, CO_LEGACY);
        }
        /* ---- ShuffleWriter: state rows bucketed by pmod(murmur3(key, 42), T), stable ---- */
        int64_t *st = starts + (size_t)t * (size_t)(T + 1);
        if (ok) {
            uint32_t *pid = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(tb.mask + 1));
            rows[t] = (gb_row *)malloc(sizeof(gb_row) * (size_t)(tb.n > 0 ? tb.n : 1));
            ok = pid && rows[t];
            if (ok) {
                for (uint64_t k = 0; k <= tb.mask; k++) if (tb.tab[k].used) {
                    uint32_t h = 42; co_murmur3_column(4, 1, &tb.tab[k].key, NULL, &h);
                    pid[k] = co_pmod(h, (uint32_t)T); st[pid[k] + 1]++;
                }
                for (int p = 0; p < T; p++) st[p + 1] += st[p];
                int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)T);
                for (int p = 0; p < T; p++) cur[p] = st[p];
                for (uint64_t k = 0; k <= tb.mask; k++) if (tb.tab[k].used) {
                    gb_row *r = &rows[t][cur[pid[k]]++];
                    r->key = tb.tab[k].key; r->sum = tb.tab[k].sum; r->valid = tb.tab[k].valid; r->empty = tb.tab[k].empty;
                }
                free(cur);
            }
            free(pid);
        }
        free(tb.tab);
        if (!ok) {
#pragma omp atomic write
            failed = 1;
        }
#pragma omp barrier
        /* ---- Final: partition t merges the rows every map task holds for it, in map-task order ---- */
        if (!failed) {
            int64_t incoming = 0;
            for (int m = 0; m < T; m++) incoming += starts[(size_t)m * (size_t)(T + 1) + t + 1] - starts[(size_t)m * (size_t)(T + 1) + t];
            uint64_t cap = 1 << 10;
            while (cap < (uint64_t)incoming * 2 + 2) cap <<= 1;
            if (gb_init(&ftabs[t], cap)) {
                for (int m = 0; m < T; m++) {
                    const int64_t *ms = starts + (size_t)m * (size_t)(T + 1);
                    for (int64_t k = ms[t]; k < ms[t + 1]; k++) {
                        const gb_row *r = &rows[m][k];
                        gb_slot *s = gb_upsert(&ftabs[t], r->key);
                        int64_t z = 0;
                        co_sum_decimal_merge(1, &r->sum, &r->valid, &r->empty, &z, &s->sum, &s->valid, &s->empty, precision, CO_LEGACY);
                    }
                }
                gcount[t + 1] = ftabs[t].n;
            } else {
#pragma omp atomic write
                failed = 1;
            }
        }
    }
    int64_t total = -1;
    if (!failed) {
        for (int p = 0; p < T; p++) gcount[p + 1] += gcount[p];
        total = gcount[T];
        if (out_keys) {
#pragma omp parallel num_threads(T)
            {
#ifdef _OPENMP
                int p = omp_get_thread_num();
#else
                int p = 0;
#endif
                int64_t o = gcount[p];
                const gb_table *ft = &ftabs[p];
                for (uint64_t k = 0; k <= ft->mask; k++) if (ft->tab[k].used) {
                    i128 r = 0; uint8_t rv = 0;
                    co_sum_decimal_evaluate(1, &ft->tab[k].sum, &ft->tab[k].valid, &ft->tab[k].empty, precision, &r, &rv);
                    out_keys[o] = ft->tab[k].key; out_sum[o] = r; out_valid[o] = rv; o++;
                }
            }
        }
    }
    for (int t = 0; t < T; t++) { free(rows[t]); free(ftabs[t].tab); }
    free(rows); free(starts); free(ftabs); free(gcount);
    return total;
}
