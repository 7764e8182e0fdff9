/*
 * comet_oracle.h -- CPU restatement of apache/datafusion-comet's hot-path semantics.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
 * library, and there only as the checker / the CPU baseline.  The product (libcomet_b200.so)
 * never links or dlopens it.
 *
 * Parity status: the in-tree Spark-semantics layer (decimal arithmetic, accumulators, murmur3,
 * pmod) is PINNED against the reference's own in-file known-answer tests (tests/test_oracle_kat.py
 * lists each vector with its reference file:line).  The third-party layers (parquet 58.4.0 page
 * decode, arrow-rs 58.4.0 plain arithmetic / compare / filter, DataFusion 54.1.0 FilterExec /
 * ProjectionExec / AggregateExec, f64 `sum`) have no source and no stored vectors under the
 * reference tree: those parts are "parity unpinned" at kernel granularity and are restated from the
 * Arrow / Parquet format specifications + SQL semantics, cross-checked against pyarrow (Arrow C++).
 *
 * Layout conventions: values are plain C arrays; validity is one byte per row (1 = valid), NULL
 * pointer = all valid.  Decimal128 = little-endian two's-complement __int128 unscaled value.
 * eval_mode: 0 = LEGACY, 1 = TRY, 2 = ANSI  (native/proto/src/proto/expr.proto:324 EvalMode).
 * Functions returning int return 0 on success, CO_ERR_* on a Spark-visible error.
 */
#ifndef COMET_ORACLE_H
#define COMET_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef __int128 co_i128;

#define CO_OK 0
#define CO_ERR_ARITHMETIC_OVERFLOW 1   /* ANSI arithmetic / decimal overflow */
#define CO_ERR_DIVIDE_BY_ZERO 2
#define CO_ERR_INVALID 3

#define CO_LEGACY 0
#define CO_TRY 1
#define CO_ANSI 2

/* ---- decimal elementwise ------------------------------------------------------------------ */
/* spark-expr/src/math_funcs/wide_decimal_binary_expr.rs:179-291 (op 0 add, 1 sub, 2 mul). */
int co_wide_decimal(int op, int64_t n, const co_i128 *l, const uint8_t *lv, int s1,
                    const co_i128 *r, const uint8_t *rv, int s2, int p_out, int s_out,
                    int eval_mode, co_i128 *out, uint8_t *outv);
/* arrow-arith 58.4.0 decimal add/sub/mul as used by DataFusion BinaryExpr (planner.rs:1126):
 * result scale max(s1,s2) (add/sub) or s1+s2 (mul); checked i128 arithmetic (error on i128
 * overflow).  Result precision/scale written to p_res/s_res. */
int co_plain_decimal(int op, int64_t n, const co_i128 *l, const uint8_t *lv, int p1, int s1,
                     const co_i128 *r, const uint8_t *rv, int p2, int s2, co_i128 *out,
                     uint8_t *outv, int *p_res, int *s_res);
/* spark-expr/src/math_funcs/internal/checkoverflow.rs:105-200 */
int co_check_overflow(int64_t n, const co_i128 *in, const uint8_t *inv, int precision,
                      int fail_on_error, co_i128 *out, uint8_t *outv);
/* spark-expr/src/math_funcs/internal/decimal_rescale_check.rs:111-150,171-240 */
int co_decimal_rescale_check(int64_t n, const co_i128 *in, const uint8_t *inv, int s_in,
                             int p_out, int s_out, int fail_on_error, co_i128 *out,
                             uint8_t *outv);
/* spark-expr/src/utils.rs:332-336 */
int co_is_valid_decimal_precision(co_i128 v, int precision);
/* spark-expr/src/math_funcs/checked_arithmetic.rs:53-128 ; op 0 add 1 sub 2 mul; width 8/16/32/64.
 * eval_mode LEGACY = wrapping (arrow-arith add_wrapping via BinaryExpr). */
int co_int_arith(int op, int width, int64_t n, const int64_t *l, const uint8_t *lv,
                 const int64_t *r, const uint8_t *rv, int eval_mode, int64_t *out, uint8_t *outv);

/* ---- hashing / partitioning ---------------------------------------------------------------- */
/* spark-expr/src/hash_funcs/murmur3.rs:73-137 */
uint32_t co_murmur3_bytes(const uint8_t *data, int64_t len, uint32_t seed);
/* hash_funcs/utils.rs:573-735 per-type rules; hashes[] updated in place (NULL rows unchanged).
 * kind: 0 bool(as i32) 1 i8 2 i16 3 i32 4 i64 5 f32 6 f64 7 date32 8 timestamp(i64)
 *       9 decimal p<=18 (as i64) 10 decimal p>18 (16 LE bytes).  `values` element width follows
 *       kind (bool = 1 byte per row). */
int co_murmur3_column(int kind, int64_t n, const void *values, const uint8_t *valid,
                      uint32_t *hashes);
/* Utf8/Binary: offsets int32[n+1] */
void co_murmur3_strings(int64_t n, const int32_t *offsets, const uint8_t *data,
                        const uint8_t *valid, uint32_t *hashes);
/* shuffle/src/comet_partitioning.rs:51-57 */
uint32_t co_pmod(uint32_t hash, uint32_t n);
/* shuffle/src/partitioners/multi_partition.rs:54-99: stable counting sort of row indices by
 * partition id.  starts has n_parts+1 entries. */
void co_partition_rows(int64_t n, const uint32_t *hashes, uint32_t n_parts, uint32_t *pids,
                       int64_t *starts, int64_t *row_idx);

/* ---- grouped accumulators (GroupsAccumulator semantics, row order) -------------------------- */
/* filter: NULL or byte-per-row (1 = keep).  group_idx: int64 per row. */
/* spark-expr/src/agg_funcs/sum_decimal.rs:418-475 ; state (sum, sum_valid, is_empty) per group,
 * caller initialises sum=0,sum_valid=1,is_empty=1 (resize_helper :403-407). */
int co_sum_decimal_update(int64_t n, const co_i128 *v, const uint8_t *valid, const uint8_t *filter,
                          const int64_t *group_idx, co_i128 *sum, uint8_t *sum_valid,
                          uint8_t *is_empty, int precision, int eval_mode);
/* sum_decimal.rs:540-607 */
int co_sum_decimal_merge(int64_t n, const co_i128 *that_sum, const uint8_t *that_sum_valid,
                         const uint8_t *that_is_empty, const int64_t *group_idx, co_i128 *sum,
                         uint8_t *sum_valid, uint8_t *is_empty, int precision, int eval_mode);
/* sum_decimal.rs:477-524 */
void co_sum_decimal_evaluate(int64_t n_groups, const co_i128 *sum, const uint8_t *sum_valid,
                             const uint8_t *is_empty, int precision, co_i128 *out, uint8_t *outv);
/* ungrouped SumDecimalAccumulator (sum_decimal.rs:176-369): state is one (sum,sum_valid,is_empty),
 * caller initialises sum=0,sum_valid=1,is_empty=1. */
int co_sum_decimal_acc_update(int64_t n, const co_i128 *v, const uint8_t *valid, co_i128 *sum,
                              uint8_t *sum_valid, uint8_t *is_empty, int precision, int eval_mode);

/* spark-expr/src/agg_funcs/avg_decimal.rs:483-495,505-540 ; caller initialises sums=0,counts=0,
 * is_not_null=1. */
int co_avg_decimal_update(int64_t n, const co_i128 *v, const uint8_t *valid, const uint8_t *filter,
                          const int64_t *group_idx, co_i128 *sums, int64_t *counts,
                          uint8_t *is_not_null, int sum_precision);
/* avg_decimal.rs:542-595 */
int co_avg_decimal_merge(int64_t n, const co_i128 *psum, const uint8_t *psum_valid,
                         const int64_t *pcount, const uint8_t *pcount_valid,
                         const int64_t *group_idx, co_i128 *sums, int64_t *counts,
                         uint8_t *is_not_null, int sum_precision, int eval_mode);
/* avg_decimal.rs:597-636,670-689 */
int co_avg_decimal_evaluate(int64_t n_groups, const co_i128 *sums, const int64_t *counts,
                            const uint8_t *is_not_null, int sum_scale, int target_precision,
                            int target_scale, int eval_mode, co_i128 *out, uint8_t *outv);

/* spark-expr/src/agg_funcs/avg.rs:229-277 (update), :279-309 (merge == update on sums with
 * counts added), :311-327 evaluate */
void co_avg_f64_update(int64_t n, const double *v, const uint8_t *valid, const uint8_t *filter,
                       const int64_t *group_idx, double *sums, int64_t *counts);
void co_avg_f64_merge(int64_t n, const double *psum, const int64_t *pcount,
                      const int64_t *group_idx, double *sums, int64_t *counts);
void co_avg_f64_evaluate(int64_t n_groups, const double *sums, const int64_t *counts, double *out,
                         uint8_t *outv);

/* spark-expr/src/agg_funcs/sum_int.rs:393-530 (Legacy groups accumulator) + Ansi/Try variants.
 * state: sums, sums_valid (None until first non-null); Try adds has_all_nulls-style tracking via
 * `overflowed` byte per group. */
int co_sum_int_update(int64_t n, const int64_t *v, const uint8_t *valid, const uint8_t *filter,
                      const int64_t *group_idx, int64_t *sums, uint8_t *sums_valid,
                      uint8_t *overflowed, int eval_mode);

/* datafusion-functions-aggregate 54.1.0 (3P, unpinned): f64 sum in row order per group, NULL
 * until first non-null; count of non-null; min/max. */
void co_sum_f64_update(int64_t n, const double *v, const uint8_t *valid, const uint8_t *filter,
                       const int64_t *group_idx, double *sums, uint8_t *sums_valid);
void co_count_update(int64_t n, const uint8_t *valid, const uint8_t *filter,
                     const int64_t *group_idx, int64_t *counts);

/* exact (correctly rounded) sum of doubles per group: the yardstick for the 1-ULP float bar. */
void co_sum_f64_exact(int64_t n, const double *v, const uint8_t *valid, const uint8_t *filter,
                      const int64_t *group_idx, int64_t n_groups, double *out);

/* ---- whole-pipeline CPU baselines (OpenMP; partition = contiguous row chunk, partial -> final
 *      exactly as the reference runs one plan per Spark partition) ------------------------------ */
/* TPC-H Q1, DECIMAL(12,2) money columns (spark/src/test/scala/org/apache/spark/sql/TPCH.scala:153-156),
 * expression tree per SURVEY.md section 8(a).  keys: dense group id = returnflag_code*n_ls + linestatus_code.
 * Outputs per group (n_groups = n_rf*n_ls): sum_qty d(22,2), sum_base d(22,2), sum_disc d(36,4),
 * sum_charge d(38,6), avg_qty d(16,6), avg_price d(16,6), avg_disc d(16,6), count. *_valid bytes. */
typedef struct {
    co_i128 sum_qty, sum_base, sum_disc_price, sum_charge, avg_qty, avg_price, avg_disc;
    int64_t count;
    uint8_t v_sum_qty, v_sum_base, v_sum_disc_price, v_sum_charge, v_avg_qty, v_avg_price,
        v_avg_disc, present;
} co_q1_dec_row;
int co_q1_dec(int64_t n, const co_i128 *qty, const co_i128 *price, const co_i128 *disc,
              const co_i128 *tax, const int32_t *shipdate, const uint8_t *rf_code,
              const uint8_t *ls_code, int n_rf, int n_ls, int32_t date_cutoff, int n_threads,
              co_q1_dec_row *out);
typedef struct {
    double sum_qty, sum_base, sum_disc_price, sum_charge, avg_qty, avg_price, avg_disc;
    int64_t count;
    uint8_t present;
} co_q1_f64_row;
int co_q1_f64(int64_t n, const double *qty, const double *price, const double *disc,
              const double *tax, const int32_t *shipdate, const uint8_t *rf_code,
              const uint8_t *ls_code, int n_rf, int n_ls, int32_t date_cutoff, int n_threads,
              co_q1_f64_row *out);
/* TPC-H Q6: sum(l_extendedprice*l_discount) where shipdate in [lo,hi), discount in [dlo,dhi],
 * quantity < qmax.  DEC: d(12,2)*d(12,2) -> d(25,4) plain mul + CheckOverflow, sum -> d(35,4). */
int co_q6_dec(int64_t n, const co_i128 *qty, const co_i128 *price, const co_i128 *disc,
              const int32_t *shipdate, int32_t date_lo, int32_t date_hi, const co_i128 *disc_lo,
              const co_i128 *disc_hi, const co_i128 *qty_max, int n_threads, co_i128 *out,
              uint8_t *out_valid);
int co_q6_f64(int64_t n, const double *qty, const double *price, const double *disc,
              const int32_t *shipdate, int32_t date_lo, int32_t date_hi, double disc_lo,
              double disc_hi, double qty_max, int n_threads, double *out, uint8_t *out_valid);
/* Config 1: SELECT l_quantity*l_extendedprice WHERE l_shipdate < cutoff.  Returns rows kept. */
int64_t co_filter_project_dec(int64_t n, const co_i128 *qty, const co_i128 *price,
                              const int32_t *shipdate, int32_t cutoff, int n_threads,
                              co_i128 *out, uint8_t *outv);
int64_t co_filter_project_f64(int64_t n, const double *qty, const double *price,
                              const int32_t *shipdate, int32_t cutoff, int n_threads, double *out);

#ifdef __cplusplus
}
#endif
/* first-touch placement of baseline inputs (see comet_oracle.c) */
void co_parallel_copy(void *dst, const void *src, int64_t n, int elem_bytes, int n_threads);

/* GROUP BY key SUM(value), d(p,s): Partial per thread-partition, murmur3/pmod repartition, Final merge (BASELINE configs[3]) */
int64_t co_groupby_sum_dec(int64_t n, const int64_t *keys, const co_i128 *vals, int precision, int n_threads,
                           int64_t *out_keys, co_i128 *out_sum, uint8_t *out_valid);

#endif
