// plan.h -- in-memory form of the reference's plan IR (spark.spark_operator.Operator and friends)
// after decoding, plus the type rules the reference's planner applies
// (native/core/src/execution/planner.rs:446-1131 create_expr / create_binary_expr_with_options,
//  :2558-2917 create_agg_expr, serde.rs:71-110 to_arrow_datatype).
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace cb200 {

struct Unsupported : std::runtime_error { // plan uses something outside the GPU hot path -> caller falls back
    using std::runtime_error::runtime_error;
};
struct PlanError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// types.proto:43-66 DataTypeId (only the ids on the hot path are accepted)
enum class TypeId : int {
    Bool = 0, Int8 = 1, Int16 = 2, Int32 = 3, Int64 = 4, Float32 = 5, Float64 = 6, String = 7, Binary = 8,
    Timestamp = 9, Decimal = 10, TimestampNtz = 11, Date = 12, Null = 13
};

struct DType {
    TypeId id = TypeId::Null;
    int precision = 0, scale = 0;
    bool is_decimal() const { return id == TypeId::Decimal; }
    bool is_integer() const { return id == TypeId::Int8 || id == TypeId::Int16 || id == TypeId::Int32 || id == TypeId::Int64; }
    bool is_float() const { return id == TypeId::Float32 || id == TypeId::Float64; }
    bool is_string() const { return id == TypeId::String || id == TypeId::Binary; }
    bool operator==(const DType& o) const {
        return id == o.id && (id != TypeId::Decimal || (precision == o.precision && scale == o.scale));
    }
    bool operator!=(const DType& o) const { return !(*this == o); }
    std::string str() const;
    // bytes of one value in Arrow layout (0 = bitmap-packed bool, -1 = variable width)
    int arrow_width() const;
};
inline DType mk_decimal(int p, int s) { DType d; d.id = TypeId::Decimal; d.precision = p; d.scale = s; return d; }
inline DType mk_type(TypeId id) { DType d; d.id = id; return d; }

enum class EvalMode : int { Legacy = 0, Try = 1, Ansi = 2 }; // expr.proto:324

enum class ExprKind {
    Literal, Bound, Unbound,
    Add, Sub, Mul, Div,
    Eq, Neq, Gt, GtEq, Lt, LtEq,
    IsNull, IsNotNull, And, Or, Not,
    Cast, CheckOverflow, UnaryMinus, If, In
};

struct Expr;
using ExprP = std::shared_ptr<Expr>;

struct Expr {
    ExprKind kind;
    std::vector<ExprP> children;
    DType type;            // resolved result type (what PhysicalExpr::data_type would return)
    // Literal
    bool lit_null = false;
    int64_t lit_i64 = 0;   // bool/int/date/timestamp
    double lit_f64 = 0;    // float/double
    unsigned __int128 lit_dec = 0; // decimal unscaled (two's complement)
    std::string lit_str;
    // Bound / Unbound
    int index = -1;
    std::string name;
    // MathExpr / Cast / CheckOverflow / UnaryMinus
    DType return_type;
    EvalMode eval_mode = EvalMode::Legacy;
    bool fail_on_error = false;
    bool negated = false;  // In
    bool integral_div = false;          // Div: IntegralDivide (expr.proto:81): the quotient without the HALF_UP digit
    bool check_divide_overflow = false; // MathExpr.check_divide_overflow (expr.proto:335-340)
    // decimal arithmetic lowering chosen by the reference's rule (planner.rs:998-1027)
    bool wide_decimal = false;
};

enum class AggKind { Count, Sum, Min, Max, Avg };
enum class AggMode : int { Partial = 0, Final = 1, PartialMerge = 2 }; // operator.proto AggregateMode

struct AggExpr {
    AggKind kind;
    std::vector<ExprP> children; // Count may have several
    DType datatype;              // result type
    DType sum_datatype;          // Avg: sum state type
    EvalMode eval_mode = EvalMode::Legacy;
    ExprP filter;                // FILTER (WHERE ...) clause, Partial mode only
};

enum class OpKind { Scan, ShuffleScan, NativeScan, Projection, Filter, HashAgg, ShuffleWriter };

struct StructField {
    std::string name;
    DType type;
    bool nullable = true;
};

struct Operator;
using OperatorP = std::shared_ptr<Operator>;

struct Operator {
    OpKind kind;
    uint32_t plan_id = 0;
    std::vector<OperatorP> children;
    std::vector<DType> schema;       // output column types (child col_i naming is positional)
    // Scan / ShuffleScan
    std::vector<DType> fields;
    std::string source;
    // NativeScan
    std::vector<StructField> required_schema, data_schema;
    std::vector<int64_t> projection_vector;
    std::vector<ExprP> data_filters;
    std::vector<std::string> files;
    std::vector<int64_t> file_start, file_length; // SparkPartitionedFile.start / length (operator.proto:103-109); 0 / 0 = the whole file
    // Projection
    std::vector<ExprP> project_list;
    // Filter
    ExprP predicate;
    // HashAgg
    std::vector<ExprP> grouping;
    std::vector<AggExpr> aggs;
    AggMode mode = AggMode::Partial;
    // ShuffleWriter (hash partitioning only)
    std::vector<ExprP> hash_exprs;
    int num_partitions = 0;
};

// Decode + resolve types.  Throws Unsupported for anything outside the GPU hot path and PlanError
// for malformed plans.
OperatorP decode_plan(const uint8_t* data, size_t len);

// state-column layout an aggregate exposes in Partial mode / consumes in Final mode
// (sum_decimal.rs:112-120, avg_decimal.rs:132-145, avg.rs:82-95, sum_int.rs:75-84)
std::vector<DType> agg_state_types(const AggExpr& a);
DType agg_result_type(const AggExpr& a);

std::string expr_str(const Expr& e);

} // namespace cb200
