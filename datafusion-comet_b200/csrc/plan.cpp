// plan.cpp -- decode the reference's protobuf plan IR and resolve expression types the way the
// reference's planner does.  Field numbers: native/proto/src/proto/{operator,expr,literal,types,
// partitioning}.proto (cited inline).  Type rules: native/core/src/execution/planner.rs.
#include "plan.h"
#include "proto_wire.h"

#include <algorithm>
#include <sstream>

namespace cb200 {

std::string DType::str() const {
    switch (id) {
    case TypeId::Bool: return "bool";
    case TypeId::Int8: return "int8";
    case TypeId::Int16: return "int16";
    case TypeId::Int32: return "int32";
    case TypeId::Int64: return "int64";
    case TypeId::Float32: return "float32";
    case TypeId::Float64: return "float64";
    case TypeId::String: return "utf8";
    case TypeId::Binary: return "binary";
    case TypeId::Timestamp: return "timestamp[us,UTC]";
    case TypeId::TimestampNtz: return "timestamp[us]";
    case TypeId::Date: return "date32";
    case TypeId::Null: return "null";
    case TypeId::Decimal: {
        std::ostringstream o;
        o << "decimal128(" << precision << "," << scale << ")";
        return o.str();
    }
    }
    return "?";
}

int DType::arrow_width() const {
    switch (id) {
    case TypeId::Bool: return 0;
    case TypeId::Int8: return 1;
    case TypeId::Int16: return 2;
    case TypeId::Int32: case TypeId::Float32: case TypeId::Date: return 4;
    case TypeId::Int64: case TypeId::Float64: case TypeId::Timestamp: case TypeId::TimestampNtz: return 8;
    case TypeId::Decimal: return 16;
    default: return -1;
    }
}

// ---- DataType (types.proto:43-114) ----------------------------------------------------------------
static DType decode_dtype(PbReader r) {
    DType d;
    int id = 0;
    while (r.next()) {
        if (r.field == 1) id = (int)r.i64();
        else if (r.field == 2) { // DataTypeInfo
            PbReader info = r.sub();
            while (info.next()) {
                if (info.field == 2) { // DecimalInfo
                    PbReader di = info.sub();
                    while (di.next()) {
                        if (di.field == 1) d.precision = (int)di.i64();
                        else if (di.field == 2) d.scale = (int)di.i64();
                        else di.skip();
                    }
                } else throw Unsupported("nested data types (list/map/struct) are outside the GPU hot path");
            }
        } else r.skip();
    }
    if (id < 0 || id > 13) throw Unsupported("data type id " + std::to_string(id) + " is outside the GPU hot path");
    d.id = (TypeId)id;
    return d;
}

static ExprP decode_expr(PbReader r);

static ExprP mk(ExprKind k) {
    auto e = std::make_shared<Expr>();
    e->kind = k;
    return e;
}

// Literal (literal.proto:26-47): value is big-endian two's-complement bytes for decimals
// (planner.rs:544-548 BigInt::from_signed_bytes_be).
static ExprP decode_literal(PbReader r) {
    auto e = mk(ExprKind::Literal);
    bool have_value = false;
    std::string dec_bytes;
    while (r.next()) {
        switch (r.field) {
        case 1: case 2: case 3: case 4: case 5: e->lit_i64 = r.i64(); have_value = true; break;
        case 6: e->lit_f64 = (double)r.f32(); have_value = true; break;
        case 7: e->lit_f64 = r.f64(); have_value = true; break;
        case 8: case 9: e->lit_str = r.bytes(); have_value = true; break;
        case 10: dec_bytes = r.bytes(); have_value = true; break;
        case 11: throw Unsupported("list literals are outside the GPU hot path");
        case 12: e->type = decode_dtype(r.sub()); break;
        case 13: e->lit_null = r.i64() != 0; break;
        default: r.skip();
        }
    }
    if (!dec_bytes.empty() || e->type.is_decimal()) {
        if (dec_bytes.size() > 16) throw PlanError("decimal literal does not fit in i128");
        unsigned __int128 v = (!dec_bytes.empty() && ((uint8_t)dec_bytes[0] & 0x80)) ? ~(unsigned __int128)0 : 0;
        for (unsigned char c : dec_bytes) v = (v << 8) | c;
        e->lit_dec = v;
    }
    if (e->type.id == TypeId::Int8) e->lit_i64 = (int8_t)e->lit_i64;
    if (e->type.id == TypeId::Int16) e->lit_i64 = (int16_t)e->lit_i64;
    if (e->type.id == TypeId::Int32 || e->type.id == TypeId::Date) e->lit_i64 = (int32_t)e->lit_i64;
    if (!have_value && !e->lit_null && e->type.id != TypeId::Null) {
        // proto3 omits default-valued scalars: a present-but-zero literal (0, false, 0.0, "")
    }
    return e;
}

static ExprP decode_binary(ExprKind k, PbReader r, bool math) {
    auto e = mk(k);
    ExprP l, rr;
    while (r.next()) {
        if (r.field == 1) l = decode_expr(r.sub());
        else if (r.field == 2) rr = decode_expr(r.sub());
        else if (math && r.field == 4) e->return_type = decode_dtype(r.sub());
        else if (math && r.field == 5) e->eval_mode = (EvalMode)r.i64();
        else if (math && r.field == 6) e->check_divide_overflow = r.i64() != 0;
        else r.skip();
    }
    if (!l || !rr) throw PlanError("binary expression is missing an operand");
    e->children = {l, rr};
    return e;
}

static ExprP decode_unary(ExprKind k, PbReader r) {
    auto e = mk(k);
    while (r.next()) {
        if (r.field == 1) e->children.push_back(decode_expr(r.sub()));
        else if (k == ExprKind::UnaryMinus && r.field == 2) e->fail_on_error = r.i64() != 0;
        else r.skip();
    }
    if (e->children.size() != 1) throw PlanError("unary expression is missing its child");
    return e;
}

// Expr (expr.proto:30-109)
static ExprP decode_expr(PbReader r) {
    ExprP out;
    while (r.next()) {
        if (r.wire != 2) { r.skip(); continue; } // expr_id (91) etc.
        switch (r.field) {
        case 2: out = decode_literal(r.sub()); break;
        case 3: { // BoundReference expr.proto:375
            out = mk(ExprKind::Bound);
            PbReader b = r.sub();
            while (b.next()) {
                if (b.field == 1) out->index = (int)b.i64();
                else if (b.field == 2) out->type = decode_dtype(b.sub());
                else b.skip();
            }
            if (out->index < 0) out->index = 0;
            break;
        }
        case 51: { // UnboundReference
            out = mk(ExprKind::Unbound);
            PbReader b = r.sub();
            while (b.next()) {
                if (b.field == 1) out->name = b.bytes();
                else if (b.field == 2) out->type = decode_dtype(b.sub());
                else b.skip();
            }
            break;
        }
        case 4: out = decode_binary(ExprKind::Add, r.sub(), true); break;
        case 5: out = decode_binary(ExprKind::Sub, r.sub(), true); break;
        case 6: out = decode_binary(ExprKind::Mul, r.sub(), true); break;
        case 7: out = decode_binary(ExprKind::Div, r.sub(), true); break;
        case 59: out = decode_binary(ExprKind::Div, r.sub(), true); out->integral_div = true; break; // IntegralDivide
        case 8: { // Cast expr.proto:337
            out = mk(ExprKind::Cast);
            PbReader c = r.sub();
            while (c.next()) {
                if (c.field == 1) out->children.push_back(decode_expr(c.sub()));
                else if (c.field == 2) out->return_type = decode_dtype(c.sub());
                else if (c.field == 4) out->eval_mode = (EvalMode)c.i64();
                else c.skip();
            }
            if (out->children.size() != 1) throw PlanError("cast is missing its child");
            break;
        }
        case 9: out = decode_binary(ExprKind::Eq, r.sub(), false); break;
        case 10: out = decode_binary(ExprKind::Neq, r.sub(), false); break;
        case 11: out = decode_binary(ExprKind::Gt, r.sub(), false); break;
        case 12: out = decode_binary(ExprKind::GtEq, r.sub(), false); break;
        case 13: out = decode_binary(ExprKind::Lt, r.sub(), false); break;
        case 14: out = decode_binary(ExprKind::LtEq, r.sub(), false); break;
        case 15: out = decode_unary(ExprKind::IsNull, r.sub()); break;
        case 16: out = decode_unary(ExprKind::IsNotNull, r.sub()); break;
        case 17: out = decode_binary(ExprKind::And, r.sub(), false); break;
        case 18: out = decode_binary(ExprKind::Or, r.sub(), false); break;
        case 25: { // CheckOverflow
            out = mk(ExprKind::CheckOverflow);
            PbReader c = r.sub();
            while (c.next()) {
                if (c.field == 1) out->children.push_back(decode_expr(c.sub()));
                else if (c.field == 2) out->return_type = decode_dtype(c.sub());
                else if (c.field == 3) out->fail_on_error = c.i64() != 0;
                else c.skip();
            }
            if (out->children.size() != 1) throw PlanError("check_overflow is missing its child");
            break;
        }
        case 39: { // In
            out = mk(ExprKind::In);
            PbReader c = r.sub();
            while (c.next()) {
                if (c.field == 1 || c.field == 2) out->children.push_back(decode_expr(c.sub()));
                else if (c.field == 3) out->negated = c.i64() != 0;
                else c.skip();
            }
            break;
        }
        case 40: out = decode_unary(ExprKind::Not, r.sub()); break;
        case 41: out = decode_unary(ExprKind::UnaryMinus, r.sub()); break;
        case 44: { // IfExpr
            out = mk(ExprKind::If);
            PbReader c = r.sub();
            ExprP a, b, d;
            while (c.next()) {
                if (c.field == 1) a = decode_expr(c.sub());
                else if (c.field == 2) b = decode_expr(c.sub());
                else if (c.field == 3) d = decode_expr(c.sub());
                else c.skip();
            }
            if (!a || !b || !d) throw PlanError("if expression is missing an operand");
            out->children = {a, b, d};
            break;
        }
        case 38: { // CaseWhen expr.proto:473 -> nested IF chain (planner.rs:677-703 builds a CaseExpr with expr = None)
            PbReader c = r.sub();
            std::vector<ExprP> whens, thens;
            ExprP els;
            while (c.next()) {
                if (c.field == 2) whens.push_back(decode_expr(c.sub()));
                else if (c.field == 3) thens.push_back(decode_expr(c.sub()));
                else if (c.field == 4) els = decode_expr(c.sub());
                else c.skip();
            }
            if (whens.empty() || whens.size() != thens.size()) throw PlanError("CASE WHEN needs matching when/then lists");
            ExprP tail = els; // may be null: ELSE NULL, typed when types are resolved
            for (size_t i = whens.size(); i-- > 0;) {
                ExprP n = mk(ExprKind::If);
                n->children = {whens[i], thens[i]};
                if (tail) n->children.push_back(tail);
                tail = n;
            }
            out = tail;
            break;
        }
        case 90: r.skip(); break; // query_context
        default:
            throw Unsupported("expression field " + std::to_string(r.field) + " is outside the GPU hot path");
        }
    }
    if (!out) throw PlanError("empty expression");
    return out;
}

// ---- type resolution ------------------------------------------------------------------------------
static bool cast_supported(const DType& from, const DType& to) {
    if (from == to) return true;
    auto numeric = [](const DType& d) { return d.is_integer() || d.is_float(); };
    if (numeric(from) && numeric(to)) {
        // widening / int->float only (narrowing needs Spark's overflow rules: conversion_funcs/numeric.rs)
        auto rank = [](const DType& d) {
            switch (d.id) {
            case TypeId::Int8: return 1; case TypeId::Int16: return 2; case TypeId::Int32: return 3;
            case TypeId::Int64: return 4; case TypeId::Float32: return 5; case TypeId::Float64: return 6;
            default: return 0;
            }
        };
        return rank(to) >= rank(from);
    }
    if (from.is_integer() && to.is_decimal()) return true;
    if (from.is_decimal() && to.is_decimal()) return true;
    if (from.is_decimal() && to.id == TypeId::Float64) return true;
    return false;
}

static void resolve(Expr& e, const std::vector<DType>& in) {
    for (auto& c : e.children) resolve(*c, in);
    auto ct = [&](int i) -> const DType& { return e.children[i]->type; };
    switch (e.kind) {
    case ExprKind::Literal: case ExprKind::Unbound: break;
    case ExprKind::Bound:
        if (e.index >= (int)in.size()) throw PlanError("bound reference index " + std::to_string(e.index) + " out of range");
        e.type = in[e.index];
        break;
    case ExprKind::Add: case ExprKind::Sub: case ExprKind::Mul: {
        const DType &l = ct(0), &r = ct(1);
        if (l.is_decimal() && r.is_decimal()) {
            // planner.rs:998-1027: wide path when the arrow-arith result could exceed precision 38
            bool wide;
            if (e.kind == ExprKind::Mul) wide = l.precision + r.precision >= 38;
            else wide = std::max(l.scale, r.scale) + std::max(l.precision - l.scale, r.precision - r.scale) >= 38;
            e.wide_decimal = wide;
            if (wide) {
                if (!e.return_type.is_decimal()) throw PlanError("Expected Decimal128 return type");
                e.type = e.return_type;
            } else if (e.kind == ExprKind::Mul) {
                e.type = mk_decimal(std::min(l.precision + r.precision + 1, 38), l.scale + r.scale);
            } else {
                int rs = std::max(l.scale, r.scale);
                e.type = mk_decimal(std::min(rs + std::max(l.precision - l.scale, r.precision - r.scale) + 1, 38), rs);
            }
        } else if ((l.is_integer() || l.is_float()) && l == r) {
            e.type = e.return_type.id == TypeId::Null ? l : e.return_type;
            if (e.type != l) throw Unsupported("arithmetic with implicit result cast " + l.str() + " -> " + e.type.str());
        } else {
            throw Unsupported("arithmetic on " + l.str() + " and " + r.str());
        }
        break;
    }
    case ExprKind::Div: {
        const DType &l = ct(0), &r = ct(1);
        if (l.is_decimal() && r.is_decimal()) { // decimal_div / decimal_integral_div UDFs (planner.rs:1028-1058): result type = the proto's
            if (!e.return_type.is_decimal()) throw PlanError("decimal division without a Decimal128 return type");
            e.type = e.return_type;
        } else if ((l.is_float() || l.is_integer()) && l == r) {
            if (e.integral_div && l.is_float()) throw Unsupported("integral division of floats");
            e.type = e.return_type.id == TypeId::Null ? l : e.return_type;
            if (e.type != l) throw Unsupported("division with implicit result cast " + l.str() + " -> " + e.type.str());
        } else throw Unsupported("division on " + l.str() + " and " + r.str());
        break;
    }
    case ExprKind::Eq: case ExprKind::Neq: case ExprKind::Gt: case ExprKind::GtEq: case ExprKind::Lt: case ExprKind::LtEq: {
        const DType &l = ct(0), &r = ct(1);
        bool ok = l == r || (l.is_decimal() && r.is_decimal() && l.scale == r.scale);
        if (!ok || l.is_string() || l.id == TypeId::Null)
            throw Unsupported("comparison between " + l.str() + " and " + r.str());
        e.type = mk_type(TypeId::Bool);
        break;
    }
    case ExprKind::And: case ExprKind::Or:
        if (ct(0).id != TypeId::Bool || ct(1).id != TypeId::Bool) throw PlanError("AND/OR over non-boolean operands");
        e.type = mk_type(TypeId::Bool);
        break;
    case ExprKind::Not:
        if (ct(0).id != TypeId::Bool) throw PlanError("NOT over non-boolean operand");
        e.type = mk_type(TypeId::Bool);
        break;
    case ExprKind::IsNull: case ExprKind::IsNotNull: e.type = mk_type(TypeId::Bool); break;
    case ExprKind::Cast:
        if (!cast_supported(ct(0), e.return_type))
            throw Unsupported("cast " + ct(0).str() + " -> " + e.return_type.str() + " is outside the GPU hot path");
        e.type = e.return_type;
        break;
    case ExprKind::CheckOverflow:
        if (!ct(0).is_decimal() || !e.return_type.is_decimal()) throw PlanError("CheckOverflow expects only Decimal128");
        e.type = e.return_type;
        break;
    case ExprKind::UnaryMinus:
        if (ct(0).is_string() || ct(0).id == TypeId::Bool) throw Unsupported("negation of " + ct(0).str());
        e.type = ct(0);
        break;
    case ExprKind::If:
        if (e.children.size() == 2) { // CASE without ELSE: NULL of the THEN type
            auto nul = std::make_shared<Expr>();
            nul->kind = ExprKind::Literal;
            nul->lit_null = true;
            nul->type = ct(1);
            e.children.push_back(nul);
        }
        if (ct(0).id != TypeId::Bool || ct(1) != ct(2)) throw Unsupported("IF with mismatched branch types");
        e.type = ct(1);
        break;
    case ExprKind::In:
        for (size_t i = 1; i < e.children.size(); i++) {
            if (e.children[i]->kind != ExprKind::Literal) throw Unsupported("IN list with non-literal members");
            const DType &l = ct(0), &r = ct((int)i);
            if (!(l == r || (l.is_decimal() && r.is_decimal() && l.scale == r.scale)) || l.is_string())
                throw Unsupported("IN over " + l.str() + " / " + r.str());
        }
        e.type = mk_type(TypeId::Bool);
        break;
    }
}

// ---- aggregates -----------------------------------------------------------------------------------
static AggExpr decode_agg(PbReader r) { // AggExpr expr.proto:143-176
    AggExpr a;
    bool have = false;
    while (r.next()) {
        if (r.wire != 2) { r.skip(); continue; }
        if (r.field >= 2 && r.field <= 6) {
            PbReader b = r.sub();
            have = true;
            switch (r.field) {
            case 2: a.kind = AggKind::Count; break;
            case 3: a.kind = AggKind::Sum; break;
            case 4: a.kind = AggKind::Min; break;
            case 5: a.kind = AggKind::Max; break;
            case 6: a.kind = AggKind::Avg; break;
            }
            while (b.next()) {
                if (b.field == 1) a.children.push_back(decode_expr(b.sub()));
                else if (a.kind != AggKind::Count && b.field == 2) a.datatype = decode_dtype(b.sub());
                else if (a.kind == AggKind::Sum && b.field == 3) a.eval_mode = (EvalMode)b.i64();
                else if (a.kind == AggKind::Avg && b.field == 3) a.sum_datatype = decode_dtype(b.sub());
                else if (a.kind == AggKind::Avg && b.field == 4) a.eval_mode = (EvalMode)b.i64();
                else b.skip();
            }
        } else if (r.field == 89) a.filter = decode_expr(r.sub());
        else if (r.field == 90) r.skip();
        else throw Unsupported("aggregate function field " + std::to_string(r.field) + " is outside the GPU hot path");
    }
    if (!have) throw PlanError("empty aggregate expression");
    if (a.kind == AggKind::Count) a.datatype = mk_type(TypeId::Int64);
    return a;
}

DType agg_result_type(const AggExpr& a) {
    switch (a.kind) {
    case AggKind::Count: return mk_type(TypeId::Int64);
    case AggKind::Avg: return a.datatype.is_decimal() ? a.datatype : mk_type(TypeId::Float64); // planner.rs:2656-2676
    default: return a.datatype;
    }
}

std::vector<DType> agg_state_types(const AggExpr& a) {
    switch (a.kind) {
    case AggKind::Count: return {mk_type(TypeId::Int64)};
    case AggKind::Sum:
        if (a.datatype.is_decimal()) return {a.datatype, mk_type(TypeId::Bool)}; // (sum, is_empty) sum_decimal.rs:112-120
        if (a.datatype.is_integer()) {
            if (a.eval_mode == EvalMode::Try) return {mk_type(TypeId::Int64), mk_type(TypeId::Bool)}; // sum_int.rs:75-84
            return {mk_type(TypeId::Int64)};
        }
        return {a.datatype};
    case AggKind::Avg:
        if (a.datatype.is_decimal()) return {a.sum_datatype, mk_type(TypeId::Int64)}; // avg_decimal.rs:132-145
        return {mk_type(TypeId::Float64), mk_type(TypeId::Int64)};                    // avg.rs:82-95
    case AggKind::Min: case AggKind::Max: return {a.datatype};
    }
    return {};
}

static void resolve_agg(AggExpr& a, const std::vector<DType>& in, AggMode mode) {
    if (mode == AggMode::Partial) {
        for (auto& c : a.children) resolve(*c, in);
        if (a.filter) resolve(*a.filter, in);
    }
    if (a.children.empty()) throw PlanError("aggregate without arguments");
    const DType& dt = a.datatype;
    switch (a.kind) {
    case AggKind::Count: break;
    case AggKind::Sum:
        if (!(dt.is_decimal() || dt.is_integer() || dt.id == TypeId::Float64 || dt.id == TypeId::Float32))
            throw Unsupported("SUM over " + dt.str());
        break;
    case AggKind::Avg:
        if (dt.is_decimal() && !a.sum_datatype.is_decimal()) throw PlanError("AVG(decimal) without a decimal sum type");
        if (!dt.is_decimal() && !(dt.id == TypeId::Float64)) throw Unsupported("AVG result type " + dt.str());
        break;
    case AggKind::Min: case AggKind::Max:
        if (!(dt.is_integer() || dt.id == TypeId::Date || dt.id == TypeId::Timestamp || dt.id == TypeId::TimestampNtz ||
              dt.id == TypeId::Float64 || (dt.is_decimal() && dt.precision <= 18)))
            throw Unsupported("MIN/MAX over " + dt.str());
        break;
    }
    if (mode == AggMode::Partial && a.kind != AggKind::Count) {
        const DType& ct = a.children[0]->type;
        bool ok = ct == dt || (a.kind == AggKind::Avg) ||
                  (a.kind == AggKind::Sum && ((dt.is_decimal() && ct.is_decimal() && ct.scale == dt.scale) ||
                                              (dt.is_integer() && ct.is_integer()) || (dt.is_float() && (ct.is_float() || ct.is_integer()))));
        if (!ok) throw Unsupported("aggregate input " + ct.str() + " for result " + dt.str());
        if (a.kind == AggKind::Avg) {
            if (dt.is_decimal() && !(ct.is_decimal() && ct.scale == a.sum_datatype.scale))
                throw Unsupported("AVG(decimal) over " + ct.str());
            if (!dt.is_decimal() && !(ct.is_integer() || ct.is_float())) throw Unsupported("AVG over " + ct.str());
        }
    }
}

// ---- operators ------------------------------------------------------------------------------------
static StructField decode_struct_field(PbReader r) { // SparkStructField operator.proto:97-102
    StructField f;
    while (r.next()) {
        if (r.field == 1) f.name = r.bytes();
        else if (r.field == 2) f.type = decode_dtype(r.sub());
        else if (r.field == 3) f.nullable = r.i64() != 0;
        else r.skip();
    }
    return f;
}

static OperatorP decode_operator(PbReader r) { // Operator operator.proto:32-86
    auto op = std::make_shared<Operator>();
    bool have = false;
    std::vector<PbReader> child_readers;
    // children may precede or follow the op payload; decode children first so schemas are known
    struct Pending { uint32_t field; PbReader rd; };
    std::vector<Pending> payload;
    while (r.next()) {
        if (r.field == 1) child_readers.push_back(r.sub());
        else if (r.field == 2) op->plan_id = (uint32_t)r.i64();
        else if (r.field >= 100 && r.wire == 2) payload.push_back({r.field, r.sub()});
        else r.skip();
    }
    for (auto& c : child_readers) op->children.push_back(decode_operator(c));
    if (payload.size() != 1) throw PlanError("operator must carry exactly one op_struct");
    uint32_t f = payload[0].field;
    PbReader b = payload[0].rd;
    auto child_schema = [&]() -> const std::vector<DType>& {
        if (op->children.size() != 1) throw PlanError("operator expects exactly one child");
        return op->children[0]->schema;
    };
    switch (f) {
    case 100: case 116: { // Scan operator.proto:104-107 / ShuffleScan
        op->kind = f == 100 ? OpKind::Scan : OpKind::ShuffleScan;
        while (b.next()) {
            if (b.field == 1) op->fields.push_back(decode_dtype(b.sub()));
            else if (b.field == 2) op->source = b.bytes();
            else b.skip();
        }
        op->schema = op->fields;
        have = true;
        break;
    }
    case 111: { // NativeScan operator.proto:141-185
        op->kind = OpKind::NativeScan;
        while (b.next()) {
            if (b.field == 1) { // NativeScanCommon
                PbReader c = b.sub();
                while (c.next()) {
                    if (c.field == 1) op->required_schema.push_back(decode_struct_field(c.sub()));
                    else if (c.field == 2) op->data_schema.push_back(decode_struct_field(c.sub()));
                    else if (c.field == 3) throw Unsupported("partition columns in NativeScan");
                    else if (c.field == 4) op->data_filters.push_back(decode_expr(c.sub()));
                    else if (c.field == 5) {
                        if (c.wire == 2) { PbReader pk = c.sub(); while (pk.p < pk.end) op->projection_vector.push_back((int64_t)pk.varint()); }
                        else op->projection_vector.push_back(c.i64());
                    } else if (c.field == 12) op->source = c.bytes();
                    else c.skip();
                }
            } else if (b.field == 2) { // SparkFilePartition
                PbReader fp = b.sub();
                while (fp.next()) {
                    if (fp.field == 1) {
                        PbReader pf = fp.sub();
                        int64_t start = 0, length = 0;
                        while (pf.next()) {
                            if (pf.field == 1) op->files.push_back(pf.bytes());
                            else if (pf.field == 2) start = pf.i64();
                            else if (pf.field == 3) length = pf.i64();
                            else if (pf.field == 5) throw Unsupported("partition values in NativeScan");
                            else pf.skip();
                        }
                        op->file_start.push_back(start);
                        op->file_length.push_back(length);
                    } else fp.skip();
                }
            } else b.skip();
        }
        for (auto& sf : op->required_schema) op->schema.push_back(sf.type);
        for (auto& e : op->data_filters) resolve(*e, op->schema);
        have = true;
        break;
    }
    case 101: { // Projection operator.proto:633
        op->kind = OpKind::Projection;
        while (b.next()) {
            if (b.field == 1) op->project_list.push_back(decode_expr(b.sub()));
            else b.skip();
        }
        for (auto& e : op->project_list) { resolve(*e, child_schema()); op->schema.push_back(e->type); }
        have = true;
        break;
    }
    case 102: { // Filter operator.proto:637
        op->kind = OpKind::Filter;
        while (b.next()) {
            if (b.field == 1) op->predicate = decode_expr(b.sub());
            else b.skip();
        }
        if (!op->predicate) throw PlanError("filter without predicate");
        resolve(*op->predicate, child_schema());
        if (op->predicate->type.id != TypeId::Bool) throw PlanError("filter predicate is not boolean");
        op->schema = child_schema();
        have = true;
        break;
    }
    case 104: { // HashAggregate operator.proto:647
        op->kind = OpKind::HashAgg;
        std::vector<int64_t> expr_modes;
        while (b.next()) {
            if (b.field == 1) op->grouping.push_back(decode_expr(b.sub()));
            else if (b.field == 2) op->aggs.push_back(decode_agg(b.sub()));
            else if (b.field == 5) op->mode = (AggMode)b.i64();
            else if (b.field == 6) {
                if (b.wire == 2) { PbReader pk = b.sub(); while (pk.p < pk.end) expr_modes.push_back((int64_t)pk.varint()); }
                else expr_modes.push_back(b.i64());
            } else b.skip();
        }
        // Operator-level PartialMerge (merge state columns, emit state columns) is the Final input path with the Partial
        // output path.  Per-expression modes that DIFFER from the operator's (the distinct rewrite: MergeAsPartialUDF,
        // merge_as_partial.rs:54-110) are not fused.
        for (int64_t m : expr_modes)
            if (m != (int64_t)op->mode) throw Unsupported("aggregate expressions whose mode differs from the operator's (distinct rewrite, merge_as_partial.rs) are outside the GPU hot path");
        const auto& cs = child_schema();
        for (auto& g : op->grouping) { resolve(*g, cs); op->schema.push_back(g->type); }
        for (auto& a : op->aggs) {
            resolve_agg(a, cs, op->mode);
            if (op->mode != AggMode::Final) for (auto& t : agg_state_types(a)) op->schema.push_back(t);
            else op->schema.push_back(agg_result_type(a));
        }
        if (op->mode != AggMode::Partial) {
            // DataFusion Final mode reads state columns positionally after the group columns
            size_t need = op->grouping.size();
            for (auto& a : op->aggs) need += agg_state_types(a).size();
            if (cs.size() < need) throw PlanError("final / partial-merge aggregate: child has fewer columns than group + state columns");
            size_t at = op->grouping.size();
            for (auto& a : op->aggs)
                for (auto& t : agg_state_types(a)) {
                    if (cs[at] != t) throw PlanError("final aggregate: state column " + std::to_string(at) + " is " + cs[at].str() + ", expected " + t.str());
                    at++;
                }
        }
        have = true;
        break;
    }
    case 106: { // ShuffleWriter operator.proto:688 (hash partitioning only)
        op->kind = OpKind::ShuffleWriter;
        while (b.next()) {
            if (b.field == 1) { // Partitioning
                PbReader pt = b.sub();
                while (pt.next()) {
                    if (pt.field == 1) { // HashPartition partitioning.proto:38
                        PbReader hp = pt.sub();
                        while (hp.next()) {
                            if (hp.field == 1) op->hash_exprs.push_back(decode_expr(hp.sub()));
                            else if (hp.field == 2) op->num_partitions = (int)hp.i64();
                            else hp.skip();
                        }
                    } else if (pt.field == 2) { pt.skip(); op->num_partitions = 1; }
                    else throw Unsupported("range / round-robin partitioning is outside the GPU hot path");
                }
            } else b.skip();
        }
        for (auto& e : op->hash_exprs) resolve(*e, child_schema());
        if (op->num_partitions <= 0) throw PlanError("shuffle writer without partitions");
        op->schema = child_schema();
        have = true;
        break;
    }
    default:
        throw Unsupported("operator field " + std::to_string(f) + " is outside the GPU hot path");
    }
    if (!have) throw PlanError("operator not decoded");
    return op;
}

OperatorP decode_plan(const uint8_t* data, size_t len) {
    try {
        return decode_operator(PbReader(data, len));
    } catch (const PbError& e) {
        throw PlanError(e.what());
    }
}

std::string expr_str(const Expr& e) {
    std::ostringstream o;
    static const char* names[] = {"lit", "col", "unbound", "+", "-", "*", "/", "=", "!=", ">", ">=", "<", "<=", "isnull",
                                  "isnotnull", "and", "or", "not", "cast", "checkoverflow", "neg", "if", "in"};
    o << names[(int)e.kind];
    if (e.kind == ExprKind::Bound) o << e.index;
    o << ":" << e.type.str();
    if (!e.children.empty()) {
        o << "(";
        for (size_t i = 0; i < e.children.size(); i++) o << (i ? "," : "") << expr_str(*e.children[i]);
        o << ")";
    }
    return o.str();
}

} // namespace cb200
