"""The hand-written plan encoder/decoder must use the reference's field numbers."""
import os
import re

import pytest

REF = "/root/reference/native/proto/src/proto"


def parse_proto(path):
    """{message_or_oneof_scope: {field_name: number}} with nested messages flattened by simple name."""
    txt = re.sub(r"//[^\n]*", "", open(path).read())
    out = {}
    stack = []
    for tok in re.finditer(r"(message|enum|oneof)\s+(\w+)\s*\{|\}|(?:repeated\s+|optional\s+)?[\w.<>, ]+?\s+(\w+)\s*=\s*(\d+)\s*(?:\[[^\]]*\])?;|(\w+)\s*=\s*(-?\d+)\s*;", txt):
        if tok.group(1):
            stack.append((tok.group(1), tok.group(2)))
            if tok.group(1) != "oneof":
                out.setdefault(tok.group(2), {})
        elif tok.group(0) == "}":
            if stack:
                stack.pop()
        else:
            name, num = (tok.group(3), tok.group(4)) if tok.group(3) else (tok.group(5), tok.group(6))
            owner = next((n for k, n in reversed(stack) if k != "oneof"), None)
            if owner:
                out[owner][name] = int(num)
    return out


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_field_numbers_match_reference():
    from comet_b200 import proto as P
    expr = parse_proto(os.path.join(REF, "expr.proto"))
    op = parse_proto(os.path.join(REF, "operator.proto"))
    types = parse_proto(os.path.join(REF, "types.proto"))
    lit = parse_proto(os.path.join(REF, "literal.proto"))
    part = parse_proto(os.path.join(REF, "partitioning.proto"))
    for k, v in P.EXPR_FIELD.items():
        assert expr["Expr"][k] == v, k
    for k, v in P.AGG_FIELD.items():
        assert expr["AggExpr"][k] == v, k
    for k, v in P.OP_FIELD.items():
        assert op["Operator"][k] == v, k
    for k, v in P.DATA_TYPE_ID.items():
        assert types["DataTypeId"][k] == v, k
    for k, v in P.LITERAL_FIELD.items():
        assert lit["Literal"][k] == v, k
    assert expr["MathExpr"] == {"left": 1, "right": 2, "return_type": 4, "eval_mode": 5, "check_divide_overflow": 6}
    assert expr["Sum"] == {"child": 1, "datatype": 2, "eval_mode": 3}
    assert expr["Avg"] == {"child": 1, "datatype": 2, "sum_datatype": 3, "eval_mode": 4}
    assert expr["CheckOverflow"] == {"child": 1, "datatype": 2, "fail_on_error": 3}
    assert expr["BoundReference"] == {"index": 1, "datatype": 2}
    assert expr["AggExpr"]["filter"] == 89
    assert op["HashAggregate"]["grouping_exprs"] == 1 and op["HashAggregate"]["agg_exprs"] == 2 and op["HashAggregate"]["mode"] == 5
    assert op["Operator"]["children"] == 1 and op["Operator"]["plan_id"] == 2
    assert op["Scan"] == {"fields": 1, "source": 2}
    assert op["NativeScanCommon"]["required_schema"] == 1 and op["NativeScanCommon"]["projection_vector"] == 5
    assert op["NativeScan"] == {"common": 1, "file_partition": 2}
    assert op["AggregateMode"] == {"Partial": 0, "Final": 1, "PartialMerge": 2}
    assert expr["EvalMode"] == {"LEGACY": 0, "TRY": 1, "ANSI": 2}
    assert part["HashPartition"] == {"hash_expression": 1, "num_partitions": 2}
    assert types["DecimalInfo"] == {"precision": 1, "scale": 2}


def test_varint_and_literal_encoding_roundtrip_through_the_decoder():
    """Negative ints are 10-byte varints, decimals big-endian two's complement (planner.rs:544-548): the C++
    decoder must read back what the encoder wrote -- checked through generated kernel source."""
    from comet_b200 import native, proto as P
    sc = P.scan([P.DECIMAL(12, 2), P.INT32])
    pred = P.and_(P.gt(P.bound(0, P.DECIMAL(12, 2)), P.literal(-12345, P.DECIMAL(12, 2))), P.gt(P.bound(1, P.INT32), P.literal(-7, P.INT32)))
    plan = P.projection(P.filter_(sc, pred), [P.bound(1, P.INT32)])
    src = native.kernel_source(plan, 1)      # the predicates live in the count pass (kernel 1); pass 2 takes its keep bits
    assert "((cb::i32)-7)" in src
    assert str((-12345) & ((1 << 64) - 1)) + "ull" in src  # sign-extended low limb of the decimal literal
