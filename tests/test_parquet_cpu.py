"""CPU-only: the hand-written Parquet footer / page-header parser (Thrift compact) agrees with pyarrow's reader."""
import numpy as np
import pyarrow.parquet as pq
import pytest


@pytest.fixture(scope="module")
def cb():
    import comet_b200
    return comet_b200


@pytest.mark.parametrize("as_int", [True, False])
def test_footer_matches_pyarrow(cb, tmp_path, as_int):
    t = cb.tpch
    cols = t.gen_lineitem(50_000, seed=4)
    path = str(tmp_path / "li.parquet")
    t.write_lineitem_parquet(cols, path, "dec", row_group_size=16_384, decimal_as_int=as_int)
    mine = cb.native.parquet_describe(path)
    ref = pq.ParquetFile(path).metadata
    assert mine["num_rows"] == ref.num_rows == 50_000
    assert len(mine["row_groups"]) == ref.num_row_groups
    names = [c["name"] for c in mine["columns"]]
    assert names == t.Q1_COLUMNS
    phys = {"INT32": 1, "INT64": 2, "BYTE_ARRAY": 6, "FIXED_LEN_BYTE_ARRAY": 7, "DOUBLE": 5}
    for i, c in enumerate(mine["columns"]):
        rc = ref.schema.column(i)
        assert c["type"] == phys[rc.physical_type]
    assert mine["columns"][0]["type"] == (2 if as_int else 7) and mine["columns"][0]["precision"] == 12 and mine["columns"][0]["scale"] == 2
    for g in range(ref.num_row_groups):
        rg = ref.row_group(g)
        assert mine["row_groups"][g]["num_rows"] == rg.num_rows
        for c in range(rg.num_columns):
            a, b = mine["row_groups"][g]["columns"][c], rg.column(c)
            assert a["num_values"] == b.num_values and a["total_compressed"] == b.total_compressed_size
            assert a["data_page_offset"] == b.data_page_offset
            assert a["codec"] == 0
            if b.has_dictionary_page:
                assert a["dictionary_page_offset"] == b.dictionary_page_offset


def test_memory_file_registration(cb, tmp_path):
    t = cb.tpch
    path = str(tmp_path / "li.parquet")
    t.write_lineitem_parquet(t.gen_lineitem(1000, seed=1), path, "f64")
    image = np.fromfile(path, dtype=np.uint8)
    url = cb.native.register_memory_file("unit-test-file", image)
    assert url == "memory://unit-test-file"
    assert cb.native.parquet_describe(url) == cb.native.parquet_describe(path)
    cb.native.register_memory_file("unit-test-file", None)
    with pytest.raises(cb.native.CometB200Error):
        cb.native.parquet_describe(url)


def test_native_scan_plan_supported(cb):
    t = cb.tpch
    plan = t.q1_partial_plan("dec", scan=t.q1_native_scan("dec", ["file:///tmp/none.parquet"]))
    ok, why = cb.native.supports(plan)
    assert ok, why
