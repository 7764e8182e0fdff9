"""CPU-only: the hand-written Parquet footer / page-header parser (Thrift compact) agrees with pyarrow's reader."""
import struct

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


@pytest.mark.parametrize("compression,version,dictionary", [("NONE", "1.0", False), ("SNAPPY", "1.0", True), ("SNAPPY", "2.0", False), ("NONE", "2.0", True)])
def test_parquet_oracle_matches_pyarrow(tmp_path, compression, version, dictionary):
    """oracle/parquet_oracle.py (page headers, Snappy, RLE hybrid levels, PLAIN / dictionary values) against pyarrow's reader."""
    import decimal
    import pyarrow as pa
    import pyarrow.parquet as pq
    from oracle import parquet_oracle as po
    rng = np.random.default_rng(3)
    n = 12_000
    ctx = decimal.Context(prec=60)
    m = [rng.random(n) < 0.2 for _ in range(5)]
    m[0][:700] = True
    i64 = rng.integers(-2**60, 2**60, n)
    low = rng.integers(0, 30, n).astype(np.int32)
    f64 = rng.standard_normal(n)
    d30 = [int(a) * 10**10 + int(b) for a, b in zip(rng.integers(-10**15, 10**15, n), rng.integers(0, 10**10, n))]
    words = np.array(["AIR", "MAIL", "SHIP", "", "TRUCK"])[rng.integers(0, 5, n)]
    tbl = pa.table({"i64": pa.array(i64, mask=m[0]), "low": pa.array(low, mask=m[1]), "f64": pa.array(f64, mask=m[2]),
                    "d30": pa.array([None if mm else decimal.Decimal(v).scaleb(-4, context=ctx) for v, mm in zip(d30, m[3])], type=pa.decimal128(30, 4)),
                    "word": pa.array(words.tolist(), mask=m[4]), "req": pa.array(i64)})
    path = str(tmp_path / "o.parquet")
    pq.write_table(tbl, path, row_group_size=5000, compression=compression, use_dictionary=True if dictionary else ["word"], data_page_version=version, data_page_size=4096)
    raw = open(path, "rb").read()
    md = pq.ParquetFile(path).metadata
    for ci, name in enumerate(tbl.column_names):
        got_v, got_ok = [], []
        for rg in range(md.num_row_groups):
            c = md.row_group(rg).column(ci)
            start = c.dictionary_page_offset if c.has_dictionary_page and c.dictionary_page_offset else c.data_page_offset
            start = min(start, c.data_page_offset)
            tl = md.schema.column(ci).length if c.physical_type == "FIXED_LEN_BYTE_ARRAY" else 0
            v, ok = po.decode_chunk(raw, start, c.total_compressed_size, c.num_values, c.physical_type, c.compression, True, tl)
            got_v.append(v)
            got_ok.append(ok)
        v, ok = np.concatenate(got_v), np.concatenate(got_ok)
        want = tbl.column(name).to_pylist()
        assert [w is not None for w in want] == ok.tolist(), name
        for g, w in zip(v[ok].tolist(), [w for w in want if w is not None]):
            if name == "d30":
                assert g == int(w.scaleb(4)), name
            elif name == "word":
                assert g.decode() == w, name
            elif name == "f64":
                assert struct.pack("<d", g) == struct.pack("<d", w), name
            else:
                assert g == w, name


def test_host_codecs_decompress_what_pyarrow_compressed():
    """ZSTD / LZ4_RAW / LZ4 (Hadoop framing) / GZIP pages are decompressed on the host (csrc/host_codecs.cpp: libzstd and liblz4 through
    dlopen, zlib linked) before the device decodes them; checked here against pyarrow's compressors without a GPU."""
    import ctypes as C
    import os
    import pyarrow as pa
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "datafusion-comet_b200", "libcomet_b200.so"))
    f = getattr(lib, "_ZN5cb20015host_decompressEiPKhmPhm")       # cb200::host_decompress(int, const uint8_t*, size_t, uint8_t*, size_t)
    f.restype = None
    f.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    rng = np.random.default_rng(0)
    raw = rng.integers(0, 1000, 300_000).astype(np.int64).tobytes() + bytes(50_000) + rng.integers(0, 256, 70_000, dtype=np.uint8).tobytes()
    for codec_id, name in ((6, "zstd"), (7, "lz4_raw"), (5, "lz4_hadoop"), (5, "lz4_raw"), (2, "gzip")):
        if name == "lz4_hadoop":     # the deprecated LZ4 codec as Hadoop frames it: [u32 BE uncompressed][u32 BE compressed][raw block], here two blocks
            half = len(raw) // 2
            comp = b""
            for part in (raw[:half], raw[half:]):
                blk = pa.compress(part, codec="lz4_raw", asbytes=True)
                comp += len(part).to_bytes(4, "big") + len(blk).to_bytes(4, "big") + blk
        else:
            comp = pa.compress(raw, codec=name, asbytes=True)
        out = C.create_string_buffer(len(raw))
        f(codec_id, comp, len(comp), out, len(raw))
        assert out.raw == raw, name
