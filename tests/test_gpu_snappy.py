"""GPU parity for the device Snappy decompressor: index pass + 64 KB segments + serial fallback, against the
oracle's decoder (oracle/parquet_oracle.py snappy_decompress, itself pinned to pyarrow on CPU) and against the original bytes."""
import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import comet_b200
    return comet_b200


def varint(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def literal(data):
    n = len(data) - 1
    if n < 60:
        return bytes([n << 2]) + data
    nb = (n.bit_length() + 7) // 8
    return bytes([(59 + nb) << 2]) + n.to_bytes(nb, "little") + data


def copy2(offset, length):                     # 16-bit offset form, 1..64 bytes
    return bytes([((length - 1) << 2) | 2]) + offset.to_bytes(2, "little")


def datasets():
    rng = np.random.default_rng(3)
    yield "int64 decimals (literal + copy per value)", (rng.integers(90000, 10**7, 200_000).astype(np.int64)).tobytes()
    yield "random bytes (64 KB literals)", rng.integers(0, 256, 300_001, dtype=np.uint8).tobytes()
    yield "long runs", (b"abc" * 50_000 + bytes(100_000) + b"xyz" * 33_333)
    yield "bit-packed-like", np.packbits(rng.integers(0, 2, 2_000_000, dtype=np.uint8)).tobytes()
    yield "tiny", b"hello snappy"
    yield "exactly one segment", rng.integers(0, 4, 65536, dtype=np.uint8).tobytes()
    yield "exactly two segments", rng.integers(0, 4, 131072, dtype=np.uint8).tobytes()


def test_stock_streams_take_the_segmented_path(cb):
    from oracle import parquet_oracle as O
    for name, raw in datasets():
        comp = pa.compress(raw, codec="snappy", asbytes=True)
        assert O.snappy_decompress(comp) == raw, name
        got, path = cb.native.snappy_decompress(comp, len(raw))
        assert got == raw, name
        assert path == 0, name                   # the stock compressor never crosses a 64 KB output boundary


def test_element_across_a_boundary_goes_to_the_serial_decoder(cb):
    from oracle import parquet_oracle as O
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, 100_000, dtype=np.uint8).tobytes()
    stream = varint(len(a) + 40) + literal(a) + copy2(1000, 40)     # one 100 000-byte literal straddles the 64 KB boundary
    exp = O.snappy_decompress(stream)
    assert exp == a + a[-1000:-960]
    got, path = cb.native.snappy_decompress(stream, len(exp))
    assert got == exp and path == 1


def test_reference_into_an_earlier_segment_goes_to_the_serial_decoder(cb):
    from oracle import parquet_oracle as O
    rng = np.random.default_rng(2)
    a = rng.integers(0, 256, 65536, dtype=np.uint8).tobytes()
    b = rng.integers(0, 256, 500, dtype=np.uint8).tobytes()
    # segment 1 starts on an element boundary (regular shape) but its first copy reaches 100 bytes back into segment 0
    stream = varint(65536 + 64 + 500 + 30) + literal(a) + copy2(100, 64) + literal(b) + copy2(64 + 500, 30)
    exp = O.snappy_decompress(stream)
    got, path = cb.native.snappy_decompress(stream, len(exp))
    assert got == exp and path == 1


def test_empty_and_malformed(cb):
    got, path = cb.native.snappy_decompress(varint(0), 0)
    assert got == b"" and path == 0
    raw = bytes(range(256)) * 600
    comp = pa.compress(raw, codec="snappy", asbytes=True)
    body = comp[len(varint(len(raw))):]
    cases = [(comp[:-3], len(raw)),                       # truncated input
             (varint(len(raw) + 5) + body, len(raw) + 5),  # declares more than the elements produce
             (varint(len(raw) - 5) + body, len(raw) - 5),  # ... and less
             (comp, len(raw) + 1),                         # the caller's size disagrees with the preamble
             (varint(20) + copy2(5, 20), 20)]              # a copy before anything was written
    for stream, n in cases:
        with pytest.raises(cb.native.CometB200Error):
            cb.native.snappy_decompress(stream, n)
