"""CPU model of the device Snappy index pass (csrc/parquet_kernels.cu k_pq_snappy_index) -- no GPU needed.

The kernel treats every byte of a 512-byte input window as a candidate element start, lets each of 32 lanes collapse the chains
inside its 16-position block with one backward sweep, hops the true chain from block to block, and records the input position of
every 64 KB boundary of the OUTPUT.  This restates exactly that in Python and checks, on streams pyarrow's (stock) Snappy produces
and on hand-made ones, that (a) the block-collapse + hop walk lands on the same element boundaries as a serial parse, (b) the stock
compressor never lets an element straddle a 64 KB output boundary (the assumption the segmented decoder rests on; streams that do
are detected and go to the serial kernel), (c) no back-reference of a stock stream reaches into an earlier 64 KB segment."""
import numpy as np
import pyarrow as pa

SEG, W, BLK = 65536, 512, 16
INVALID = 0xFFFFFFFF


def preamble(src):
    n = shift = pos = 0
    while True:
        c = src[pos]
        pos += 1
        n |= (c & 0x7F) << shift
        shift += 7
        if not c & 0x80:
            return n, pos


def elem_len(src, p):
    """(bytes to the next element, bytes produced, copy distance or 0); adv 0 = malformed / runs off the end"""
    if p >= len(src):
        return 0, 0, 0
    tag = src[p]
    t = tag & 3
    if t == 0:
        ln, hdr = tag >> 2, 1
        if ln >= 60:
            extra = ln - 59
            if p + 1 + extra > len(src):
                return 0, 0, 0
            ln = int.from_bytes(src[p + 1:p + 1 + extra], "little")
            hdr += extra
        ln += 1
        return (hdr + ln, ln, 0) if p + hdr + ln <= len(src) else (0, 0, 0)
    nb = 1 if t == 1 else 2 if t == 2 else 4
    if p + 1 + nb > len(src):
        return 0, 0, 0
    if t == 1:
        return 2, ((tag >> 2) & 7) + 4, ((tag >> 5) << 8) | src[p + 1]
    return 1 + nb, (tag >> 2) + 1, int.from_bytes(src[p + 1:p + 1 + nb], "little")


def serial_index(src):
    """reference walk: (checkpoints, irregular?, cross-segment reference?)"""
    body, pos = preamble(src)
    o, bnd, ck, irregular, cross = 0, 0, [], False, False
    while pos < len(src):
        if o == bnd:
            ck.append(pos)
            bnd += SEG
        elif o > bnd:
            irregular = True
            break
        adv, out, dist = elem_len(src, pos)
        assert adv > 0
        if dist and dist > o - (o // SEG) * SEG:
            cross = True
        o += out
        pos += adv
    return ck, irregular, cross, body


def model_index(src):
    """the kernel's algorithm: windows of W - 16 positions, blocks of BLK positions collapsed backwards, hops between blocks"""
    body, pos = preamble(src)
    n = len(src)
    o, bnd, ck = 0, 0, []
    while pos < n:
        if o == bnd:
            ck.append(pos)
            bnd += SEG
        elif o > bnd:
            return ck, True
        L = min(W - 16, n - pos)
        nxt, sm = [INVALID] * W, [0] * W
        for lane in range(32):                                   # all lanes at once on the GPU
            b0, bend = lane * BLK, lane * BLK + BLK
            for k in range(BLK - 1, -1, -1):
                i = b0 + k
                if i >= L:
                    continue
                adv, out, _ = elem_len(src, pos + i)
                if adv == 0:
                    continue
                t = i + adv
                if t < bend and t < L:
                    nxt[i], sm[i] = nxt[t], out + sm[t]
                else:
                    nxt[i], sm[i] = t, out
        E = S = 0
        while E < L:
            assert nxt[E] != INVALID
            S += sm[E]
            E = nxt[E]
        if o + S <= bnd:
            o += S
            pos += E
            continue
        while pos < n and o < bnd:                               # a boundary lies inside this window: element by element
            adv, out, _ = elem_len(src, pos)
            assert adv > 0
            o += out
            pos += adv
    assert o == body
    return ck, False


def payloads():
    rng = np.random.default_rng(1)
    n = 200_000
    yield "plain int64 prices", (rng.integers(1, 51, n) * rng.integers(90000, 210001, n)).astype(np.int64).tobytes()
    yield "random bytes", rng.integers(0, 256, 300_001, dtype=np.uint8).tobytes()
    yield "bit-packed indices", np.packbits(rng.integers(0, 2, 1_500_000).astype(np.uint8)).tobytes()
    yield "runs", np.repeat(rng.integers(0, 3, 4000).astype(np.uint8), 64).tobytes()
    yield "exactly two segments", rng.integers(0, 4, 131072, dtype=np.uint8).tobytes()


def test_block_collapse_and_hops_find_the_serial_checkpoints():
    for name, raw in payloads():
        comp = pa.compress(raw, codec="snappy", asbytes=True)
        ck, irregular, cross, body = serial_index(comp)
        assert body == len(raw) and not irregular, name          # (b) the stock compressor never straddles a 64 KB output boundary
        assert not cross, name                                   # (c) ... and never references across one
        assert len(ck) == (len(raw) + SEG - 1) // SEG, name
        got, irr = model_index(comp)
        assert not irr and got == ck, name                       # (a)


def test_hand_made_irregular_stream_is_detected():
    def varint(v):
        out = bytearray()
        while True:
            b = v & 0x7F
            v >>= 7
            out.append(b | (0x80 if v else 0))
            if not v:
                return bytes(out)
    a = bytes(np.random.default_rng(2).integers(0, 256, 100_000, dtype=np.uint8))
    lit = bytes([(59 + 3) << 2]) + (len(a) - 1).to_bytes(3, "little") + a
    stream = varint(len(a) + 40) + lit + bytes([((40 - 1) << 2) | 2]) + (1000).to_bytes(2, "little")
    _, irregular, _, _ = serial_index(stream)
    assert irregular
    _, irr = model_index(stream)
    assert irr
