"""Round-1 design study of the segmented Snappy page decoder (not product code; runs on the CPU, no GPU needed).
The decoder was built in round 2 (csrc/parquet_kernels.cu: k_pq_snappy_index / k_pq_snappy_seg); the algorithm as shipped is modelled and
checked in tests/test_snappy_index_model.py.  Kept for the stream statistics it prints (elements per window, cross-segment references).


Today's kernel (parquet_kernels.cu k_pq_snappy) lets lane 0 parse one element per loop trip: ~115 dependent instructions per
element at ~4.5 cycles each.  The plan for the next round is a two-pass decode per page:

  pass A  (one warp per page)  lane-parallel speculative parse: each lane treats input byte pos+lane as the start of an
          element and computes where the next element would start; the true chain is then walked through the lanes' answers
          (one shuffle per element instead of ~115 instructions).  Every SEG elements it records a checkpoint (input position,
          output position).
  pass B  (one warp per checkpoint segment)  decodes its segment exactly like today's kernel, but many segments of a page
          run concurrently.  Measured below on Parquet-like pages: the stock compressor works on independent 64 KB fragments,
          so NO back-reference crosses a 64 KB output boundary (median distance 880 B, 99% < 40 KB) -- checkpoints at 64 KB
          output boundaries give 16 fully independent segments per 1 MB page; a stream that does reference across a boundary
          (legal, never produced by snappy itself) is detected in pass B and sent to the serial kernel.

This file checks the part that can be checked without a GPU: that the speculative parse + chain walk reproduces the element
boundaries of real Snappy streams (pyarrow-compressed Parquet-like pages), and it reports the statistics that size the design
(elements per 32-byte window, cross-segment references)."""
import sys

import numpy as np
import pyarrow as pa


def elem_at(src, p):
    """(advance to next element, output bytes, copy offset or 0) for an element starting at src[p]; None if it runs off the end"""
    if p >= len(src):
        return None
    tag = src[p]
    t = tag & 3
    if t == 0:
        ln, hdr = tag >> 2, 1
        if ln >= 60:
            extra = ln - 59
            if p + 1 + extra > len(src):
                return None
            ln = int.from_bytes(src[p + 1:p + 1 + extra], "little")
            hdr += extra
        ln += 1
        return hdr + ln, ln, 0
    if t == 1:
        return (2, ((tag >> 2) & 7) + 4, ((tag >> 5) << 8) | src[p + 1]) if p + 1 < len(src) else None
    nb = 2 if t == 2 else 4
    if p + 1 + nb > len(src):
        return None
    return 1 + nb, (tag >> 2) + 1, int.from_bytes(src[p + 1:p + 1 + nb], "little")


def preamble(src):
    n = shift = pos = 0
    while True:
        c = src[pos]
        pos += 1
        n |= (c & 0x7F) << shift
        shift += 7
        if not c & 0x80:
            return n, pos


def serial_boundaries(src):
    n, pos = preamble(src)
    out, b = 0, []
    while pos < len(src):
        adv, ln, off = elem_at(src, pos)
        b.append((pos, out, ln, off))
        pos += adv
        out += ln
    assert out == n
    return b


def lane_parallel_boundaries(src, lanes=32):
    """pass A: per window of `lanes` bytes every lane parses speculatively; the chain walk only follows `next` pointers"""
    n, pos = preamble(src)
    out, b, windows, hops = 0, [], 0, 0
    while pos < len(src):
        spec = [elem_at(src, pos + i) for i in range(lanes)]          # all lanes at once on the GPU
        windows += 1
        s = 0
        while s < lanes and spec[s] is not None:                      # the walk: one shuffle per hop
            adv, ln, off = spec[s]
            b.append((pos + s, out, ln, off))
            out += ln
            s += adv
            hops += 1
            if pos + s >= len(src):
                break
        pos += s
    assert out == n, (out, n)
    return b, windows, hops


def page_like_payloads(rng):
    n = 131072
    price = (rng.integers(1, 51, n) * rng.integers(90000, 210001, n)).astype(np.int64)       # PLAIN INT64 page (the slow case today)
    qty = rng.integers(1, 51, n).astype(np.int64) * 100
    codes6 = np.packbits(rng.integers(0, 2, n * 6).astype(np.uint8))                          # bit-packed dictionary indices
    flags = np.repeat(rng.integers(0, 3, n // 64).astype(np.uint8), 64)
    return {"plain_int64_price": price.tobytes(), "plain_int64_qty": qty.tobytes(), "bitpacked_indices": codes6.tobytes(), "rle_like_flags": flags.tobytes()}


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    seg = 256
    for name, raw in page_like_payloads(rng).items():
        comp = pa.compress(raw, codec="snappy", asbytes=True)
        ser = serial_boundaries(comp)
        par, windows, hops = lane_parallel_boundaries(comp)
        assert ser == par, name
        lens = np.array([e[2] for e in ser])
        offs = np.array([e[3] for e in ser])
        outs = np.array([e[1] for e in ser])
        cross = 0
        for k in range(0, len(ser), seg):                                                     # back-references that leave their segment
            seg_start = outs[k]
            sl = slice(k, min(k + seg, len(ser)))
            cross += int(((offs[sl] > 0) & (outs[sl] - offs[sl] < seg_start)).sum())
        print(f"{name:20s} raw {len(raw):8d} B  snappy {len(comp):8d} B  elements {len(ser):7d}  bytes/elem out {len(raw) / len(ser):7.1f}  "
              f"elements per 32-byte window {hops / windows:5.2f}  copies {int((offs > 0).sum()):7d}  cross-segment copies (SEG={seg}) {cross}")
    print("speculative parse + chain walk reproduces the serial element boundaries on every payload")
