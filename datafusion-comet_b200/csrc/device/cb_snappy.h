// cb_snappy.h -- Snappy raw-format element parser shared by the device decompressor (parquet_kernels.cu, one
// warp per page) and the host decoder used for the few pages the host itself must read (string dictionaries).
//
// Format (google/snappy format_description.txt): a ULEB128 uncompressed length, then elements whose tag byte's
// low two bits select  00 literal (len-1 in the upper six bits; 60..63 => that many-59 extra length bytes),
// 01 copy with 11-bit offset (len 4..11), 10 copy with 16-bit offset, 11 copy with 32-bit offset (len 1..64).
// The reference gets this from the `snap` crate through the third-party parquet reader (native/core/Cargo.toml:40).
#ifndef CB_SNAPPY_H
#define CB_SNAPPY_H
#include "cb_math.h"

namespace cb {

struct SnappyElem {
    int kind;      // 0 literal, 1 copy, -1 malformed
    int len;       // bytes produced
    long long src; // literal: input position of the bytes; copy: distance back in the output
    long long next;// input position of the next element
};

// ULEB128 preamble; returns the position after it (or -1) and the declared uncompressed length
CB_HD long long snappy_preamble(const u8* in, long long n, u64& ulen) {
    ulen = 0;
    int shift = 0;
    for (long long p = 0; p < n && shift <= 63; p++) {
        u8 b = in[p];
        ulen |= (u64)(b & 0x7f) << shift;
        shift += 7;
        if (!(b & 0x80)) return p + 1;
    }
    return -1;
}

// parse the element whose tag is at in[pos]; never reads at or beyond in[n]
CB_HD SnappyElem snappy_next(const u8* in, long long n, long long pos) {
    SnappyElem e;
    e.kind = -1; e.len = 0; e.src = 0; e.next = n;
    if (pos >= n) return e;
    const u32 tag = in[pos++];
    const u32 t = tag & 3u;
    if (t == 0) {
        u32 len = tag >> 2;
        if (len >= 60) {
            const int extra = (int)len - 59;
            if (pos + extra > n) return e;
            len = 0;
            for (int k = 0; k < extra; k++) len |= (u32)in[pos + k] << (8 * k);
            pos += extra;
        }
        if (len >= 0x7fffffffu) return e;
        e.kind = 0; e.len = (int)len + 1; e.src = pos; e.next = pos + e.len;
        if (e.next > n) e.kind = -1;
        return e;
    }
    if (t == 1) {
        if (pos + 1 > n) return e;
        e.kind = 1; e.len = (int)((tag >> 2) & 7u) + 4; e.src = (long long)(((tag >> 5) << 8) | in[pos]); e.next = pos + 1;
        return e;
    }
    const int ob = t == 2 ? 2 : 4;
    if (pos + ob > n) return e;
    u32 off = 0;
    for (int k = 0; k < ob; k++) off |= (u32)in[pos + k] << (8 * k);
    e.kind = 1; e.len = (int)(tag >> 2) + 1; e.src = (long long)off; e.next = pos + ob;
    return e;
}

// serial decoder (host use; also the definition the warp kernel is tested against). Returns bytes produced or -1.
CB_HD long long snappy_decode_serial(const u8* in, long long n, u8* out, long long cap) {
    u64 ulen;
    long long pos = snappy_preamble(in, n, ulen);
    if (pos < 0 || ulen > (u64)cap) return -1;
    long long o = 0;
    while (pos < n) {
        SnappyElem e = snappy_next(in, n, pos);
        if (e.kind < 0 || o + e.len > (long long)ulen) return -1;
        if (e.kind == 0) for (int i = 0; i < e.len; i++) out[o + i] = in[e.src + i];
        else {
            if (e.src <= 0 || e.src > o) return -1;
            for (int i = 0; i < e.len; i++) out[o + i] = out[o + i - e.src];
        }
        o += e.len;
        pos = e.next;
    }
    return o == (long long)ulen ? o : -1;
}

} // namespace cb
#endif
