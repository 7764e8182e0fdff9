// parquet_kernels.cu -- Parquet page decode on sm_100a.
//
// Encodings per the Apache Parquet specification (Encodings.md): PLAIN for fixed-width physical types,
// RLE/bit-packed hybrid for dictionary indices (RLE_DICTIONARY: 1 byte bit width, then runs) and for
// definition levels.  The reference reaches the third-party `parquet` crate for this
// (native/core/src/parquet/parquet_exec.rs:139-141); its source is not under the reference tree, so the
// decoders follow the format specification and are checked against pyarrow-written files.
#include "parquet_kernels.h"
#include "device/cb_math.h"

namespace cb200 {
using namespace cb;

// page bytes start at arbitrary offsets: assemble unaligned little-endian words from aligned loads
__device__ __forceinline__ u64 load_u64_unaligned(const u8* p) {
    size_t a = (size_t)p;
    const u64* q = (const u64*)(a & ~(size_t)7);
    int sh = (int)(a & 7) * 8;
    u64 lo = q[0];
    if (sh == 0) return lo;
    u64 hi = q[1];
    return (lo >> sh) | (hi << (64 - sh));
}
__device__ __forceinline__ u32 load_u32_unaligned(const u8* p) {
    size_t a = (size_t)p;
    const u32* q = (const u32*)(a & ~(size_t)3);
    int sh = (int)(a & 3) * 8;
    u32 lo = q[0];
    if (sh == 0) return lo;
    u32 hi = q[1];
    return (lo >> sh) | (hi << (32 - sh));
}

// ---- PLAIN ---------------------------------------------------------------------------------------------------
template <int CONV> __global__ void k_pq_plain(const u8* chunk, const PqPage* pages, int flba_len, u8* out) {
    const PqPage pg = pages[blockIdx.y];
    if (pg.encoding != 0) return; // dictionary-encoded page: decoded by k_pq_rle_decode
    const u8* src = chunk + pg.values_off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pg.num_values; i += gridDim.x * blockDim.x) {
        long long row = pg.dst_row + i;
        if (CONV == PQ_COPY32) ((u32*)out)[row] = load_u32_unaligned(src + (size_t)i * 4);
        else if (CONV == PQ_COPY64) ((u64*)out)[row] = load_u64_unaligned(src + (size_t)i * 8);
        else if (CONV == PQ_I32_TO_I64) ((i64*)out)[row] = (i64)(i32)load_u32_unaligned(src + (size_t)i * 4);
        else if (CONV == PQ_I64_TO_I128) ((i128*)out)[row] = i128_from_i64((i64)load_u64_unaligned(src + (size_t)i * 8));
        else if (CONV == PQ_I32_TO_I128) ((i128*)out)[row] = i128_from_i64((i64)(i32)load_u32_unaligned(src + (size_t)i * 4));
        else { // FIXED_LEN_BYTE_ARRAY: big-endian two's complement of flba_len bytes
            const u8* b = src + (size_t)i * flba_len;
            u64 hi = (b[0] & 0x80) ? ~0ull : 0ull, lo = hi;
            for (int k = 0; k < flba_len; k++) {
                hi = (hi << 8) | (lo >> 56);
                lo = (lo << 8) | b[k];
            }
            if (CONV == PQ_FLBA_TO_I64) ((i64*)out)[row] = (i64)lo;
            else ((i128*)out)[row] = mk128(lo, (i64)hi);
        }
    }
}
void launch_pq_plain(const unsigned char* chunk, const PqPage* pages, int n_pages, int conv, int flba_len, void* out, cudaStream_t st) {
    if (n_pages <= 0) return;
    dim3 grid(64, (unsigned)n_pages), block(256);
    switch (conv) {
    case PQ_COPY32: k_pq_plain<PQ_COPY32><<<grid, block, 0, st>>>(chunk, pages, flba_len, (u8*)out); break;
    case PQ_COPY64: k_pq_plain<PQ_COPY64><<<grid, block, 0, st>>>(chunk, pages, flba_len, (u8*)out); break;
    case PQ_I32_TO_I64: k_pq_plain<PQ_I32_TO_I64><<<grid, block, 0, st>>>(chunk, pages, flba_len, (u8*)out); break;
    case PQ_I64_TO_I128: k_pq_plain<PQ_I64_TO_I128><<<grid, block, 0, st>>>(chunk, pages, flba_len, (u8*)out); break;
    case PQ_I32_TO_I128: k_pq_plain<PQ_I32_TO_I128><<<grid, block, 0, st>>>(chunk, pages, flba_len, (u8*)out); break;
    case PQ_FLBA_TO_I64: k_pq_plain<PQ_FLBA_TO_I64><<<grid, block, 0, st>>>(chunk, pages, flba_len, (u8*)out); break;
    default: k_pq_plain<PQ_FLBA_TO_I128><<<grid, block, 0, st>>>(chunk, pages, flba_len, (u8*)out); break;
    }
}

// ---- RLE / bit-packed hybrid ------------------------------------------------------------------------------------------
// walk the run headers of [p, end): calls f(is_bit_packed, count, value, data_ptr); returns values seen
template <typename F> __device__ long long walk_hybrid(const u8* p, const u8* end, int bit_width, long long max_values, F f) {
    long long seen = 0;
    const int vbytes = (bit_width + 7) / 8;
    while (p < end && seen < max_values) {
        u32 header = 0;
        int shift = 0;
        while (p < end) { // ULEB128
            u8 b = *p++;
            header |= (u32)(b & 0x7f) << shift;
            shift += 7;
            if (!(b & 0x80)) break;
        }
        if (header & 1) { // bit-packed: (header >> 1) groups of 8 values
            long long count = (long long)(header >> 1) * 8;
            long long take = count < max_values - seen ? count : max_values - seen;
            f(1, (int)take, 0u, p);
            p += (size_t)(header >> 1) * bit_width;
            seen += take;
        } else {
            long long count = header >> 1;
            u32 v = 0;
            for (int k = 0; k < vbytes && p + k < end; k++) v |= (u32)p[k] << (8 * k);
            p += vbytes;
            long long take = count < max_values - seen ? count : max_values - seen;
            f(0, (int)take, v, p);
            seen += take;
        }
    }
    return seen;
}

__global__ void k_pq_rle_scan(const u8* chunk, const PqPage* pages, int n_pages, PqRun* runs, int* run_counts, int* err) {
    int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= n_pages) return;
    const PqPage pg = pages[pi];
    if (pg.encoding == 0) { run_counts[pi] = 0; return; }
    const u8* p = chunk + pg.values_off;
    const u8* end = p + pg.values_bytes;
    int bw = *p++; // RLE_DICTIONARY: the first byte is the index bit width
    PqRun* out = runs + pg.run_base;
    int n = 0;
    long long row = pg.dst_row;
    long long seen = walk_hybrid(p, end, bw, pg.num_values, [&](int packed, int count, u32 value, const u8* data) {
        if (n < pg.max_runs) {
            PqRun r;
            r.out_row = row;
            r.src_off = (long long)(data - chunk);
            r.count = count;
            r.value = value;
            r.bit_packed = packed;
            r.bit_width = bw;
            out[n] = r;
        }
        n++;
        row += count;
    });
    if (n > pg.max_runs || seen != pg.num_values) atomicOr(err, 1);
    run_counts[pi] = n < pg.max_runs ? n : pg.max_runs;
}
void launch_pq_rle_scan(const unsigned char* chunk, const PqPage* pages, int n_pages, PqRun* runs, int* run_counts, int* err, cudaStream_t st) {
    if (n_pages > 0) k_pq_rle_scan<<<(n_pages + 63) / 64, 64, 0, st>>>(chunk, pages, n_pages, runs, run_counts, err);
}

template <int DW> __device__ __forceinline__ void store_dict(const void* dict, int dict_size, u32 idx, void* out, long long row, int* err) {
    if ((int)idx >= dict_size) { atomicOr(err, 4); idx = 0; }
    if (DW == 4) ((u32*)out)[row] = ((const u32*)dict)[idx];
    else if (DW == 8) ((u64*)out)[row] = ((const u64*)dict)[idx];
    else ((ulonglong2*)out)[row] = ((const ulonglong2*)dict)[idx];
}
// one warp per run; blockIdx.y = page
template <int DW> __global__ void k_pq_rle_decode(const u8* chunk, const PqPage* pages, const PqRun* runs, const int* run_counts, const void* dict_all,
                                                 void* out, int* err) {
    const PqPage pg = pages[blockIdx.y];
    if (pg.encoding == 0) return;
    const void* dict = (const u8*)dict_all + (size_t)pg.dict_off * DW;
    const int dict_size = pg.dict_size;
    const int n_runs = run_counts[blockIdx.y];
    const int lane = threadIdx.x & 31, warps_per_block = blockDim.x >> 5;
    for (int ri = blockIdx.x * warps_per_block + (threadIdx.x >> 5); ri < n_runs; ri += gridDim.x * warps_per_block) {
        const PqRun r = runs[pg.run_base + ri];
        if (!r.bit_packed) {
            for (int i = lane; i < r.count; i += 32) store_dict<DW>(dict, dict_size, r.value, out, r.out_row + i, err);
        } else {
            const u8* src = chunk + r.src_off;
            const int bw = r.bit_width;
            const u32 mask = bw >= 32 ? 0xffffffffu : ((1u << bw) - 1u);
            for (int i = lane; i < r.count; i += 32) {
                long long bit = (long long)i * bw;
                const u8* q = src + (bit >> 3);
                u64 w = 0;
                for (int k = 0; k < 5; k++) w |= (u64)q[k] << (8 * k); // bw <= 32: value spans at most 5 bytes
                u32 idx = (u32)(w >> (bit & 7)) & mask;
                store_dict<DW>(dict, dict_size, idx, out, r.out_row + i, err);
            }
        }
    }
}
void launch_pq_rle_decode(const unsigned char* chunk, const PqPage* pages, int n_pages, const PqRun* runs, const int* run_counts, const void* dict, int dict_width,
                          void* out, int* err, cudaStream_t st) {
    if (n_pages <= 0) return;
    dim3 grid(32, (unsigned)n_pages), block(256);
    if (dict_width == 4) k_pq_rle_decode<4><<<grid, block, 0, st>>>(chunk, pages, runs, run_counts, dict, out, err);
    else if (dict_width == 8) k_pq_rle_decode<8><<<grid, block, 0, st>>>(chunk, pages, runs, run_counts, dict, out, err);
    else k_pq_rle_decode<16><<<grid, block, 0, st>>>(chunk, pages, runs, run_counts, dict, out, err);
}

// definition levels (bit width 1): all levels must be 1 until NULL scatter lands
__global__ void k_pq_check_def(const u8* chunk, const PqPage* pages, int n_pages, int* err) {
    int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= n_pages) return;
    const PqPage pg = pages[pi];
    if (pg.def_bytes <= 0) return;
    const u8* p = chunk + pg.def_off;
    bool bad = false;
    long long seen = walk_hybrid(p, p + pg.def_bytes, 1, pg.num_values, [&](int packed, int count, u32 value, const u8* data) {
        if (!packed) { if (value != 1u) bad = true; }
        else for (int i = 0; i < count; i++) if (!((data[i >> 3] >> (i & 7)) & 1)) { bad = true; break; }
    });
    if (bad) atomicOr(err, 2);
    if (seen != pg.num_values) atomicOr(err, 1);
}
void launch_pq_check_def_levels(const unsigned char* chunk, const PqPage* pages, int n_pages, int* err, cudaStream_t st) {
    if (n_pages > 0) k_pq_check_def<<<(n_pages + 63) / 64, 64, 0, st>>>(chunk, pages, n_pages, err);
}

} // namespace cb200
