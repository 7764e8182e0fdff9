// parquet_kernels.h -- device-side Parquet page decode (PLAIN, RLE_DICTIONARY, definition levels).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace cb200 {

struct PqPage { // one data page of a column chunk already resident on the device
    long long values_off;   // byte offset of the encoded VALUES (after the levels) inside the chunk buffer
    int values_bytes;
    long long def_off;      // byte offset of the RLE-encoded definition levels (0 length: column is required)
    int def_bytes;
    long long dst_row;      // first output row of this page
    int num_values;
    int encoding;           // 0 PLAIN, 8 RLE_DICTIONARY (2 PLAIN_DICTIONARY is the same on the wire)
    long long run_base;     // first entry of this page in the run table (RLE pages)
    int max_runs;           // capacity reserved for it
    long long dict_off;     // element offset of this page's dictionary inside the column's combined dictionary buffer
    int dict_size;
};

struct PqRun {              // one run of the RLE / bit-packed hybrid
    long long out_row;      // absolute output row of the run's first value
    long long src_off;      // byte offset (chunk buffer) of the packed data (bit-packed) -- unused for RLE runs
    int count;              // values in the run
    unsigned value;         // RLE runs: the repeated value
    int bit_packed;
    int bit_width;
};

enum PqConv { PQ_COPY32, PQ_COPY64, PQ_I32_TO_I64, PQ_FLBA_TO_I64, PQ_FLBA_TO_I128, PQ_I64_TO_I128, PQ_I32_TO_I128 };

// PLAIN fixed-width pages -> output column (element width given by the conversion)
void launch_pq_plain(const unsigned char* chunk, const PqPage* pages_dev, int n_pages, int conv, int flba_len, void* out, cudaStream_t st);
// RLE_DICTIONARY pages: (1) scan run headers, one thread per page
void launch_pq_rle_scan(const unsigned char* chunk, const PqPage* pages_dev, int n_pages, PqRun* runs, int* run_counts, int* err, cudaStream_t st);
// (2) decode runs (warp per run) and gather through the dictionary: dict_width 4/8/16 bytes per entry
void launch_pq_rle_decode(const unsigned char* chunk, const PqPage* pages_dev, int n_pages, const PqRun* runs, const int* run_counts, const void* dict,
                          int dict_width, void* out, int* err, cudaStream_t st);
// definition levels of flat optional columns (max level 1): verify "no NULLs" (sets err bit 1 if a 0 level appears)
void launch_pq_check_def_levels(const unsigned char* chunk, const PqPage* pages_dev, int n_pages, int* err, cudaStream_t st);

} // namespace cb200
