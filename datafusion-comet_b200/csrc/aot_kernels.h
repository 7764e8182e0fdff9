// aot_kernels.h -- host-callable launchers of the plan-independent sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace cb200 {

void launch_bitmap_append(uint32_t* dst, long long dst_off, const uint8_t* src, long long src_off, long long n, cudaStream_t st);
void launch_bytes_to_bitmap(const uint8_t* bytes, long long n, uint32_t* out, cudaStream_t st);
void launch_remap_codes(const void* in, int in_width, long long n, const int* table, int table_len, int* out, cudaStream_t st);

enum { CB_DICT_FULL = 1, CB_DICT_COLLISION = 2 };
struct StringDictDev {
    unsigned long long* tags; // [capacity] 0 = empty
    int* slot_code;           // [capacity]
    long long capacity;       // power of two
    int* n_codes;             // running number of codes
    int max_codes;
    long long* code_off;      // [max_codes]
    int* code_len;            // [max_codes]
    unsigned char* bytes;     // string storage
    long long bytes_cap;
    unsigned long long* bytes_used;
    int* err;
};
void launch_dict_encode(const StringDictDev& d, const int* offsets, const unsigned char* chars, const unsigned char* validity, long long n,
                        int* row_slot, int* codes, cudaStream_t st);

// stream compaction (hash-aggregate results): per-1024-row-block counts + exclusive scan, then one scatter per column
void launch_key_presence(const unsigned long long* keys, long long n, unsigned char* present, cudaStream_t st);
void launch_compact_plan(const unsigned char* present, long long n, int* counts, long long* offsets, long long* total, cudaStream_t st);
void launch_compact_scatter(const unsigned char* present, long long n, const long long* offsets, const void* in, int width, void* out, cudaStream_t st);

} // namespace cb200
