// exec.h -- executor: pull-based operator tree over device batches.
//
// Mirrors what the reference builds in PhysicalPlanner::create_plan (native/core/src/execution/
// planner.rs:1211) but with pipeline fusion: every maximal chain Scan -> (Filter|Projection)* ->
// {output | HashAggregate} becomes ONE JIT-specialised kernel launch per device chunk.
#pragma once
#include "arrow_abi.h"
#include "codegen.h"
#include "jit.h"
#include "plan.h"

#include <cuda_runtime.h>
#include <memory>
#include <string>
#include <vector>

namespace cb200 {

struct ExecError : std::runtime_error {
    int code;
    std::string error_class;
    ExecError(int c, const std::string& cls, const std::string& msg) : std::runtime_error(msg), code(c), error_class(cls) {}
};

void cuda_check(cudaError_t e, const char* what);

// CB200_TRACE=1: wall-clock spans to stderr (the reference's spark.comet.tracing.enabled analogue,
// native/common/src/tracing.rs:27-96)
struct TraceSpan {
    const char* name;
    double t0;
    explicit TraceSpan(const char* n);
    ~TraceSpan();
};
bool trace_on();
double now_ms();
void set_alloc_stream(cudaStream_t s); // stream used by DeviceBuf allocations made on this thread
size_t release_cached_device_memory();  // frees the library's recycled large blocks on the current device; returns the bytes released

struct DeviceBuf {
    void* ptr = nullptr;
    size_t bytes = 0;
    bool owned = true;
    cudaStream_t stream = nullptr;
    std::shared_ptr<void> owner;               // non-owning views: the block this pointer lives in (kept alive with the view)
    DeviceBuf() {}
    DeviceBuf(size_t n);                       // cudaMalloc, padded
    DeviceBuf(void* p, size_t n) : ptr(p), bytes(n), owned(false) {}
    ~DeviceBuf();
    DeviceBuf(const DeviceBuf&) = delete;
    DeviceBuf& operator=(const DeviceBuf&) = delete;
};
using DeviceBufP = std::shared_ptr<DeviceBuf>;

struct Dictionary { // string dictionary of a key column (host copy; tiny)
    std::vector<std::string> values;
};
using DictionaryP = std::shared_ptr<Dictionary>;

struct Column {
    DType type;
    Phys phys = Phys::I32;          // physical encoding of `data` (device) -- see codegen.h
    bool is_dict = false;           // data holds dictionary codes (type = String)
    DictionaryP dict;
    DeviceBufP data, validity;      // device (validity: Arrow bitmap) ...
    DeviceBufP offsets, chars;      // ... device Utf8 (offsets int32[n+1], chars)
    DeviceBufP valid_bytes;         // optional byte-per-row validity (exchange-friendly form; see PartitionNode / cb200_table_add_column_bytes)
    DeviceBufP bool_bytes;          // optional byte-per-row form of a boolean column
    int64_t null_count = 0;
    // host-resident alternative (small aggregate results)
    bool on_host = false;
    std::vector<uint8_t> h_data;        // fixed-width values or chars
    std::vector<uint8_t> h_valid;       // one byte per row; empty = all valid
    std::vector<int32_t> h_offsets;     // Utf8
};

struct Batch {
    int64_t n_rows = 0;
    std::vector<Column> cols;
};

struct ExecContext {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    int64_t chunk_rows = 1ll << 26;
    int hash_threads = 512;   // consumer threads per CTA of the hash-aggregate kernel (tuning knob)
    // Partial hash aggregates over inputs of at least this many rows sample whether equal keys are adjacent and, if so, emit one state
    // row per run instead of building a key table (spark.comet.b200.streamAgg.minRows; -1 disables, 0 = always sample)
    int64_t stream_agg_min_rows = 4 << 20;
    double stream_agg_max_ratio = 0.5; // state rows per input row above which the key table is used (spark.comet.b200.streamAgg.maxRatio)
    int batch_size = 8192;
    int* d_err = nullptr;   // device error flags
    int* h_err = nullptr;   // pinned host mirror
    int64_t kernel_launches = 0;
    std::string last_kernel_key;
    // measurement (bench.py / cb200_plan_stats): CUDA events around each fused pipeline kernel, on the
    // stream the kernel is launched on
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_pending = false;
    double pipeline_ms = 0;       // sum of fused-pipeline kernel durations
    int64_t pipeline_launches = 0;
    int64_t pipeline_rows = 0;    // rows those launches scanned
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    int64_t scan_pruned_row_groups = 0, scan_pruned_rows = 0; // Parquet row groups skipped by statistics (parquet_exec.rs:143-196)
    std::vector<int64_t> partition_starts; // last ShuffleWriter batch: partition p = rows [starts[p], starts[p+1])
    void check_device_errors();
    void collect_timing();
};

struct ExecNode {
    std::vector<DType> schema;
    virtual ~ExecNode() {}
    virtual bool next(Batch& out) = 0; // false = end of stream
    // predicates (over this node's output columns) that the consumer applies to every row anyway: a source may use them to skip
    // data that cannot pass (Parquet row groups whose statistics rule them out)
    virtual void push_filters(const std::vector<ExprP>&) {}
    // rows this node will still produce, if it knows (-1: unknown): lets a hash aggregate size its table once instead of growing it
    virtual int64_t rows_hint() const { return -1; }
};
using ExecNodeP = std::shared_ptr<ExecNode>;

struct DeviceTable { // caller-owned device-resident columns bound as a plan input (bench "value" path)
    int64_t n_rows = 0;
    std::vector<Column> cols;
    bool needs_packing = false; // some columns were given byte-per-row validity / booleans: packed to Arrow bitmaps at first use
};

// Build the executor tree for a decoded plan.  `inputs` are consumed in Scan order.
struct PlanInputs {
    std::vector<ArrowArrayStream*> streams;
    std::vector<std::shared_ptr<DeviceTable>> tables; // parallel to streams; non-null entry overrides
};
ExecNodeP build_exec(const OperatorP& op, ExecContext* ctx, PlanInputs* inputs);

// native Parquet scan (scan_parquet.cpp)
ExecNodeP make_native_scan(const OperatorP& op, ExecContext* ctx);

// Export helpers (host-visible Arrow C Data)
void export_batch(Batch& b, ExecContext* ctx, ArrowArray* out_arrays, ArrowSchema* out_schemas, int n_cols, int64_t row0, int64_t n_rows);

// Debug / build-time: generate (and NVRTC-compile, no device needed) the kernels a plan would use,
// assuming inputs without nulls and dictionary-encoded string keys.
void register_memory_file(const std::string& name, const uint8_t* p, size_t n); // p == nullptr unregisters
std::string describe_parquet(const std::string& path);

std::vector<GeneratedKernel> plan_kernels_for_build(const OperatorP& op, const std::vector<int>& assume_bits = {});

} // namespace cb200
