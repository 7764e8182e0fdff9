// jit.h -- NVRTC compilation of generated pipeline kernels for sm_100a + module cache.
#pragma once
#include "codegen.h"

#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include <memory>
#include <string>
#include <vector>

namespace cb200 {

struct CompiledModule {
    std::vector<char> cubin;
    cudaLibrary_t lib = nullptr;
    std::map<std::string, cudaKernel_t> kernels;
    std::mutex mu; // modules are shared by every plan (and task thread) that uses the same pipeline
    bool loaded = false;
    ~CompiledModule();
    cudaKernel_t kernel(const std::string& name);
};

struct JitError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Compile (or fetch from the in-memory / on-disk cache) the module for `g`.  With load=false the
// cubin is produced but not loaded on a device (works on a CPU-only box: NVRTC needs no GPU).
std::shared_ptr<CompiledModule> jit_get(const GeneratedKernel& g, bool load);
std::string jit_cache_dir();

} // namespace cb200
