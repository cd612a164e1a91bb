// knockoffgwas_b200/csrc/kgw_common.h -- helpers shared by the translation units of libkgw.so
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

// cudaMalloc on the current device that knows about the buffers destroyed plans parked in the library's pool
// (kgw_api.cu): on out-of-memory the pool of this device is released and the allocation retried once.
cudaError_t kgw_dev_malloc_(void **out, size_t bytes);
template <typename T>
static inline cudaError_t kgw_malloc(T **out, size_t bytes) { return kgw_dev_malloc_(reinterpret_cast<void **>(out), bytes); }
