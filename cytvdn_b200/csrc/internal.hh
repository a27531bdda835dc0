// internal.hh -- helpers shared by the translation units of libcytvdn_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

namespace cytvdn_internal {
// records the calling thread's error message (cytvdn_last_error) and returns `code`
int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void count_launch();
// create the reduction workspace of (current device, stream) now (its first use allocates, which may synchronise the device)
int warm_workspace(cudaStream_t stream);
// page-locked host memory with the huge-page allocator of cytvdn_host_alloc
int pinned_alloc(void **out, size_t bytes);
int pinned_free(void *p);
}  // namespace cytvdn_internal

#define CYTVDN_CUDA_TRY(expr)                                                                               \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return cytvdn_internal::fail(_e == cudaErrorMemoryAllocation ? CYTVDN_E_NOMEM : CYTVDN_E_CUDA,  \
                                         "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
