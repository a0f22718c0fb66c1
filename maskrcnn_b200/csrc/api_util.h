// api_util.h — host-side helpers shared by the extern "C" entry points.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/mrcnn_b200.h"

namespace mrcnn {

void set_last_error(const char* fmt, ...);

// true iff p is device (or managed) memory.  NULL is not.
bool is_device_ptr(const void* p);

inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    set_last_error("%s", buf);
    return code;
}

#define MRCNN_REQUIRE(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) return ::mrcnn::fail(MRCNN_E_INVALID_ARG, __VA_ARGS__); \
    } while (0)

#define MRCNN_REQUIRE_DEV(ptr)                                                                         \
    do {                                                                                               \
        if (!::mrcnn::is_device_ptr(ptr))                                                              \
            return ::mrcnn::fail(MRCNN_E_NOT_DEVICE_PTR, "%s: '%s' is not a device pointer (no CPU path)", \
                                 __func__, #ptr);                                                      \
    } while (0)

#define MRCNN_CUDA(call)                                                                          \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return ::mrcnn::fail(MRCNN_E_CUDA, "%s: %s -> %s", __func__, #call, cudaGetErrorString(e__)); \
    } while (0)

#define MRCNN_LAUNCH_CHECK() MRCNN_CUDA(cudaGetLastError())

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int sm_count();

}  // namespace mrcnn
