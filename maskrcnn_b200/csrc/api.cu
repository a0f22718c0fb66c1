// api.cu — error plumbing shared by the extern "C" entry points of libmrcnn_b200.so.
#include <stdint.h>
#include <string.h>

#include "api_util.h"
#include "common.cuh"

namespace mrcnn {

int* device_error_word() {
    static int* words[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (words[dev] == nullptr) {
        int* p = nullptr;
        if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) return nullptr;
        cudaMemset(p, 0, sizeof(int));
        words[dev] = p;
    }
    return words[dev];
}

static thread_local char t_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_last_error, sizeof(t_last_error), fmt, ap);
    va_end(ap);
}

// cudaPointerGetAttributes costs a driver round trip per pointer; a small RoIAlign launch (~20 us of kernel) names six of
// them.  Pointers that were classified as device memory are remembered in a per-thread direct-mapped table, so the steady
// state of a loop that reuses its buffers (or a caching allocator that hands the same blocks back) validates for free.
// Only POSITIVE verdicts are cached: device virtual addresses stay device addresses while they are mapped, and a stale hit
// after the block was freed is no worse than the use-after-free it already is.  Host pointers are re-checked every time.
bool is_device_ptr(const void* p) {
    if (p == nullptr) return false;
    constexpr int kSlots = 256;
    static thread_local const void* seen[kSlots] = {nullptr};
    const uintptr_t v = reinterpret_cast<uintptr_t>(p);
    const int slot = (int)(((v >> 4) ^ (v >> 12) ^ (v >> 21)) & (kSlots - 1));
    if (seen[slot] == p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    const bool dev = a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
    if (dev) seen[slot] = p;
    return dev;
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

}  // namespace mrcnn

extern "C" {

int mrcnn_abi_version(void) { return MRCNN_ABI_VERSION; }

const char* mrcnn_last_error(void) { return mrcnn::t_last_error; }

int mrcnn_poll_device_errors(mrcnn_stream_t stream) {
    int flag = 0;
    int* word = mrcnn::device_error_word();
    if (word == nullptr) return mrcnn::fail(MRCNN_E_CUDA, "cannot allocate the device error word");
    MRCNN_CUDA(cudaMemcpyAsync(&flag, word, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MRCNN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (flag) {
        MRCNN_CUDA(cudaMemsetAsync(word, 0, sizeof(int), (cudaStream_t)stream));
        MRCNN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        if (flag & 1) return mrcnn::fail(MRCNN_E_BOX_INDEX, "box_index out of range [0, batch) in a crop/RoIAlign call");
        if (flag & 2) return mrcnn::fail(MRCNN_E_CLASS_ID, "class id out of range [0, num_classes) in mrcnn_full_masks");
    }
    return MRCNN_OK;
}

}  // extern "C"
