// exp_bw.cu — what HBM bandwidth does a B200 give a streaming kernel as a function of its read : write mix?
// MEASURED_PEAKS.json holds the copy figure (1 : 1).  The RoIAlign forward writes 70 % of its bytes, the mask paste-back
// 100 %, the gather backward 43 %: this tool measures the ceiling for each mix with a plain grid-stride kernel that reads
// R and writes W 16-byte vectors per thread and iteration from / to 2 GB buffers.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_bw tools/exp_bw.cu ; run: ./exp_bw
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int W, int ST>
__global__ void __launch_bounds__(256) mix_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n_iter, uint4* sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_iter; i += stride) {
        uint4 v[R > 0 ? R : 1];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = __ldcs(src + (size_t)r * n_iter + i);
#pragma unroll
        for (int r = 0; r < R; ++r) { acc.x ^= v[r].x; acc.y += v[r].y; acc.z ^= v[r].z; acc.w += v[r].w; }
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const uint4 o = make_uint4(acc.x + w, acc.y, acc.z, (unsigned)i);
            if (ST == 0) dst[(size_t)w * n_iter + i] = o;
            else if (ST == 1) __stcs(dst + (size_t)w * n_iter + i, o);
            else __stwt(dst + (size_t)w * n_iter + i, o);
        }
    }
    if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) *sink = acc;
}

// the same mix, but a CTA streams whole contiguous 32 KB blocks (block-cyclic) with U vectors in flight per thread and stream
template <int R, int W, int U>
__global__ void __launch_bounds__(256) mix_block_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n_iter, uint4* sink) {
    constexpr int kBlock = 256 * U;  // vectors per CTA block and stream
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t n_blocks = n_iter / kBlock;
    for (size_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const size_t base = b * kBlock + threadIdx.x;
        uint4 v[(R > 0 ? R : 1) * U];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int u = 0; u < U; ++u) v[r * U + u] = __ldcs(src + (size_t)r * n_iter + base + u * 256);
#pragma unroll
        for (int i = 0; i < R * U; ++i) { acc.x ^= v[i].x; acc.y += v[i].y; acc.z ^= v[i].z; acc.w += v[i].w; }
#pragma unroll
        for (int w = 0; w < W; ++w)
#pragma unroll
            for (int u = 0; u < U; ++u) __stcs(dst + (size_t)w * n_iter + base + u * 256, make_uint4(acc.x + w, acc.y, acc.z, u));
    }
    if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) *sink = acc;
}

template <int R, int W, int U>
void run_block(const uint4* src, uint4* dst, size_t bytes_each, uint4* sink, int ctas_per_sm) {
    const size_t vec_total = bytes_each / 16;
    const int m = R > W ? R : W;
    const size_t n_iter = vec_total / m / (256 * U) * (256 * U);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    for (int i = 0; i < 2; ++i) mix_block_kernel<R, W, U><<<grid, 256>>>(src, dst, n_iter, sink);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) mix_block_kernel<R, W, U><<<grid, 256>>>(src, dst, n_iter, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)n_iter * 16 * (R + W);
    printf("block-cyclic R=%d W=%d U=%d ctas/sm=%d : %8.1f GB/s  (%.0f %% writes)\n", R, W, U, ctas_per_sm,
           bytes / (ms / reps * 1e-3) / 1e9, 100.0 * W / (R + W));
}

template <int R, int W, int ST>
void run(const uint4* src, uint4* dst, size_t bytes_each, uint4* sink, int ctas_per_sm) {
    const size_t vec_total = bytes_each / 16;
    const int m = R > W ? R : W;
    const size_t n_iter = vec_total / m;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    for (int i = 0; i < 2; ++i) mix_kernel<R, W, ST><<<grid, 256>>>(src, dst, n_iter, sink);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) mix_kernel<R, W, ST><<<grid, 256>>>(src, dst, n_iter, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)n_iter * 16 * (R + W);
    printf("R=%d W=%d st=%s ctas/sm=%d : %8.1f GB/s  (%.0f %% writes, %.1f us / %.2f GB)\n", R, W, ST == 0 ? "default" : (ST == 1 ? "cs" : "wt"),
           ctas_per_sm, bytes / (ms / reps * 1e-3) / 1e9, 100.0 * W / (R + W), ms / reps * 1e3, bytes / 1e9);
}

int main() {
    const size_t bytes_each = (size_t)2 << 30;
    uint4 *src, *dst, *sink;
    cudaMalloc(&src, bytes_each);
    cudaMalloc(&dst, bytes_each);
    cudaMalloc(&sink, 16);
    cudaMemset(src, 1, bytes_each);
    cudaMemset(dst, 0, bytes_each);
    for (int c : {8}) {
        run<1, 0, 0>(src, dst, bytes_each, sink, c);
        run<0, 1, 0>(src, dst, bytes_each, sink, c);
        run<0, 1, 1>(src, dst, bytes_each, sink, c);
        run<0, 1, 2>(src, dst, bytes_each, sink, c);
        run<1, 1, 0>(src, dst, bytes_each, sink, c);
        run<1, 1, 1>(src, dst, bytes_each, sink, c);
        run<2, 1, 1>(src, dst, bytes_each, sink, c);
        run<4, 3, 1>(src, dst, bytes_each, sink, c);
        run<1, 2, 1>(src, dst, bytes_each, sink, c);
        run<3, 7, 1>(src, dst, bytes_each, sink, c);
        run<3, 2, 1>(src, dst, bytes_each, sink, c);
    }
    for (int c : {2, 4, 8}) {
        run_block<0, 1, 4>(src, dst, bytes_each, sink, c);
        run_block<0, 1, 8>(src, dst, bytes_each, sink, c);
        run_block<1, 0, 4>(src, dst, bytes_each, sink, c);
        run_block<1, 1, 4>(src, dst, bytes_each, sink, c);
        run_block<3, 7, 2>(src, dst, bytes_each, sink, c);
        run_block<3, 7, 4>(src, dst, bytes_each, sink, c);
        run_block<4, 3, 2>(src, dst, bytes_each, sink, c);
        run_block<4, 3, 4>(src, dst, bytes_each, sink, c);
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaMemsetAsync(dst, 0, bytes_each);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) cudaMemsetAsync(dst, 0, bytes_each);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemsetAsync 2 GB: %.1f GB/s\n", bytes_each / (ms / 10 * 1e-3) / 1e9);
    cudaMemcpyAsync(dst, src, bytes_each, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) cudaMemcpyAsync(dst, src, bytes_each, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpyAsync D2D 2 GB: %.1f GB/s (read + write)\n", 2.0 * bytes_each / (ms / 10 * 1e-3) / 1e9);
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
