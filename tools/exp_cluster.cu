// exp_cluster.cu - what a cluster-wide synchronisation costs on B200, for 8-CTA clusters of 1024-thread CTAs (the shape of
// proposal_select_kernel / proposal_lazy_nms_kernel):
//   (a) cooperative-groups cluster.sync() (barrier.cluster.arrive + wait by every thread),
//   (b) __syncthreads + ONE thread per CTA arriving on every peer's mbarrier + everybody waiting on the local mbarrier,
//   (c) a dependent load through distributed shared memory (pointer chase in a peer's shared memory),
//   (d) __syncthreads alone, for scale.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/exp_cluster tools/exp_cluster.cu ; run on the GPU box.
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(long long* out, int iters, int nthreads_active) {
    __shared__ uint64_t bar;
    __shared__ int chase[1024];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), tid = threadIdx.x;
    const int csize = (int)cluster.num_blocks();
    chase[tid] = (tid * 37 + 11) & 1023;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(csize));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();
    long long t0 = clock64();
    int acc = tid;
    if (MODE == 0) {
        for (int i = 0; i < iters; ++i) cluster.sync();
    } else if (MODE == 1) {
        for (int i = 0; i < iters; ++i) {
            __syncthreads();
            if (tid < csize) {
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32(&bar), (uint32_t)tid)) : "memory");
            }
            uint32_t done = 0;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(done) : "r"(smem_u32(&bar)), "r"((uint32_t)(i & 1)) : "memory");
            } while (!done);
        }
    } else if (MODE == 2) {
        const int* peer = cluster.map_shared_rank(chase, (unsigned)((rank + 1) % csize));
        for (int i = 0; i < iters; ++i) acc = peer[acc & 1023];
    } else if (MODE == 3) {
        for (int i = 0; i < iters; ++i) __syncthreads();
    } else if (MODE == 4) {
        for (int i = 0; i < iters; ++i) acc = chase[acc & 1023];
    }
    long long t1 = clock64();
    cluster.sync();
    if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == -12345) out[1] = acc;
}

template <int MODE>
static void run(const char* name, int csize, int iters) {
    long long* d;
    cudaMalloc(&d, 16);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize * 8);
    cfg.blockDim = dim3(1024);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, k<MODE>, d, iters, 1024);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-58s cluster %d: %8.1f cycles per iteration (%s)\n", name, csize, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    for (int cs : {2, 8}) {
        run<0>("(a) cluster.sync(), 1024 threads per CTA", cs, 200);
        run<1>("(b) __syncthreads + 1 remote mbarrier arrive per peer + wait", cs, 200);
        run<2>("(c) dependent load through distributed shared memory", cs, 2000);
    }
    run<3>("(d) __syncthreads, 1024 threads", 1, 2000);
    run<4>("(e) dependent load from the CTA's own shared memory", 1, 2000);
    return 0;
}
