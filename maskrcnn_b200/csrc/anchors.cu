// anchors.cu — RPN anchor matching for sm_100a.
//
// Replaces data.rpn_samples (data.py:449-591), numpy code that runs in the DataLoader for every training sample:
// boxes_overlaps materialises a [261,888 * G, 4] tiled box matrix on the CPU, then argmax / where / random.choice
// and a Python loop over the positive anchors.  Here:
//   rpn_match_kernel      one thread per anchor, gt boxes in shared memory: fp32 IoU exactly as data.py:151-189 (the
//                         numpy inputs go through .float()), running max / first argmax per anchor in registers, and
//                         the per-gt column argmax (first anchor with the largest IoU, np.argmax) by a warp
//                         max-reduction + one 64-bit atomicMax of (IoU key, ~anchor index) per warp and gt box;
//   rpn_apply_gt_kernel   every gt box claims its best anchor (data.py:531-532);
//   count / compact       ascending index lists of the anchors equal to +1 / -1 (np.where order) without a host pass;
//   scatter_fill_kernel   resets the subsampled-away anchors to neutral (data.py:543-553);
//   rpn_deltas_kernel     float64 box deltas / std for the positive anchors (data.py:557-589).
// The two np.random.choice draws stay with the caller (see ops.rpn_samples), so a seeded run reproduces the reference.
#include <limits.h>

#include "api_util.h"
#include "nms_core.cuh"

namespace mrcnn {

constexpr int kAncThreads = 256;
constexpr int kCompactBlock = 1024;

__device__ __forceinline__ float tmaxf2(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : (a > b ? a : b); }
__device__ __forceinline__ float tminf2(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : (a < b ? a : b); }

// data.py:151-189 for one pair (no +1), one rounding per operation
__device__ __forceinline__ float iou_plain(const float4 a, const float4 b) {
    const float y1 = tmaxf2(a.x, b.x), x1 = tmaxf2(a.y, b.y);
    const float y2 = tminf2(a.z, b.z), x2 = tminf2(a.w, b.w);
    const float inter = __fmul_rn(tmaxf2(__fsub_rn(x2, x1), 0.0f), tmaxf2(__fsub_rn(y2, y1), 0.0f));
    const float a1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float a2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(a1, a2), inter));
}

struct RpnMatchParams {
    const double* anchors;    // [A,4] px
    const int32_t* gt_boxes;  // [G,4] px
    const int32_t* gt_class;  // [G]
    int A, G;
    int32_t* match;           // [A]
    int32_t* argmax;          // [A]
    unsigned long long* gt_best;  // [G] (IoU key << 32) | ~anchor index, zero-initialised
};

__global__ void __launch_bounds__(kAncThreads) rpn_match_kernel(const RpnMatchParams p) {
    extern __shared__ __align__(16) unsigned char anc_smem[];
    float4* s_gt = reinterpret_cast<float4*>(anc_smem);                                  // [G]
    int* s_cls = reinterpret_cast<int*>(anc_smem + (size_t)p.G * 16);                     // [G]
    unsigned long long* s_best = reinterpret_cast<unsigned long long*>(anc_smem + (((size_t)p.G * 20 + 7) & ~(size_t)7));  // [G]
    __shared__ int s_any_crowd;
    const int tid = threadIdx.x, lane = tid & 31;
    const int G = p.G;
    if (tid == 0) s_any_crowd = 0;
    __syncthreads();
    for (int g = tid; g < G; g += blockDim.x) {
        const int32_t* b = p.gt_boxes + 4 * g;
        s_gt[g] = make_float4((float)__ldg(b), (float)__ldg(b + 1), (float)__ldg(b + 2), (float)__ldg(b + 3));
        const int c = __ldg(p.gt_class + g);
        s_cls[g] = c;
        s_best[g] = 0ull;
        if (c < 0) s_any_crowd = 1;
    }
    __syncthreads();
    const bool any_crowd = s_any_crowd != 0;
    const int i = blockIdx.x * blockDim.x + tid;
    const bool live = i < p.A;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        const double* q = p.anchors + 4 * (size_t)i;
        a = make_float4((float)q[0], (float)q[1], (float)q[2], (float)q[3]);  // torch.from_numpy(boxes1).float(), data.py:156
    }
    float best = -INFINITY, crowd_best = -INFINITY;
    int besti = -1;
    bool nan_seen = false, crowd_nan = false, first = true;
    for (int g = 0; g < G; ++g) {
        const int c = s_cls[g];
        if (any_crowd && c == 0) continue;  // data.py:500 non_crowd_ix = class > 0
        const float v = iou_plain(a, s_gt[g]);
        if (any_crowd && c < 0) {  // data.py:503-506
            if (v != v) crowd_nan = true;
            else if (v > crowd_best) crowd_best = v;
            continue;
        }
        if (first) {
            besti = g;
            first = false;
        }
        if (v != v) {
            if (!nan_seen) besti = g;
            nan_seen = true;
        } else if (!nan_seen && v > best) {
            best = v;
            besti = g;
        }
        // column argmax (data.py:531): first anchor with the largest IoU; NaN is maximal for np.argmax
        const uint32_t key = live ? ((v != v) ? 0xffffffffu : float_to_key(v)) : 0u;
        const uint32_t mx = __reduce_max_sync(0xffffffffu, key);
        const unsigned who = __ballot_sync(0xffffffffu, key == mx && live);
        if (who && lane == __ffs(who) - 1)
            atomicMax(&s_best[g], ((unsigned long long)mx << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i));
    }
    if (live) {
        const float m = nan_seen ? __int_as_float(0x7fc00000) : best;
        const bool no_crowd = any_crowd ? (!crowd_nan && crowd_best < 0.001f) : true;  // np.amax propagates NaN
        int lab = 0;
        if (m < 0.3f && no_crowd) lab = -1;  // data.py:527
        if (m >= 0.7f) lab = 1;              // data.py:535
        p.match[i] = lab;
        p.argmax[i] = besti;
    }
    __syncthreads();
    for (int g = tid; g < G; g += blockDim.x)
        if (s_best[g]) atomicMax(p.gt_best + g, s_best[g]);
}

__global__ void rpn_apply_gt_kernel(const unsigned long long* __restrict__ gt_best, const int32_t* __restrict__ gt_class, int G,
                                    int32_t* __restrict__ match) {
    __shared__ int s_any_crowd;
    if (threadIdx.x == 0) s_any_crowd = 0;
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x)
        if (__ldg(gt_class + g) < 0) s_any_crowd = 1;
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const int c = __ldg(gt_class + g);
        if (s_any_crowd && c <= 0) continue;
        const unsigned long long b = gt_best[g];
        if (b) match[0xffffffffu - (uint32_t)(b & 0xffffffffull)] = 1;  // data.py:532
    }
}

// ---- ordered compaction of the indices i with values[i] == target (np.where order) ----
__global__ void __launch_bounds__(kCompactBlock) count_equal_kernel(const int32_t* __restrict__ values, int n, int target,
                                                                    int32_t* __restrict__ block_counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = __syncthreads_count(i < n && __ldg(values + i) == target);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(kCompactBlock) compact_equal_kernel(const int32_t* __restrict__ values, int n, int target,
                                                                      const int32_t* __restrict__ block_counts,
                                                                      int32_t* __restrict__ ids_out, int32_t* __restrict__ total_out) {
    __shared__ int s_warp[kCompactBlock / 32];
    __shared__ int s_red[kCompactBlock / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // offset of this block = sum of the counts of the blocks before it
    int part = 0;
    for (int b = tid; b < (int)blockIdx.x; b += blockDim.x) part += __ldg(block_counts + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < kCompactBlock / 32; ++w) base += s_red[w];
    const int i = blockIdx.x * blockDim.x + tid;
    const bool flag = i < n && __ldg(values + i) == target;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < kCompactBlock / 32; ++w) {
        const int v = s_warp[w];
        if (w < warp) before += v;
        total += v;
    }
    if (flag) ids_out[base + before + __popc(m & ((1u << lane) - 1u))] = i;
    if (blockIdx.x == gridDim.x - 1 && tid == 0) *total_out = base + total;
}

__global__ void scatter_fill_kernel(int32_t* __restrict__ values, const int32_t* __restrict__ ids, const int32_t* __restrict__ perm,
                                    int count, int32_t fill) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) values[__ldg(ids + __ldg(perm + t))] = fill;
}

__global__ void rpn_deltas_kernel(const double* __restrict__ anchors, const int32_t* __restrict__ gt_boxes,
                                  const int32_t* __restrict__ argmax, const int32_t* __restrict__ ids, const int32_t* __restrict__ count,
                                  int T, double s0, double s1, double s2, double s3, double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double4 o = make_double4(0.0, 0.0, 0.0, 0.0);
    if (t < __ldg(count)) {  // data.py:557-589, float64 like the numpy code (gt boxes: int32 arithmetic first)
        const int i = __ldg(ids + t);
        const double* a = anchors + 4 * (size_t)i;
        const int32_t* gt = gt_boxes + 4 * (size_t)__ldg(argmax + i);
        const double gt_h = (double)(gt[2] - gt[0]), gt_w = (double)(gt[3] - gt[1]);
        const double gcy = __dadd_rn((double)gt[0], __dmul_rn(0.5, gt_h)), gcx = __dadd_rn((double)gt[1], __dmul_rn(0.5, gt_w));
        const double a_h = __dsub_rn(a[2], a[0]), a_w = __dsub_rn(a[3], a[1]);
        const double acy = __dadd_rn(a[0], __dmul_rn(0.5, a_h)), acx = __dadd_rn(a[1], __dmul_rn(0.5, a_w));
        o.x = __ddiv_rn(__ddiv_rn(__dsub_rn(gcy, acy), a_h), s0);
        o.y = __ddiv_rn(__ddiv_rn(__dsub_rn(gcx, acx), a_w), s1);
        o.z = __ddiv_rn(log(__ddiv_rn(gt_h, a_h)), s2);
        o.w = __ddiv_rn(log(__ddiv_rn(gt_w, a_w)), s3);
    }
    reinterpret_cast<double4*>(out)[t] = o;
}

}  // namespace mrcnn

using namespace mrcnn;

extern "C" {

size_t mrcnn_rpn_match_workspace_bytes(int G) { return align_up((size_t)(G > 0 ? G : 1) * 8, 256); }

int mrcnn_rpn_match(const double* anchors, int A, const int32_t* gt_boxes, const int32_t* gt_class_ids, int G, int32_t* match,
                    int32_t* argmax, void* workspace, size_t workspace_bytes, mrcnn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRCNN_REQUIRE(A > 0 && G > 0, "mrcnn_rpn_match: needs at least one anchor and one gt box (the reference fails on G == 0)");
    MRCNN_REQUIRE((size_t)G * 28 + 64 <= 200 * 1024, "mrcnn_rpn_match: too many gt boxes");
    MRCNN_REQUIRE_DEV(anchors);
    MRCNN_REQUIRE_DEV(gt_boxes);
    MRCNN_REQUIRE_DEV(gt_class_ids);
    MRCNN_REQUIRE_DEV(match);
    MRCNN_REQUIRE_DEV(argmax);
    MRCNN_REQUIRE_DEV(workspace);
    if (workspace_bytes < mrcnn_rpn_match_workspace_bytes(G) || (reinterpret_cast<uintptr_t>(workspace) & 7u))
        return fail(MRCNN_E_WORKSPACE, "mrcnn_rpn_match: workspace too small or misaligned");
    RpnMatchParams p = {anchors, gt_boxes, gt_class_ids, A, G, match, argmax, reinterpret_cast<unsigned long long*>(workspace)};
    MRCNN_CUDA(cudaMemsetAsync(workspace, 0, (size_t)G * 8, stream));
    const size_t smem = (((size_t)G * 20 + 7) & ~(size_t)7) + (size_t)G * 8;
    if (smem > 48 * 1024)
        MRCNN_CUDA(cudaFuncSetAttribute(rpn_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rpn_match_kernel<<<(A + kAncThreads - 1) / kAncThreads, kAncThreads, smem, stream>>>(p);
    MRCNN_LAUNCH_CHECK();
    rpn_apply_gt_kernel<<<1, 256, 0, stream>>>(p.gt_best, gt_class_ids, G, match);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

size_t mrcnn_compact_equal_workspace_bytes(int n) {
    return align_up((size_t)((n > 0 ? n : 1) + kCompactBlock - 1) / kCompactBlock * 4, 256);
}

int mrcnn_compact_equal(const int32_t* values, int n, int32_t target, int32_t* ids_out, int32_t* count_out, void* workspace,
                        size_t workspace_bytes, mrcnn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRCNN_REQUIRE(n > 0, "mrcnn_compact_equal: n must be positive");
    MRCNN_REQUIRE_DEV(values);
    MRCNN_REQUIRE_DEV(ids_out);
    MRCNN_REQUIRE_DEV(count_out);
    MRCNN_REQUIRE_DEV(workspace);
    if (workspace_bytes < mrcnn_compact_equal_workspace_bytes(n) || (reinterpret_cast<uintptr_t>(workspace) & 3u))
        return fail(MRCNN_E_WORKSPACE, "mrcnn_compact_equal: workspace too small or misaligned");
    const int blocks = (n + kCompactBlock - 1) / kCompactBlock;
    int32_t* counts = reinterpret_cast<int32_t*>(workspace);
    count_equal_kernel<<<blocks, kCompactBlock, 0, stream>>>(values, n, target, counts);
    MRCNN_LAUNCH_CHECK();
    compact_equal_kernel<<<blocks, kCompactBlock, 0, stream>>>(values, n, target, counts, ids_out, count_out);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

int mrcnn_scatter_fill(int32_t* values, const int32_t* ids, const int32_t* perm, int count, int32_t fill, mrcnn_stream_t stream) {
    MRCNN_REQUIRE(count >= 0, "mrcnn_scatter_fill: bad count");
    if (count == 0) return MRCNN_OK;
    MRCNN_REQUIRE_DEV(values);
    MRCNN_REQUIRE_DEV(ids);
    MRCNN_REQUIRE_DEV(perm);
    scatter_fill_kernel<<<(count + 255) / 256, 256, 0, (cudaStream_t)stream>>>(values, ids, perm, count, fill);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

int mrcnn_rpn_deltas(const double* anchors, const int32_t* gt_boxes, const int32_t* argmax, const int32_t* ids,
                     const int32_t* count, int T, const double* std4_host, double* rpn_bbox_out, mrcnn_stream_t stream) {
    MRCNN_REQUIRE(T > 0 && std4_host, "mrcnn_rpn_deltas: bad sizes");
    MRCNN_REQUIRE_DEV(anchors);
    MRCNN_REQUIRE_DEV(gt_boxes);
    MRCNN_REQUIRE_DEV(argmax);
    MRCNN_REQUIRE_DEV(ids);
    MRCNN_REQUIRE_DEV(count);
    MRCNN_REQUIRE_DEV(rpn_bbox_out);
    MRCNN_REQUIRE((reinterpret_cast<uintptr_t>(rpn_bbox_out) & 31u) == 0, "mrcnn_rpn_deltas: rpn_bbox_out must be 32-byte aligned");
    rpn_deltas_kernel<<<(T + 127) / 128, 128, 0, (cudaStream_t)stream>>>(anchors, gt_boxes, argmax, ids, count, T, std4_host[0],
                                                                          std4_host[1], std4_host[2], std4_host[3], rpn_bbox_out);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

}  // extern "C"
