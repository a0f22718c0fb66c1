// targets.cu — the detection-target layer for sm_100a, batched over images.
//
// Replaces mrn_samples (model.py:396-576) with data.boxes_overlaps (data.py:151-189) and data.boxes_deltas
// (data.py:103-121): the reference tiles an [N*G, 4] box matrix, takes max / nonzero / randperm on the host side of
// several syncs, materialises gt_masks[assignment] (P x 4 MB) and only then crops 28x28 targets.  Here:
//   1. target_classify_kernel  one CTA per image: IoU of every proposal against the gt boxes held in shared
//                              memory (crowd rows apart), running max / first argmax in registers, then an ordered
//                              block compaction into the positive and negative index lists (torch.nonzero order).
//   2. target_select_kernel    (sync-free mode) one CTA per image: the reference's two torch.randperm draws are
//                              replaced by caller-supplied random keys; perm = stable argsort(keys[:count]) by a
//                              shared-memory bitonic sort; how many to keep comes from a host-built table so that
//                              int(r * p - p) is evaluated in Python doubles exactly like the reference.
//   3. target_emit_kernel      one CTA per output row: gather the RoI, class id and gt box, box deltas / std with a
//                              correctly rounded log, and the 28x28 mask target cropped straight out of the assigned
//                              gt mask (no gt_masks[assignment] copy), rounded half-to-even; zero rows for
//                              negatives and padding.
// Parity: selections, rois, class ids, dy/dx and masks bit-exact; dh/dw are the correctly rounded fp32 logs.
#include <limits.h>

#include "api_util.h"
#include "nms_core.cuh"

namespace mrcnn {

constexpr int kTgtThreads = 1024;
constexpr int kTgtMaxSort = 8192;  // proposals per image the select kernel can rank

// torch.max / torch.min of two tensors: NaN propagates
__device__ __forceinline__ float tmaxf(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : (a > b ? a : b); }
__device__ __forceinline__ float tminf(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : (a < b ? a : b); }

// data.py:151-189 for one pair of normalised boxes (no +1), one rounding per operation
__device__ __forceinline__ float box_iou_norm(const float4 a, const float4 b) {
    const float y1 = tmaxf(a.x, b.x), x1 = tmaxf(a.y, b.y);
    const float y2 = tminf(a.z, b.z), x2 = tminf(a.w, b.w);
    const float inter = __fmul_rn(tmaxf(__fsub_rn(x2, x1), 0.0f), tmaxf(__fsub_rn(y2, y1), 0.0f));
    const float a1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float a2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(a1, a2), inter));
}

struct ClassifyParams {
    const float* rois;          // [B,N,4] normalised
    const float* gt_boxes;      // [B,G,4] normalised
    const int32_t* gt_class;    // [B,G]
    int B, N, G;
    int32_t* pos_idx;  // [B,N]
    int32_t* neg_idx;  // [B,N]
    int32_t* assign;   // [B,N] gt row of the max IoU (-1: none)
    float* iou_max;    // [B,N] or null
    int32_t* counts;   // [B,2]
};

// Ordered compaction of `flag` over the CTA: returns this thread's position (valid if flag) and adds the total to base.
__device__ __forceinline__ int block_compact(bool flag, int* s_warp, int& base) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        const int v = s_warp[w];
        if (w < warp) before += v;
        total += v;
    }
    const int pos = base + before + __popc(m & ((1u << lane) - 1u));
    base += total;
    __syncthreads();
    return pos;
}

__global__ void __launch_bounds__(kTgtThreads) target_classify_kernel(const ClassifyParams p) {
    extern __shared__ __align__(16) unsigned char tgt_smem[];
    float4* s_gt = reinterpret_cast<float4*>(tgt_smem);              // [G]
    int* s_cls = reinterpret_cast<int*>(tgt_smem + (size_t)p.G * 16);  // [G]
    __shared__ int s_warp[kTgtThreads / 32];
    __shared__ int s_any_crowd;
    const int img = blockIdx.x, tid = threadIdx.x;
    const int N = p.N, G = p.G;
    if (tid == 0) s_any_crowd = 0;
    __syncthreads();
    for (int g = tid; g < G; g += blockDim.x) {
        const float* b = p.gt_boxes + ((size_t)img * G + g) * 4;
        s_gt[g] = make_float4(__ldg(b), __ldg(b + 1), __ldg(b + 2), __ldg(b + 3));
        const int c = __ldg(p.gt_class + (size_t)img * G + g);
        s_cls[g] = c;
        if (c < 0) s_any_crowd = 1;  // benign race: every writer stores 1
    }
    __syncthreads();
    const bool any_crowd = s_any_crowd != 0;
    int npos = 0, nneg = 0;
    for (int i0 = 0; i0 < N; i0 += blockDim.x) {
        const int i = i0 + tid;
        bool pos = false, neg = false;
        if (i < N) {
            const float* r = p.rois + ((size_t)img * N + i) * 4;
            const float4 box = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
            float best = -INFINITY, crowd_best = -INFINITY;
            int besti = -1;
            bool nan_seen = false, crowd_nan = false, have = false, have_crowd = false;
            for (int g = 0; g < G; ++g) {
                const int c = s_cls[g];
                if (any_crowd && c == 0) continue;  // model.py:439 non_crowd_ix = class > 0
                const float v = box_iou_norm(box, s_gt[g]);
                if (any_crowd && c < 0) {  // model.py:444-447 crowd_iou_max
                    have_crowd = true;
                    if (v != v) crowd_nan = true;
                    else if (v > crowd_best) crowd_best = v;
                    continue;
                }
                have = true;
                if (v != v) {  // torch.max propagates NaN
                    if (!nan_seen) besti = g;
                    nan_seen = true;
                } else if (!nan_seen && v > best) {
                    best = v;
                    besti = g;
                }
            }
            const float m = nan_seen ? __int_as_float(0x7fc00000) : best;
            const bool no_crowd = any_crowd ? (have_crowd && !crowd_nan && crowd_best < 0.001f) : true;  // :447-449
            pos = have && m >= 0.5f;             // :459
            neg = have && m < 0.5f && no_crowd;  // :513-514
            p.assign[(size_t)img * N + i] = besti;
            if (p.iou_max) p.iou_max[(size_t)img * N + i] = m;
        }
        const int pp = block_compact(pos, s_warp, npos);
        if (pos) p.pos_idx[(size_t)img * N + pp] = i;
        const int pn = block_compact(neg, s_warp, nneg);
        if (neg) p.neg_idx[(size_t)img * N + pn] = i;
    }
    if (tid == 0) {
        p.counts[2 * img] = npos;
        p.counts[2 * img + 1] = nneg;
    }
}

struct SelectParams {
    const int32_t* counts;     // [B,2]
    const float* keys_pos;     // [B,N] random keys, one per positive-list slot
    const float* keys_neg;     // [B,N]
    const int32_t* neg_table;  // [pos_cap + 1]: negatives to keep for a given number of kept positives
    int B, N, pos_cap, P2;     // P2: power of two >= N
    int32_t* perm_pos;  // [B,N]
    int32_t* perm_neg;  // [B,N]
    int32_t* take;      // [B,2]
};

// perm[0..n) = stable argsort(keys[0..n)) ascending
__device__ __forceinline__ void block_argsort_asc(const float* keys, int n, int P2, uint64_t* s, int32_t* perm) {
    const int tid = threadIdx.x;
    if (P2 == kTgtThreads) {  // up to 1024 proposals: one key per thread, register / shuffle network (s holds 2 x 1024 keys)
        uint64_t kv = 0ull;
        if (tid < n) kv = ~(((uint64_t)float_to_key(__ldg(keys + tid)) << 32) | (uint32_t)tid);
        kv = block_bitonic_desc_1024_reg(kv, s);
        if (tid < n) perm[tid] = (int32_t)((~kv) & 0xffffffffu);
        __syncthreads();  // the exchange buffer is reused by the next call
        return;
    }
    for (int i = tid; i < P2; i += blockDim.x) {
        // descending sort of ~(key, index) == ascending (key, index); padding sorts last
        uint64_t kv = 0ull;
        if (i < n) kv = ~(((uint64_t)float_to_key(__ldg(keys + i)) << 32) | (uint32_t)i);
        s[i] = kv;
    }
    __syncthreads();
    block_bitonic_desc(s, P2, 0u, 2u, 1u, (unsigned)P2);
    for (int i = tid; i < n; i += blockDim.x) perm[i] = (int32_t)((~s[i]) & 0xffffffffu);
    __syncthreads();
}

__global__ void __launch_bounds__(kTgtThreads) target_select_kernel(const SelectParams p) {
    extern __shared__ __align__(16) unsigned char tgt_smem[];
    uint64_t* s = reinterpret_cast<uint64_t*>(tgt_smem);
    const int img = blockIdx.x;
    const int P = __ldg(p.counts + 2 * img), Q = __ldg(p.counts + 2 * img + 1);
    block_argsort_asc(p.keys_pos + (size_t)img * p.N, P, p.P2, s, p.perm_pos + (size_t)img * p.N);
    block_argsort_asc(p.keys_neg + (size_t)img * p.N, Q, p.P2, s, p.perm_neg + (size_t)img * p.N);
    if (threadIdx.x == 0) {
        const int tp = min(P, p.pos_cap);                                // model.py:466-471
        const int tn = (tp > 0) ? min(Q, __ldg(p.neg_table + tp)) : 0;    // model.py:517-522
        p.take[2 * img] = tp;
        p.take[2 * img + 1] = tn;
    }
}

struct EmitParams {
    const float* rois;        // [B,N,4]
    const float* gt_boxes;    // [B,G,4]
    const int32_t* gt_class;  // [B,G]
    const float* gt_masks;    // [B,G,H,W]
    int B, N, G, H, W;
    const int32_t* pos_idx;   // [B,N]
    const int32_t* neg_idx;   // [B,N]
    const int32_t* perm_pos;  // [B,N] or null (identity)
    const int32_t* perm_neg;  // [B,N] or null
    const int32_t* take;      // [B,2] kept positives / negatives
    const int32_t* assign;    // [B,N]
    float std0, std1, std2, std3;
    int mh, mw, T;            // T: output rows per image (zero padded)
    float* rois_out;          // [B,T,4]
    int32_t* class_out;       // [B,T]
    float* deltas_out;        // [B,T,4]
    float* masks_out;         // [B,T,mh,mw]
};

__device__ __forceinline__ float log_cr(float x) { return (float)log((double)x); }

__global__ void __launch_bounds__(256) target_emit_kernel(const EmitParams p) {
    const int t = blockIdx.x, img = blockIdx.y, tid = threadIdx.x;
    const int tp = __ldg(p.take + 2 * img), tn = __ldg(p.take + 2 * img + 1);
    const size_t row = (size_t)img * p.T + t;
    float* mo = p.masks_out + row * p.mh * p.mw;
    const int npx = p.mh * p.mw;
    if (t >= tp) {  // negatives and padding: zero deltas / masks (model.py:525-541)
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < tp + tn) {
            const int k = t - tp;
            const int slot = p.perm_neg ? __ldg(p.perm_neg + (size_t)img * p.N + k) : k;
            const int i = __ldg(p.neg_idx + (size_t)img * p.N + slot);
            const float* r = p.rois + ((size_t)img * p.N + i) * 4;
            box = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
        }
        if (tid == 0) {
            reinterpret_cast<float4*>(p.rois_out)[row] = box;
            p.class_out[row] = 0;
            reinterpret_cast<float4*>(p.deltas_out)[row] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int k = tid; k < npx; k += blockDim.x) mo[k] = 0.f;
        return;
    }
    const int slot = p.perm_pos ? __ldg(p.perm_pos + (size_t)img * p.N + t) : t;
    const int i = __ldg(p.pos_idx + (size_t)img * p.N + slot);
    const int g = __ldg(p.assign + (size_t)img * p.N + i);
    const float* r = p.rois + ((size_t)img * p.N + i) * 4;
    const float4 b = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
    if (tid == 0) {
        const float* q = p.gt_boxes + ((size_t)img * p.G + g) * 4;
        const float4 gb = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
        // data.py:103-121, then / std (model.py:489)
        const float h = __fsub_rn(b.z, b.x), w = __fsub_rn(b.w, b.y);
        const float cy = __fadd_rn(b.x, __fmul_rn(0.5f, h)), cx = __fadd_rn(b.y, __fmul_rn(0.5f, w));
        const float gh = __fsub_rn(gb.z, gb.x), gw = __fsub_rn(gb.w, gb.y);
        const float gcy = __fadd_rn(gb.x, __fmul_rn(0.5f, gh)), gcx = __fadd_rn(gb.y, __fmul_rn(0.5f, gw));
        float4 d;
        d.x = __fdiv_rn(__fdiv_rn(__fsub_rn(gcy, cy), h), p.std0);
        d.y = __fdiv_rn(__fdiv_rn(__fsub_rn(gcx, cx), w), p.std1);
        d.z = __fdiv_rn(log_cr(__fdiv_rn(gh, h)), p.std2);
        d.w = __fdiv_rn(log_cr(__fdiv_rn(gw, w)), p.std3);
        reinterpret_cast<float4*>(p.rois_out)[row] = b;
        p.class_out[row] = __ldg(p.gt_class + (size_t)img * p.G + g);
        reinterpret_cast<float4*>(p.deltas_out)[row] = d;
    }
    // model.py:492-507: crop the assigned gt mask to the RoI, resize to mh x mw, round half to even
    const float* mask = p.gt_masks + ((size_t)img * p.G + g) * p.H * p.W;
    for (int k = tid; k < npx; k += blockDim.x) {
        const int y = k / p.mw, x = k - y * p.mw;
        const AxisTap ty = axis_tap(b.x, b.z, p.H, p.mh, y);
        const AxisTap tx = axis_tap(b.y, b.w, p.W, p.mw, x);
        float v = 0.0f;  // extrapolation value 0
        if (ty.lo >= 0 && tx.lo >= 0) {
            const float* r0 = mask + (size_t)ty.lo * p.W;
            const float* r1 = mask + (size_t)ty.hi * p.W;
            v = bilerp(__ldg(r0 + tx.lo), __ldg(r0 + tx.hi), __ldg(r1 + tx.lo), __ldg(r1 + tx.hi), tx.lerp, ty.lerp);
        }
        mo[k] = rintf(v);
    }
}

static int next_pow2(int v) {
    int p = 32;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace mrcnn

using namespace mrcnn;

extern "C" {

int mrcnn_target_classify(const float* rois, const float* gt_boxes, const int32_t* gt_class_ids, int B, int N, int G,
                          int32_t* pos_idx, int32_t* neg_idx, int32_t* assign, float* iou_max, int32_t* counts,
                          mrcnn_stream_t stream) {
    MRCNN_REQUIRE(B > 0 && N >= 0 && G >= 0, "mrcnn_target_classify: bad sizes");
    MRCNN_REQUIRE((size_t)G * 20 <= 200 * 1024, "mrcnn_target_classify: at most %d gt boxes per image", 200 * 1024 / 20);
    MRCNN_REQUIRE_DEV(counts);
    if (N > 0) {
        MRCNN_REQUIRE_DEV(rois);
        MRCNN_REQUIRE_DEV(pos_idx);
        MRCNN_REQUIRE_DEV(neg_idx);
        MRCNN_REQUIRE_DEV(assign);
    }
    if (G > 0) {
        MRCNN_REQUIRE_DEV(gt_boxes);
        MRCNN_REQUIRE_DEV(gt_class_ids);
    }
    ClassifyParams p = {rois, gt_boxes, gt_class_ids, B, N, G, pos_idx, neg_idx, assign, iou_max, counts};
    const size_t smem = (size_t)G * 20 + 16;
    if (smem > 48 * 1024)
        MRCNN_CUDA(cudaFuncSetAttribute(target_classify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    target_classify_kernel<<<B, kTgtThreads, smem, (cudaStream_t)stream>>>(p);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

int mrcnn_target_select(const int32_t* counts, const float* keys_pos, const float* keys_neg, const int32_t* neg_table,
                        int B, int N, int pos_cap, int32_t* perm_pos, int32_t* perm_neg, int32_t* take,
                        mrcnn_stream_t stream) {
    MRCNN_REQUIRE(B > 0 && N > 0 && pos_cap >= 0, "mrcnn_target_select: bad sizes");
    MRCNN_REQUIRE(N <= kTgtMaxSort, "mrcnn_target_select: at most %d proposals per image", kTgtMaxSort);
    MRCNN_REQUIRE_DEV(counts);
    MRCNN_REQUIRE_DEV(keys_pos);
    MRCNN_REQUIRE_DEV(keys_neg);
    MRCNN_REQUIRE_DEV(neg_table);
    MRCNN_REQUIRE_DEV(perm_pos);
    MRCNN_REQUIRE_DEV(perm_neg);
    MRCNN_REQUIRE_DEV(take);
    SelectParams p = {counts, keys_pos, keys_neg, neg_table, B, N, pos_cap, next_pow2(N), perm_pos, perm_neg, take};
    const size_t smem = (size_t)p.P2 * (p.P2 == kTgtThreads ? 16 : 8);  // + exchange buffer of the register network
    if (smem > 48 * 1024)
        MRCNN_CUDA(cudaFuncSetAttribute(target_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    target_select_kernel<<<B, kTgtThreads, smem, (cudaStream_t)stream>>>(p);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

int mrcnn_target_emit(const float* rois, const float* gt_boxes, const int32_t* gt_class_ids, const float* gt_masks, int B,
                      int N, int G, int H, int W, const int32_t* pos_idx, const int32_t* neg_idx, const int32_t* perm_pos,
                      const int32_t* perm_neg, const int32_t* take, const int32_t* assign, const float* std4_host, int mask_h,
                      int mask_w, int T, float* rois_out, int32_t* class_out, float* deltas_out, float* masks_out,
                      mrcnn_stream_t stream) {
    MRCNN_REQUIRE(B > 0 && N > 0 && G > 0 && H > 0 && W > 0 && mask_h > 0 && mask_w > 0 && T >= 0 && std4_host,
                  "mrcnn_target_emit: bad sizes");
    if (T == 0) return MRCNN_OK;
    MRCNN_REQUIRE(B <= 65535, "mrcnn_target_emit: at most 65535 images per call");
    MRCNN_REQUIRE_DEV(rois);
    MRCNN_REQUIRE_DEV(gt_boxes);
    MRCNN_REQUIRE_DEV(gt_class_ids);
    MRCNN_REQUIRE_DEV(gt_masks);
    MRCNN_REQUIRE_DEV(pos_idx);
    MRCNN_REQUIRE_DEV(neg_idx);
    MRCNN_REQUIRE_DEV(take);
    MRCNN_REQUIRE_DEV(assign);
    MRCNN_REQUIRE_DEV(rois_out);
    MRCNN_REQUIRE_DEV(class_out);
    MRCNN_REQUIRE_DEV(deltas_out);
    MRCNN_REQUIRE_DEV(masks_out);
    if (perm_pos) MRCNN_REQUIRE_DEV(perm_pos);
    if (perm_neg) MRCNN_REQUIRE_DEV(perm_neg);
    EmitParams p = {rois, gt_boxes, gt_class_ids, gt_masks, B, N, G, H, W, pos_idx, neg_idx, perm_pos, perm_neg, take, assign,
                    std4_host[0], std4_host[1], std4_host[2], std4_host[3], mask_h, mask_w, T, rois_out, class_out, deltas_out,
                    masks_out};
    target_emit_kernel<<<dim3(T, B), 256, 0, (cudaStream_t)stream>>>(p);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

}  // extern "C"
