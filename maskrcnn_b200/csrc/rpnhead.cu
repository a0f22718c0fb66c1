// rpnhead.cu — RPN head output plumbing for sm_100a: everything between the two 1x1 convolutions of the RPN head and the
// proposal layer.
//
// The reference does this per pyramid level (model.py:624-641) and then across levels (model.rpn_detect, :1294-1304):
//     logits = conv_class(x)  [B,2K,H,W] -> permute(0,2,3,1).contiguous().view(B,-1,2);  rpn_class = softmax(logits, dim=2)
//     bbox   = conv_bbox(x)   [B,4K,H,W] -> permute(0,2,3,1).contiguous().view(B,-1,4)
//     torch.cat over the five levels, three times
// = 5 x (2 layout copies + softmax) + 3 concatenations, ~18 launches and two extra passes over every byte.  Here ONE
// launch reads each conv output once (NCHW as cuDNN's default, or channels-last, where the permute is free) and writes
// the concatenated [B,A,2] logits, [B,A,2] probabilities, [B,A,4] deltas - and the foreground probability alone as
// [B,A], which is all the proposal layer reads (mrcnn_proposal_layer_fg: half the bytes of its only pass over the scores).
//
// One thread per (image, level position): K anchors x (2 logits + 4 deltas) strided loads that coalesce across the warp
// (x is the fastest index), contiguous 8 / 16-byte stores.  Softmax over the (bg, fg) pair in torch's operation order
// (subtract the max, exp, sum, multiply by the reciprocal of the sum, SoftMaxKernel.cpp vec_softmax_lastdim) with a
// correctly rounded exp.
#include "api_util.h"
#include "common.cuh"

namespace mrcnn {

constexpr int kMaxRpnLevels = 8;

struct RpnPackParams {
    const float* logits[kMaxRpnLevels];  // [B,2K,H,W]
    const float* bbox[kMaxRpnLevels];    // [B,4K,H,W]
    int hw[kMaxRpnLevels];               // H*W of the level
    int pos_base[kMaxRpnLevels + 1];     // prefix sums of hw
    int L, B, K, layout;
    float* out_logits;  // [B,A,2] or null
    float* out_class;   // [B,A,2] or null
    float* out_bbox;    // [B,A,4] or null
    float* out_fg;      // [B,A]   or null
};

__global__ void __launch_bounds__(256) rpn_pack_kernel(const RpnPackParams p) {
    // outputs of one CTA are contiguous ranges (anchor index = K * global position index + k): they are staged in shared
    // memory and written back as whole 32-byte sectors - per-thread stores would leave 8 of every 24 bytes per
    // instruction, i.e. partial-sector writes that L2 has to fill from DRAM first (ncu: +37 MB of reads).
    extern __shared__ __align__(16) unsigned char rp_smem[];
    const int K = p.K, nA = 256 * K;
    float4* s_bbox = reinterpret_cast<float4*>(rp_smem);          // [256*K]
    float2* s_logits = reinterpret_cast<float2*>(s_bbox + nA);    // [256*K]
    float2* s_class = s_logits + nA;                              // [256*K]
    float* s_fg = reinterpret_cast<float*>(s_class + nA);         // [256*K]
    const int P = p.pos_base[p.L];
    const long long total = (long long)p.B * P;
    const long long first = (long long)blockIdx.x * 256;
    const long long gid = first + threadIdx.x;
    if (gid < total) {
        const int b = (int)(gid / P), pos = (int)(gid - (long long)b * P);
        int l = 0;
#pragma unroll
        for (int i = 1; i < kMaxRpnLevels; ++i)
            if (i < p.L && pos >= p.pos_base[i]) l = i;
        const int q = pos - p.pos_base[l], hw = p.hw[l];
        // element (channel c, position q) of image b
        const bool nhwc = p.layout == MRCNN_NHWC;
        const float* lg = p.logits[l] + (size_t)b * 2 * K * hw + (nhwc ? (size_t)q * 2 * K : (size_t)q);
        const float* bx = p.bbox[l] + (size_t)b * 4 * K * hw + (nhwc ? (size_t)q * 4 * K : (size_t)q);
        const size_t cs = nhwc ? 1 : (size_t)hw;  // channel stride
        for (int k = 0; k < K; ++k) {
            const int a = threadIdx.x * K + k;
            const float x0 = __ldg(lg + (size_t)(2 * k) * cs), x1 = __ldg(lg + (size_t)(2 * k + 1) * cs);
            s_logits[a] = make_float2(x0, x1);
            if (p.out_bbox)
                s_bbox[a] = make_float4(__ldg(bx + (size_t)(4 * k) * cs), __ldg(bx + (size_t)(4 * k + 1) * cs),
                                        __ldg(bx + (size_t)(4 * k + 2) * cs), __ldg(bx + (size_t)(4 * k + 3) * cs));
            if (p.out_class || p.out_fg) {
                // exp(x - max) of the larger logit is exp(0) = 1 exactly: only the smaller one needs the exp
                const float m = fmaxf(x0, x1);
                const float d0 = __fsub_rn(x0, m), d1 = __fsub_rn(x1, m);
                const float e0 = d0 == 0.0f ? 1.0f : exp_cr(d0), e1 = d1 == 0.0f ? 1.0f : exp_cr(d1);
                const float inv = __fdiv_rn(1.0f, __fadd_rn(e0, e1));
                const float p1 = __fmul_rn(e1, inv);
                s_class[a] = make_float2(__fmul_rn(e0, inv), p1);
                s_fg[a] = p1;
            }
        }
    }
    __syncthreads();
    const long long left = total - first;
    const int n = (int)(left < 256 ? left : 256) * K;  // anchors of this CTA
    const size_t a0 = (size_t)first * K;
    for (int i = threadIdx.x; i < n; i += 256) {
        if (p.out_logits) reinterpret_cast<float2*>(p.out_logits)[a0 + i] = s_logits[i];
        if (p.out_class) reinterpret_cast<float2*>(p.out_class)[a0 + i] = s_class[i];
        if (p.out_bbox) reinterpret_cast<float4*>(p.out_bbox)[a0 + i] = s_bbox[i];
        if (p.out_fg) p.out_fg[a0 + i] = s_fg[i];
    }
}

// Adjoint of the layout part (the losses read rpn_class_logits and rpn_bbox, model.py:1256-1262; the probabilities only
// feed the non-differentiable proposal layer): gradients [B,A,2] / [B,A,4] back to the conv outputs' shape and layout.
struct RpnUnpackParams {
    float* g_logits[kMaxRpnLevels];  // [B,2K,H,W] or null
    float* g_bbox[kMaxRpnLevels];    // [B,4K,H,W] or null
    int hw[kMaxRpnLevels];
    int pos_base[kMaxRpnLevels + 1];
    int L, B, K, layout;
    const float* grad_logits;  // [B,A,2] or null
    const float* grad_bbox;    // [B,A,4] or null
};

__global__ void __launch_bounds__(256) rpn_unpack_kernel(const RpnUnpackParams p) {
    const int P = p.pos_base[p.L];
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)p.B * P) return;
    const int b = (int)(gid / P), pos = (int)(gid - (long long)b * P);
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxRpnLevels; ++i)
        if (i < p.L && pos >= p.pos_base[i]) l = i;
    const int q = pos - p.pos_base[l], hw = p.hw[l], K = p.K;
    const bool nhwc = p.layout == MRCNN_NHWC;
    const size_t cs = nhwc ? 1 : (size_t)hw;
    const size_t a0 = ((size_t)b * P + pos) * K;
    if (p.grad_logits) {
        float* lg = p.g_logits[l] + (size_t)b * 2 * K * hw + (nhwc ? (size_t)q * 2 * K : (size_t)q);
        for (int k = 0; k < K; ++k) {
            const float2 g = __ldg(reinterpret_cast<const float2*>(p.grad_logits) + a0 + k);
            lg[(size_t)(2 * k) * cs] = g.x;
            lg[(size_t)(2 * k + 1) * cs] = g.y;
        }
    }
    if (p.grad_bbox) {
        float* bx = p.g_bbox[l] + (size_t)b * 4 * K * hw + (nhwc ? (size_t)q * 4 * K : (size_t)q);
        for (int k = 0; k < K; ++k) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.grad_bbox) + a0 + k);
            bx[(size_t)(4 * k) * cs] = g.x;
            bx[(size_t)(4 * k + 1) * cs] = g.y;
            bx[(size_t)(4 * k + 2) * cs] = g.z;
            bx[(size_t)(4 * k + 3) * cs] = g.w;
        }
    }
}

}  // namespace mrcnn

using namespace mrcnn;

extern "C" {

int mrcnn_rpn_pack(const float* const* logits, const float* const* bbox, const int* H, const int* W, int levels, int B,
                   int anchors_per_location, int layout, float* logits_out, float* class_out, float* bbox_out, float* fg_out,
                   mrcnn_stream_t stream) {
    MRCNN_REQUIRE(levels > 0 && levels <= kMaxRpnLevels && B > 0 && anchors_per_location > 0 && logits && bbox && H && W,
                  "mrcnn_rpn_pack: bad sizes (1..%d levels)", kMaxRpnLevels);
    MRCNN_REQUIRE(layout == MRCNN_NCHW || layout == MRCNN_NHWC, "mrcnn_rpn_pack: layout must be MRCNN_NCHW or MRCNN_NHWC");
    RpnPackParams p = {};
    p.L = levels; p.B = B; p.K = anchors_per_location; p.layout = layout;
    long long total = 0;
    for (int l = 0; l < levels; ++l) {
        MRCNN_REQUIRE(H[l] > 0 && W[l] > 0, "mrcnn_rpn_pack: empty level %d", l);
        MRCNN_REQUIRE_DEV(logits[l]);
        MRCNN_REQUIRE_DEV(bbox[l]);
        p.logits[l] = logits[l];
        p.bbox[l] = bbox[l];
        p.hw[l] = H[l] * W[l];
        p.pos_base[l] = (int)total;
        total += (long long)H[l] * W[l];
    }
    MRCNN_REQUIRE(total * anchors_per_location * B < (1ll << 31), "mrcnn_rpn_pack: too many anchors");
    p.pos_base[levels] = (int)total;
    if (logits_out) MRCNN_REQUIRE_DEV(logits_out);
    if (class_out) MRCNN_REQUIRE_DEV(class_out);
    if (bbox_out) MRCNN_REQUIRE_DEV(bbox_out);
    if (fg_out) MRCNN_REQUIRE_DEV(fg_out);
    MRCNN_REQUIRE(((reinterpret_cast<uintptr_t>(logits_out) | reinterpret_cast<uintptr_t>(class_out)) & 7u) == 0 &&
                      (reinterpret_cast<uintptr_t>(bbox_out) & 15u) == 0 && (reinterpret_cast<uintptr_t>(fg_out) & 3u) == 0,
                  "mrcnn_rpn_pack: outputs must be 8 / 8 / 16 / 4-byte aligned");
    p.out_logits = logits_out; p.out_class = class_out; p.out_bbox = bbox_out; p.out_fg = fg_out;
    const long long threads = total * B;
    const size_t smem = (size_t)256 * anchors_per_location * (16 + 8 + 8 + 4);
    MRCNN_REQUIRE(smem <= 200 * 1024, "mrcnn_rpn_pack: too many anchors per location");
    if (smem > 48 * 1024) MRCNN_CUDA(cudaFuncSetAttribute(rpn_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rpn_pack_kernel<<<(unsigned)((threads + 255) / 256), 256, smem, (cudaStream_t)stream>>>(p);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

int mrcnn_rpn_unpack(const float* grad_logits, const float* grad_bbox, const int* H, const int* W, int levels, int B,
                     int anchors_per_location, int layout, float* const* g_logits, float* const* g_bbox, mrcnn_stream_t stream) {
    MRCNN_REQUIRE(levels > 0 && levels <= kMaxRpnLevels && B > 0 && anchors_per_location > 0 && H && W,
                  "mrcnn_rpn_unpack: bad sizes (1..%d levels)", kMaxRpnLevels);
    MRCNN_REQUIRE(layout == MRCNN_NCHW || layout == MRCNN_NHWC, "mrcnn_rpn_unpack: layout must be MRCNN_NCHW or MRCNN_NHWC");
    MRCNN_REQUIRE((grad_logits == nullptr) == (g_logits == nullptr) && (grad_bbox == nullptr) == (g_bbox == nullptr),
                  "mrcnn_rpn_unpack: a gradient and its destinations go together");
    if (!grad_logits && !grad_bbox) return MRCNN_OK;
    RpnUnpackParams p = {};
    p.L = levels; p.B = B; p.K = anchors_per_location; p.layout = layout;
    long long total = 0;
    for (int l = 0; l < levels; ++l) {
        MRCNN_REQUIRE(H[l] > 0 && W[l] > 0, "mrcnn_rpn_unpack: empty level %d", l);
        if (g_logits) {
            MRCNN_REQUIRE_DEV(g_logits[l]);
            p.g_logits[l] = g_logits[l];
        }
        if (g_bbox) {
            MRCNN_REQUIRE_DEV(g_bbox[l]);
            p.g_bbox[l] = g_bbox[l];
        }
        p.hw[l] = H[l] * W[l];
        p.pos_base[l] = (int)total;
        total += (long long)H[l] * W[l];
    }
    MRCNN_REQUIRE(total * anchors_per_location * B < (1ll << 31), "mrcnn_rpn_unpack: too many anchors");
    p.pos_base[levels] = (int)total;
    if (grad_logits) MRCNN_REQUIRE_DEV(grad_logits);
    if (grad_bbox) MRCNN_REQUIRE_DEV(grad_bbox);
    MRCNN_REQUIRE((reinterpret_cast<uintptr_t>(grad_logits) & 7u) == 0 && (reinterpret_cast<uintptr_t>(grad_bbox) & 15u) == 0,
                  "mrcnn_rpn_unpack: gradients must be 8 / 16-byte aligned");
    p.grad_logits = grad_logits; p.grad_bbox = grad_bbox;
    const long long threads = total * B;
    rpn_unpack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

}  // extern "C"
