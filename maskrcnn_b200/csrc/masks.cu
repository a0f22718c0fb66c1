// masks.cu — mask paste-back (data.full_masks, data.py:287-314) for sm_100a.
//
// The reference turns the [D,81,28,28] mask-head output into [D,H,W] boolean image masks one detection at a time on the
// CPU: `.item()` / `.tolist()` syncs, a PIL image per detection (mask * 255 -> 'F' -> 'L'), torchvision Resize to the box
// size (Pillow's 8-bit two-pass bilinear resample), Pad to the image, '> 127', stack, copy back to the GPU.  Here the
// whole batch is one launch and the only HBM traffic that matters is the output itself (H*W bytes per detection, each
// written exactly once with 128-bit stores): a CTA walks kGroupBands bands of kBandRows image rows of one detection; for a
// band the box touches it keeps the 8-bit source mask, the horizontal pass of the source rows the band needs and the
// per-row vertical taps in shared memory and streams the band out; bands that miss the box are pure zero fill.
//
// Arithmetic = Pillow's (Resample.c precompute_coeffs / normalize_coeffs_8bpc / ImagingResample{Horizontal,Vertical}_8bpc,
// Convert.c f2l), bit for bit: tap bounds and weights in double, 22-bit fixed-point weights, 8-bit intermediate image.
#include "api_util.h"
#include "common.cuh"

namespace mrcnn {

constexpr int kBandRows = 32;
constexpr int kGroupBands = 4;  // bands per CTA
constexpr int kMaskThreads = 256;
constexpr int kPrecBits = 22;   // Resample.c: PRECISION_BITS = 32 - 8 - 2
constexpr int kMaxMaskSide = 64;

// Taps of one output position of one axis (Resample.c precompute_coeffs with the triangle filter, box = whole input).
struct AxisTaps {
    int lo, n;         // input positions [lo, lo + n)
    int k0, k1, k2;    // fixed-point weights when n <= 3 (every upscale)
    double center, ss, ww;
};

__device__ __forceinline__ double tri_weight(const AxisTaps& t, int j) {
    double x = __dmul_rn(__dadd_rn(__dsub_rn((double)(j + t.lo), t.center), 0.5), t.ss);
    if (x < 0.0) x = -x;
    return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
}

__device__ __forceinline__ int fixed_weight(const AxisTaps& t, int j) {
    double v = tri_weight(t, j);
    if (t.ww != 0.0) v = __ddiv_rn(v, t.ww);
    return (int)__dadd_rn(0.5, __dmul_rn(v, (double)(1 << kPrecBits)));  // weights are >= 0: normalize_coeffs_8bpc's + branch
}

__device__ __forceinline__ AxisTaps axis_taps(int in_size, int out_size, int xx) {
    AxisTaps t;
    t.k0 = t.k1 = t.k2 = 0;
    t.center = t.ss = t.ww = 0.0;
    if (in_size == out_size) {  // ImagingResample skips the pass (need_horizontal / need_vertical): identity
        t.lo = xx;
        t.n = 1;
        t.k0 = 1 << kPrecBits;
        return t;
    }
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;  // triangle filter support 1.0
    t.center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
    t.ss = __ddiv_rn(1.0, filterscale);
    int lo = (int)__dadd_rn(__dsub_rn(t.center, support), 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)__dadd_rn(__dadd_rn(t.center, support), 0.5);
    if (hi > in_size) hi = in_size;
    t.lo = lo;
    t.n = hi - lo;
    double ww = 0.0;
    for (int j = 0; j < t.n; ++j) ww = __dadd_rn(ww, tri_weight(t, j));
    t.ww = ww;
    if (t.n <= 3) {
        t.k0 = fixed_weight(t, 0);
        if (t.n > 1) t.k1 = fixed_weight(t, 1);
        if (t.n > 2) t.k2 = fixed_weight(t, 2);
    }
    return t;
}

__device__ __forceinline__ int clip8(int v) {
    v >>= kPrecBits;  // arithmetic shift, like Resample.c clip8
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

struct RowTaps {  // vertical taps of one band row, relative to the band's first needed source row
    int lo, n, k0, k1, k2;
};

struct PasteParams {
    const int64_t* class_ids;  // [D]
    const float* boxes;        // [D,4] px
    const float* masks;        // [D,NC,mh,mw]
    int D, NC, mh, mw, H, W, bands;  // bands = CTAs per detection
    int tmp_stride;            // bytes per row of the horizontal-pass buffer (multiple of 16)
    uint8_t* out;              // [D,H,W]
    int* err;
};

template <bool kVec>
__global__ void __launch_bounds__(kMaskThreads) full_masks_kernel(const PasteParams p) {
    extern __shared__ __align__(16) unsigned char mk_smem[];
    uint8_t* s_tmp = mk_smem;                                              // [mh][tmp_stride]
    uint8_t* s_src = mk_smem + (size_t)p.mh * p.tmp_stride;               // [mh*mw]
    __shared__ RowTaps s_rows[kBandRows];
    __shared__ AxisTaps s_slow[kBandRows];  // only read for rows with more than 3 taps (downscale)
    const int tid = threadIdx.x;
    const int d = blockIdx.x / p.bands, group = blockIdx.x - d * p.bands;

    // data.py:294-300: Python floats, int() truncates towards zero
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.boxes) + d);
    const int bh = (int)__dsub_rn((double)b.z, (double)b.x), bw = (int)__dsub_rn((double)b.w, (double)b.y);
    const int top = (int)b.x, left = (int)b.y;
    const long long cls = __ldg(p.class_ids + d);
    const bool cls_ok = cls >= 0 && cls < p.NC;
    if (!cls_ok && tid == 0 && group == 0) atomicOr(p.err, 2);
    const int x_lo = max(0, left), x_hi = min(p.W, left + bw);
    const int left16 = x_lo & ~15;  // the horizontal-pass buffer is aligned with the 16-byte output chunks
    bool src_staged = false;

    // a CTA walks kGroupBands bands of kBandRows rows: the box decode above is paid once per 128 KB of output
    for (int band = group * kGroupBands; band < min((group + 1) * kGroupBands, (p.H + kBandRows - 1) / kBandRows); ++band) {
    const int r0 = band * kBandRows;
    const int r1 = min(r0 + kBandRows, p.H);
    // an empty box gives an empty mask (PIL raises ValueError; zero-padded detection rows land here)
    const int y_lo = max(r0, top), y_hi = min(r1, top + bh);
    const bool live = cls_ok && bh > 0 && bw > 0 && y_lo < y_hi && x_lo < x_hi;
    uint8_t* out = p.out + ((size_t)d * p.H + r0) * p.W;

    if (live) {
        // 'F' -> 'L' (Convert.c f2l) of mask * 255.0 (data.py:291)
        if (!src_staged) {
            const float* m = p.masks + ((size_t)d * p.NC + (size_t)cls) * p.mh * p.mw;
            for (int i = tid; i < p.mh * p.mw; i += kMaskThreads) {
                const float v = __fmul_rn(__ldg(m + i), 255.0f);
                s_src[i] = v <= 0.0f ? 0 : (v >= 255.0f ? 255 : (uint8_t)(int)v);
            }
            src_staged = true;
        }
        // vertical taps of the band's rows
        if (tid < y_hi - y_lo) {
            const AxisTaps t = axis_taps(p.mh, bh, y_lo + tid - top);
            s_rows[tid] = {t.lo, t.n, t.k0, t.k1, t.k2};
            if (t.n > 3) s_slow[tid] = t;
        }
        __syncthreads();
        const int src_lo = s_rows[0].lo;
        const int src_hi = s_rows[y_hi - y_lo - 1].lo + s_rows[y_hi - y_lo - 1].n;
        // horizontal pass of source rows [src_lo, src_hi) for the visible columns -> 8-bit intermediate
        for (int x = x_lo + tid; x < x_hi; x += kMaskThreads) {
            const AxisTaps t = axis_taps(p.mw, bw, x - left);
            uint8_t* dst = s_tmp + (x - left16);
            for (int r = src_lo; r < src_hi; ++r) {
                const uint8_t* s = s_src + r * p.mw + t.lo;
                int acc = 1 << (kPrecBits - 1);
                if (t.n <= 3) {
                    acc += (int)s[0] * t.k0;
                    if (t.n > 1) acc += (int)s[1] * t.k1;
                    if (t.n > 2) acc += (int)s[2] * t.k2;
                } else {
                    for (int j = 0; j < t.n; ++j) acc += (int)s[j] * fixed_weight(t, j);
                }
                dst[(size_t)(r - src_lo) * p.tmp_stride] = (uint8_t)clip8(acc);
            }
        }
        __syncthreads();
    }

    if (kVec) {
        const int chunks = p.W >> 4;  // 16-byte chunks per row
        if (!live) {
            // a band the box misses (94 % of them at the bench size) is one contiguous block of zeros
            uint4* o = reinterpret_cast<uint4*>(out);
            const int n = (r1 - r0) * chunks;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            for (int i = tid; i < n; i += kMaskThreads) __stcs(o + i, z);
            continue;
        }
        // warp w owns rows w, w + 8, ...; a lane owns 16-byte chunks lane, lane + 32, ...: every byte is written once
        const int warp = tid >> 5, lane = tid & 31;
        const int src_lo = s_rows[0].lo;
        const int c_lo = x_lo >> 4, c_hi = (x_hi + 15) >> 4;  // chunks that intersect the box
        for (int y = r0 + warp; y < r1; y += kMaskThreads / 32) {
            uint4* orow = reinterpret_cast<uint4*>(out + (size_t)(y - r0) * p.W);
            const bool row_live = y >= y_lo && y < y_hi;
            RowTaps rt = {0, 0, 0, 0, 0};
            if (row_live) rt = s_rows[y - y_lo];
            for (int c = lane; c < chunks; c += 32) {
                uint4 o = make_uint4(0u, 0u, 0u, 0u);
                if (row_live && c >= c_lo && c < c_hi) {
                    const int x0 = c << 4;
                    int acc[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = 1 << (kPrecBits - 1);
                    const uint8_t* col = s_tmp + (size_t)(rt.lo - src_lo) * p.tmp_stride + (x0 - left16);
                    for (int j = 0; j < rt.n; ++j) {
                        const int k = rt.n <= 3 ? (j == 0 ? rt.k0 : (j == 1 ? rt.k1 : rt.k2)) : fixed_weight(s_slow[y - y_lo], j);
                        const uint4 v = *reinterpret_cast<const uint4*>(col + (size_t)j * p.tmp_stride);
                        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int i = 0; i < 16; ++i) acc[i] += (int)((w[i >> 2] >> ((i & 3) * 8)) & 0xffu) * k;
                    }
                    unsigned w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int x = x0 + i;
                        // clip8(acc) > 127 (data.py:307) <=> acc >= 128 << 22: the clamp to 255 cannot change the comparison
                        const unsigned bit = (x >= x_lo && x < x_hi && acc[i] >= (128 << kPrecBits)) ? 1u : 0u;
                        w[i >> 2] |= bit << ((i & 3) * 8);
                    }
                    o = make_uint4(w[0], w[1], w[2], w[3]);
                }
                __stcs(orow + c, o);
            }
        }
    } else {
        const int n = (r1 - r0) * p.W;
        for (int idx = tid; idx < n; idx += kMaskThreads) {
            const int row = idx / p.W, x = idx - row * p.W;
            const int y = r0 + row;
            uint8_t o = 0;
            if (live && y >= y_lo && y < y_hi && x >= x_lo && x < x_hi) {
                const RowTaps rt = s_rows[y - y_lo];
                const uint8_t* col = s_tmp + (size_t)(rt.lo - s_rows[0].lo) * p.tmp_stride + (x - left16);
                int acc = 1 << (kPrecBits - 1);
                for (int j = 0; j < rt.n; ++j) {
                    const int k = rt.n <= 3 ? (j == 0 ? rt.k0 : (j == 1 ? rt.k1 : rt.k2)) : fixed_weight(s_slow[y - y_lo], j);
                    acc += (int)col[(size_t)j * p.tmp_stride] * k;
                }
                o = clip8(acc) > 127;
            }
            out[idx] = o;
        }
    }
    if (live) __syncthreads();  // the next band overwrites s_rows / s_tmp
    }
}

}  // namespace mrcnn

using namespace mrcnn;

extern "C" {

int mrcnn_full_masks(const int64_t* class_ids, const float* boxes, const float* masks, int D, int NC, int mask_h, int mask_w,
                     int H, int W, uint8_t* out, mrcnn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRCNN_REQUIRE(D >= 0 && NC > 0 && H > 0 && W > 0, "mrcnn_full_masks: bad sizes");
    MRCNN_REQUIRE(mask_h > 0 && mask_w > 0 && mask_h <= kMaxMaskSide && mask_w <= kMaxMaskSide,
                  "mrcnn_full_masks: mask side must be in [1, 64]");
    if (D == 0) return MRCNN_OK;
    MRCNN_REQUIRE_DEV(class_ids);
    MRCNN_REQUIRE_DEV(boxes);
    MRCNN_REQUIRE_DEV(masks);
    MRCNN_REQUIRE_DEV(out);
    MRCNN_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15u) == 0, "mrcnn_full_masks: boxes must be 16-byte aligned");
    PasteParams p;
    p.class_ids = class_ids; p.boxes = boxes; p.masks = masks;
    p.D = D; p.NC = NC; p.mh = mask_h; p.mw = mask_w; p.H = H; p.W = W;
    p.bands = (H + kBandRows * kGroupBands - 1) / (kBandRows * kGroupBands);  // CTAs per detection
    p.tmp_stride = (int)align_up((size_t)W, 16) + 16;
    p.out = out;
    p.err = device_error_word();
    MRCNN_REQUIRE(p.err != nullptr, "cannot allocate device error word");
    MRCNN_REQUIRE((long long)D * p.bands < (1ll << 31), "mrcnn_full_masks: too many detections");
    const size_t smem = (size_t)mask_h * p.tmp_stride + align_up((size_t)mask_h * mask_w, 16);
    MRCNN_REQUIRE(smem <= 200 * 1024, "mrcnn_full_masks: image too wide for the shared-memory row buffer");
    const bool vec = (W % 16 == 0) && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    auto kern = vec ? full_masks_kernel<true> : full_masks_kernel<false>;
    if (smem > 48 * 1024) MRCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<D * p.bands, kMaskThreads, smem, stream>>>(p);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

}  // extern "C"
