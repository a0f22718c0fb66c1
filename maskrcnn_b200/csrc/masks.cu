// masks.cu — mask paste-back (data.full_masks, data.py:287-314) for sm_100a.
//
// The reference turns the [D,81,28,28] mask-head output into [D,H,W] boolean image masks one detection at a time on the
// CPU: `.item()` / `.tolist()` syncs, a PIL image per detection (mask * 255 -> 'F' -> 'L'), torchvision Resize to the box
// size (Pillow's 8-bit two-pass bilinear resample), Pad to the image, '> 127', stack, copy back to the GPU.  Here the
// whole batch is two launches and the only HBM traffic that matters is the output itself (H*W bytes per detection, each
// written exactly once with 128-bit streaming stores):
//   paste_prepare_kernel  one CTA per detection: box decode, the 8-bit source mask, Pillow's resampling taps in double
//                         (22-bit fixed-point weights) for the box's rows and columns, and the HORIZONTAL pass of all
//                         mask rows into an 8-bit intermediate [mask_h][box width] - a few KB per detection, L2-resident;
//   full_masks_kernel     pure streaming, no shared memory, no barrier, no fp64: a CTA walks 128 image rows of one
//                         detection; rows / 16-byte chunks outside the box are zero stores, a chunk inside is <= 3
//                         128-bit loads of the intermediate (vertical taps) + 16 fixed-point MACs per tap, '> 127', store.
// First version (one kernel, taps + horizontal pass recomputed per 32-row band behind two barriers): 275 us for 800
// detections at 1024x1024; with a zero-band fast path 221 us; the barrier / fp64 chains of the ~25 % of bands that touch a
// box were the rest (a plain write-only stream reaches 6.2 TB/s, profiles/r01_bw_mix.txt).
//
// Arithmetic = Pillow's (Resample.c precompute_coeffs / normalize_coeffs_8bpc / ImagingResample{Horizontal,Vertical}_8bpc,
// Convert.c f2l), bit for bit: tap bounds and weights in double, 22-bit fixed-point weights, 8-bit intermediate image.
#include "api_util.h"
#include "common.cuh"

namespace mrcnn {

constexpr int kGroupRows = 64;   // image rows per CTA of the streaming kernel
constexpr int kPrepParts = 4;    // CTAs per detection in the prepare kernel
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

// Per-detection plan written by paste_prepare_kernel (all offsets in bytes from the detection's workspace slice).
struct PasteHeader {
    int live;        // box non-empty, class id valid, box intersects the image
    int y_lo, y_hi;  // visible rows
    int x_lo, x_hi;  // visible columns
    int left16;      // x_lo & ~15: column origin of the intermediate image
    int ksize;       // weight slots per row
    int pad;
};

struct PasteLayout {
    size_t rows_off;     // int4 {lo, n, 0, 0} per image row y            [H]
    size_t weights_off;  // int32 [visible row][ksize]
    size_t hrows_off;    // uint8 [mh][hstride]: horizontal pass of every mask row, column x stored at x - left16
    size_t det_bytes;
    int hstride;
};

__host__ __device__ inline PasteLayout paste_layout(int mh, int H, int W) {
    PasteLayout l;
    l.hstride = ((W + 15) / 16) * 16 + 16;
    l.rows_off = 32;
    l.weights_off = l.rows_off + (size_t)H * 16;
    // upscale: <= 3 taps per row; downscale (box shorter than the mask, < 64 rows): <= 2 * 64 + 1 taps per row
    const size_t wcap = (size_t)3 * H > (size_t)64 * 132 ? (size_t)3 * H : (size_t)64 * 132;
    l.hrows_off = l.weights_off + wcap * 4;
    l.hrows_off = (l.hrows_off + 15) / 16 * 16;
    l.det_bytes = (l.hrows_off + (size_t)mh * l.hstride + 255) / 256 * 256;
    return l;
}

struct PasteParams {
    const int64_t* class_ids;  // [D]
    const float* boxes;        // [D,4] px
    const float* masks;        // [D,NC,mh,mw]
    int D, NC, mh, mw, H, W, groups;
    unsigned char* ws;
    uint8_t* out;              // [D,H,W]
    int* err;
};

__global__ void __launch_bounds__(kMaskThreads) paste_prepare_kernel(const PasteParams p) {
    __shared__ uint8_t s_src[kMaxMaskSide * kMaxMaskSide];
    __shared__ int s_acc[kMaxMaskSide * kMaxMaskSide];  // [mask row][column], downscale path only
    const int tid = threadIdx.x, d = blockIdx.x / kPrepParts, part = blockIdx.x - d * kPrepParts;
    const PasteLayout L = paste_layout(p.mh, p.H, p.W);
    unsigned char* ws = p.ws + (size_t)d * L.det_bytes;
    // data.py:294-300: Python floats, int() truncates towards zero
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.boxes) + d);
    const int bh = (int)__dsub_rn((double)b.z, (double)b.x), bw = (int)__dsub_rn((double)b.w, (double)b.y);
    const int top = (int)b.x, left = (int)b.y;
    const long long cls = __ldg(p.class_ids + d);
    const bool cls_ok = cls >= 0 && cls < p.NC;
    // an empty box gives an empty mask (PIL raises ValueError; zero-padded detection rows land here)
    const int y_lo = max(0, top), y_hi = min(p.H, top + bh);
    const int x_lo = max(0, left), x_hi = min(p.W, left + bw);
    const bool live = cls_ok && bh > 0 && bw > 0 && y_lo < y_hi && x_lo < x_hi;
    const int left16 = x_lo & ~15;
    const double vscale = (double)p.mh / (double)max(bh, 1);
    const int ksize = 2 * (int)ceil(vscale < 1.0 ? 1.0 : vscale) + 1;
    if (tid == 0 && part == 0) {
        if (!cls_ok) atomicOr(p.err, 2);
        PasteHeader h = {live ? 1 : 0, y_lo, y_hi, x_lo, x_hi, left16, ksize, 0};
        *reinterpret_cast<PasteHeader*>(ws) = h;
    }
    if (!live || part * kMaskThreads >= (y_hi - y_lo) + (x_hi - x_lo)) return;  // nothing to do for this quarter
    // 'F' -> 'L' (Convert.c f2l) of mask * 255.0 (data.py:291)
    const float* m = p.masks + ((size_t)d * p.NC + (size_t)cls) * p.mh * p.mw;
    for (int i = tid; i < p.mh * p.mw; i += kMaskThreads) {
        const float v = __fmul_rn(__ldg(m + i), 255.0f);
        s_src[i] = v <= 0.0f ? 0 : (v >= 255.0f ? 255 : (uint8_t)(int)v);
    }
    __syncthreads();
    // work items of this detection: its visible rows (vertical taps) and its visible columns (horizontal taps + the
    // horizontal pass of every mask row into the 8-bit intermediate the vertical pass reads); a quarter per CTA
    int4* rows = reinterpret_cast<int4*>(ws + L.rows_off);
    int* weights = reinterpret_cast<int*>(ws + L.weights_off);
    uint8_t* hrows = ws + L.hrows_off;
    const int n_rows = y_hi - y_lo, n_items = n_rows + (x_hi - x_lo);
    for (int it = part * kMaskThreads + tid; it < n_items; it += kMaskThreads * kPrepParts) {
        if (it < n_rows) {
            const int y = y_lo + it;
            const AxisTaps t = axis_taps(p.mh, bh, y - top);
            rows[y] = make_int4(t.lo, t.n, 0, 0);
            int* w = weights + (size_t)it * ksize;
            if (t.n <= 3) {
                w[0] = t.k0;
                if (t.n > 1) w[1] = t.k1;
                if (t.n > 2) w[2] = t.k2;
            } else {
                for (int j = 0; j < t.n; ++j) w[j] = fixed_weight(t, j);
            }
        } else {
            const int x = x_lo + (it - n_rows);
            const AxisTaps t = axis_taps(p.mw, bw, x - left);
            uint8_t* dst = hrows + (x - left16);
            if (t.n <= 3) {
                for (int r = 0; r < p.mh; ++r) {
                    const uint8_t* s = s_src + r * p.mw + t.lo;
                    int acc = (1 << (kPrecBits - 1)) + (int)s[0] * t.k0;
                    if (t.n > 1) acc += (int)s[1] * t.k1;
                    if (t.n > 2) acc += (int)s[2] * t.k2;
                    dst[(size_t)r * L.hstride] = (uint8_t)clip8(acc);
                }
            } else {
                // downscale (box narrower than the mask, so fewer than kMaxMaskSide columns): tap-major, every fixed-point
                // weight (an fp64 division) is computed once and the per-row sums live in shared memory
                int* acc = s_acc + (x - x_lo);
                for (int r = 0; r < p.mh; ++r) acc[r * kMaxMaskSide] = 1 << (kPrecBits - 1);
                for (int j = 0; j < t.n; ++j) {
                    const int k = fixed_weight(t, j);
                    for (int r = 0; r < p.mh; ++r) acc[r * kMaxMaskSide] += (int)s_src[r * p.mw + t.lo + j] * k;
                }
                for (int r = 0; r < p.mh; ++r) dst[(size_t)r * L.hstride] = (uint8_t)clip8(acc[r * kMaxMaskSide]);
            }
        }
    }
}

template <bool kVec>
__global__ void __launch_bounds__(kMaskThreads) full_masks_kernel(const PasteParams p) {
    const int tid = threadIdx.x;
    const int d = blockIdx.x / p.groups, group = blockIdx.x - d * p.groups;
    const int r0 = group * kGroupRows, r1 = min(r0 + kGroupRows, p.H);
    const PasteLayout L = paste_layout(p.mh, p.H, p.W);
    const unsigned char* ws = p.ws + (size_t)d * L.det_bytes;
    const int4 h0 = __ldg(reinterpret_cast<const int4*>(ws)), h1 = __ldg(reinterpret_cast<const int4*>(ws) + 1);
    const int y_lo = h0.y, y_hi = h0.z, x_lo = h0.w, x_hi = h1.x, left16 = h1.y, ksize = h1.z;
    const bool live = h0.x != 0 && y_lo < r1 && y_hi > r0;  // the box touches this CTA's rows
    uint8_t* out = p.out + ((size_t)d * p.H + r0) * p.W;
    const int4* rows = reinterpret_cast<const int4*>(ws + L.rows_off);
    const int* weights = reinterpret_cast<const int*>(ws + L.weights_off);
    const uint8_t* hrows = ws + L.hrows_off;
    if (kVec) {
        const int chunks = p.W >> 4;  // 16-byte chunks per row
        if (!live) {
            // rows the box misses (three quarters of the CTAs at the bench size) are one contiguous block of zeros
            uint4* o = reinterpret_cast<uint4*>(out);
            const int n = (r1 - r0) * chunks;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            for (int i = tid; i < n; i += kMaskThreads) __stcs(o + i, z);
            return;
        }
        // (A) zero stores: whole rows above / below the box, then the chunks left / right of it
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        const int ya = max(y_lo, r0), yb = min(y_hi, r1);     // box rows of this CTA
        const int c_lo = x_lo >> 4, c_hi = (x_hi + 15) >> 4;  // chunks that intersect the box
        {
            uint4* o = reinterpret_cast<uint4*>(out);
            const int n_above = (ya - r0) * chunks;
            for (int i = tid; i < n_above; i += kMaskThreads) __stcs(o + i, z);
            uint4* o2 = o + (size_t)(yb - r0) * chunks;
            const int n_below = (r1 - yb) * chunks;
            for (int i = tid; i < n_below; i += kMaskThreads) __stcs(o2 + i, z);
            const int nz = chunks - (c_hi - c_lo);  // zero chunks per box row
            if (nz > 0) {
                uint4* o3 = o + (size_t)(ya - r0) * chunks;
                const int n_side = (yb - ya) * nz;
                for (int i = tid; i < n_side; i += kMaskThreads) {
                    const int row = i / nz, j = i - row * nz;
                    __stcs(o3 + (size_t)row * chunks + (j < c_lo ? j : j - c_lo + c_hi), z);
                }
            }
        }
        // (B) the chunks inside the box, flattened over (row, chunk) so that every thread has live work
        const int ncl = c_hi - c_lo;
        const int n_live = (yb - ya) * ncl;
#pragma unroll 2
        for (int i = tid; i < n_live; i += kMaskThreads) {
            const int row = i / ncl, c = c_lo + (i - row * ncl);
            const int y = ya + row, x0 = c << 4;
            const int4 rt = __ldg(rows + y);
            const int* wrow = weights + (size_t)(y - y_lo) * ksize;
            const uint8_t* col = hrows + (size_t)rt.x * L.hstride + (x0 - left16);
            int acc[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = 1 << (kPrecBits - 1);
            for (int j = 0; j < rt.y; ++j) {
                const int k = __ldg(wrow + j);
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(col + (size_t)j * L.hstride));
                const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 16; ++q) acc[q] += (int)((w[q >> 2] >> ((q & 3) * 8)) & 0xffu) * k;
            }
            // clip8(acc) > 127 (data.py:307) <=> acc >= 128 << 22: the clamp to 255 cannot change the comparison
            unsigned w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int q = 0; q < 16; ++q) w[q >> 2] |= (acc[q] >= (128 << kPrecBits) ? 1u : 0u) << ((q & 3) * 8);
            if (x0 < x_lo || x0 + 16 > x_hi) {  // a chunk on the box's left / right edge: drop the columns outside
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (x0 + q < x_lo || x0 + q >= x_hi) w[q >> 2] &= ~(0xffu << ((q & 3) * 8));
            }
            __stcs(reinterpret_cast<uint4*>(out + (size_t)(y - r0) * p.W) + c, make_uint4(w[0], w[1], w[2], w[3]));
        }
    } else {
        const int n = (r1 - r0) * p.W;
        for (int idx = tid; idx < n; idx += kMaskThreads) {
            const int row = idx / p.W, x = idx - row * p.W;
            const int y = r0 + row;
            uint8_t o = 0;
            if (live && y >= y_lo && y < y_hi && x >= x_lo && x < x_hi) {
                const int4 rt = __ldg(rows + y);
                const int* wrow = weights + (size_t)(y - y_lo) * ksize;
                const uint8_t* col = hrows + (size_t)rt.x * L.hstride + (x - left16);
                int acc = 1 << (kPrecBits - 1);
                for (int j = 0; j < rt.y; ++j) acc += (int)__ldg(col + (size_t)j * L.hstride) * __ldg(wrow + j);
                o = clip8(acc) > 127;
            }
            out[idx] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// decode_masks (data.py:265-284): the [D,H,W] image masks back in the original frame.  The reference takes every mask
// to the CPU, PIL '1' -> 'L' (0 / 255), torchvision CenterCrop to the window, Resize = the same Pillow 8-bit two-pass
// bilinear resample, np.array, stack, copy back: uint8 [D,nh,nw], not thresholded.
//   decode_taps_kernel   the taps of every output column and row once per call (they do not depend on the detection):
//                        {first input position, count} + 22-bit fixed-point weights, in double as Resample.c does;
//   decode_masks_kernel  CTA = 256 output columns x 64 output rows of one mask: the HORIZONTAL pass of the input rows the
//                        tile's vertical taps reach are staged in shared memory, their HORIZONTAL pass goes into shared memory
//                        as the 8-bit intermediate Pillow keeps (so the double rounding is Pillow's), the vertical pass
//                        reads it 16 columns at a time and writes 128-bit
//                        streaming stores.  A tile whose input bytes are all equal (the inside and the outside of a mask:
//                        all but the edge tiles) is a constant fill.  HBM traffic = the window read once (+ ~2 rows of halo
//                        per tile) and the output written once.
// ------------------------------------------------------------------------------------------------
constexpr int kDecTX = 256;  // output columns per CTA (= threads)
constexpr int kDecTY = 64;   // output rows per CTA (32 / 64 / 128 measured 143 / 126 / 122 us on the bench case); halved for
                             // downscales until the intermediate rows fit in shared memory (DecodeParams::ty)

struct DecodeParams {
    const uint8_t* masks;  // [D,H,W] 'L' pixels, or bool bytes (src_bool: non-zero -> 255)
    int src_bool;
    int src_vec4;          // rows are 32-bit aligned: W % 4 == 0 and a 4-byte aligned base
    int D, H, W;
    int top, left, ch, cw;  // CenterCrop window
    int nh, nw;             // target size
    int kx, ky;             // weight slots per output column / row
    int ty;                 // output rows per CTA
    int rmax;               // rows of the shared-memory intermediate
    int cstride;            // row pitch of the staged input bytes
    const int2* xmeta;      // [nw] {lo, n}
    const int* xw;          // [nw][kx]
    const int2* ymeta;      // [nh]
    const int* yw;          // [nh][ky]
    uint8_t* out;           // [D,nh,nw]
};

__global__ void __launch_bounds__(256) decode_taps_kernel(int cw, int nw, int kx, int2* xmeta, int* xw, int ch, int nh, int ky,
                                                          int2* ymeta, int* yw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nw + nh) return;
    const bool is_x = i < nw;
    const int pos = is_x ? i : i - nw;
    const AxisTaps t = is_x ? axis_taps(cw, nw, pos) : axis_taps(ch, nh, pos);
    int* w = is_x ? xw + (size_t)pos * kx : yw + (size_t)pos * ky;
    (is_x ? xmeta : ymeta)[pos] = make_int2(t.lo, t.n);
    if (t.n <= 3) {
        w[0] = t.k0;
        if (t.n > 1) w[1] = t.k1;
        if (t.n > 2) w[2] = t.k2;
    } else {
        for (int j = 0; j < t.n; ++j) w[j] = fixed_weight(t, j);
    }
}

// Sums of this file's resampling passes never clip: the weights of a position are >= 0 and sum to 2^22 + e with |e| <= half
// the tap count, so 2^21 + sum(v * k) < 256 * 2^22 for 8-bit v and clip8() is a plain shift.  With the weights scaled by 4
// the result is the top byte of an unsigned 32-bit sum: (2^23 + sum(v * 4k)) >> 24.
template <bool kVec>
__global__ void __launch_bounds__(kDecTX) decode_masks_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) uint8_t s_dec[];
    uint8_t* s_tmp = s_dec;                               // [rmax][kDecTX]   horizontal pass (Pillow's 8-bit intermediate)
    uint8_t* s_src = s_dec + (size_t)p.rmax * kDecTX;     // [rmax][cstride]  the input bytes under the tile
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kDecTX, y0 = blockIdx.y * p.ty, d = blockIdx.z;
    const int y1 = min(y0 + p.ty, p.nh);
    // input rows / columns (window coordinates) the tile's taps reach: lo and lo + n are non-decreasing along an axis
    const int2 mfirst = __ldg(p.ymeta + y0), mlast = __ldg(p.ymeta + (y1 - 1));
    const int r0 = mfirst.x, r1 = mlast.x + mlast.y;
    const int x_last = min(x0 + kDecTX, p.nw) - 1;
    const int2 cx0 = __ldg(p.xmeta + x0), cx1 = __ldg(p.xmeta + x_last);
    const int c0 = cx0.x, c1 = cx1.x + cx1.y;
    const uint8_t* rows = p.masks + ((size_t)d * p.H + p.top) * p.W;  // window row r, absolute column a: rows[r * W + a]
    uint8_t* out = p.out + (size_t)d * p.nh * p.nw;
    // staged columns [s0, s1) in absolute coordinates: the tap range [a0, a1), widened to whole 32-bit words when the rows
    // are word-aligned (W % 4 == 0), so that the staging loop moves four pixels per instruction
    const int a0 = p.left + c0, a1 = p.left + c1;
    const int s0 = p.src_vec4 ? (a0 & ~3) : a0;
    {   // stage the input bytes (coalesced), and probe them: a mask is two flat regions and an edge, and if every byte under
        // the tile's taps is the same value v both passes return v - the tile is a constant fill
        int v0 = __ldg(rows + (size_t)r0 * p.W + a0);
        if (p.src_bool) v0 = v0 ? 255 : 0;
        unsigned differs = 0;
        if (p.src_vec4) {
            const int nwords = (((a1 + 3) & ~3) - s0) >> 2;
            const unsigned v0w = (unsigned)v0 * 0x01010101u;
            auto consume = [&](unsigned w, int r, int wi) {
                if (p.src_bool) w = __vcmpne4(w, 0u);  // per byte: non-zero -> 0xff
                reinterpret_cast<unsigned*>(s_src + (r - r0) * p.cstride)[wi] = w;
                unsigned diff = w ^ v0w;
                const int col = s0 + 4 * wi;  // the bytes of this word that lie outside [a0, a1) do not count
                if (col < a0) diff &= 0xffffffffu << (8 * (a0 - col));
                if (col + 4 > a1) diff &= 0xffffffffu >> (8 * (col + 4 - a1));
                differs |= diff;
            };
            constexpr int kRI = 5, kWI = 2;  // rows / words per thread of the register-staged form
            if (nwords <= 32 * kWI && r1 - r0 <= (kDecTX / 32) * kRI) {
                // every load of the thread is issued before the first one is used (a warp stalls at the first use of a
                // load: the plain loop below keeps ONE load in flight per warp and the tile pays DRAM latency per row)
                unsigned w[kRI][kWI];
#pragma unroll
                for (int i = 0; i < kRI; ++i) {
                    const int r = r0 + (tid >> 5) + (kDecTX / 32) * i;
                    const unsigned* q = reinterpret_cast<const unsigned*>(rows + (size_t)min(r, r1 - 1) * p.W + s0);
#pragma unroll
                    for (int j = 0; j < kWI; ++j) {
                        const int wi = (tid & 31) + 32 * j;
                        w[i][j] = __ldg(q + min(wi, nwords - 1));
                    }
                }
#pragma unroll
                for (int i = 0; i < kRI; ++i) {
                    const int r = r0 + (tid >> 5) + (kDecTX / 32) * i;
#pragma unroll
                    for (int j = 0; j < kWI; ++j) {
                        const int wi = (tid & 31) + 32 * j;
                        if (r < r1 && wi < nwords) consume(w[i][j], r, wi);
                    }
                }
            } else {
                for (int r = r0 + (tid >> 5); r < r1; r += kDecTX / 32) {
                    const unsigned* q = reinterpret_cast<const unsigned*>(rows + (size_t)r * p.W + s0);
                    for (int wi = tid & 31; wi < nwords; wi += 32) consume(__ldg(q + wi), r, wi);
                }
            }
        } else {
            for (int r = r0 + (tid >> 5); r < r1; r += kDecTX / 32) {
                const uint8_t* q = rows + (size_t)r * p.W;
                uint8_t* sq = s_src + (r - r0) * p.cstride - s0;
                for (int a = a0 + (tid & 31); a < a1; a += 32) {
                    int v = __ldg(q + a);
                    if (p.src_bool) v = v ? 255 : 0;
                    sq[a] = (uint8_t)v;
                    differs |= (unsigned)(v ^ v0);
                }
            }
        }
        if (!__syncthreads_or(differs)) {
            const unsigned w = (unsigned)v0 * 0x01010101u;
            if (kVec) {
                const int c = tid & 15, x = x0 + c * 16;
                if (x < p.nw)
                    for (int y = y0 + (tid >> 4); y < y1; y += kDecTX / 16)
                        __stcs(reinterpret_cast<uint4*>(out + (size_t)y * p.nw + x), make_uint4(w, w, w, w));
            } else {
                const int x = x0 + tid;
                if (x < p.nw)
                    for (int y = y0; y < y1; ++y) out[(size_t)y * p.nw + x] = (uint8_t)v0;
            }
            return;
        }
    }
    {   // horizontal pass: column x0 + tid of rows [r0, r1), from shared memory into shared memory
        const int x = x0 + tid;
        if (x < p.nw) {
            const int2 m = __ldg(p.xmeta + x);
            const int* w = p.xw + (size_t)x * p.kx;
            const uint8_t* s = s_src + (p.left + m.x - s0);
            if (m.y <= 3) {
                const unsigned k0 = (unsigned)__ldg(w) << 2, k1 = m.y > 1 ? (unsigned)__ldg(w + 1) << 2 : 0u,
                               k2 = m.y > 2 ? (unsigned)__ldg(w + 2) << 2 : 0u;
                const int o1 = m.y > 1 ? 1 : 0, o2 = m.y > 2 ? 2 : 0;  // a zero weight on a valid address
                for (int r = 0; r < r1 - r0; ++r) {
                    const uint8_t* q = s + r * p.cstride;
                    const unsigned acc = (1u << (kPrecBits + 1)) + q[0] * k0 + q[o1] * k1 + q[o2] * k2;
                    s_tmp[r * kDecTX + tid] = (uint8_t)(acc >> 24);
                }
            } else {
                for (int r = 0; r < r1 - r0; ++r) {
                    const uint8_t* q = s + r * p.cstride;
                    int acc = 1 << (kPrecBits - 1);
                    for (int j = 0; j < m.y; ++j) acc += (int)q[j] * __ldg(w + j);
                    s_tmp[r * kDecTX + tid] = (uint8_t)clip8(acc);
                }
            }
        }
    }
    __syncthreads();
    if (kVec) {  // nw % 16 == 0: 16 columns per thread, 16 rows per sweep
        const int c = tid & 15, x = x0 + c * 16;
        if (x < p.nw) {
            for (int y = y0 + (tid >> 4); y < y1; y += kDecTX / 16) {
                const int2 m = __ldg(p.ymeta + y);
                const int* w = p.yw + (size_t)y * p.ky;
                const uint8_t* col = s_tmp + (m.x - r0) * kDecTX + c * 16;
                unsigned acc[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) acc[q] = 1u << (kPrecBits + 1);
                for (int j = 0; j < m.y; ++j) {
                    const unsigned k = (unsigned)__ldg(w + j) << 2;
                    if (k == 0u) continue;  // an upscale has three taps per row and often a zero weight among them
                    const uint4 v = *reinterpret_cast<const uint4*>(col + j * kDecTX);
                    const unsigned vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 16; ++q) acc[q] += __byte_perm(vw[q >> 2], 0u, 0x4440u + (q & 3)) * k;
                }
                unsigned o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)  // the top bytes of four sums
                    o[q] = __byte_perm(__byte_perm(acc[4 * q], acc[4 * q + 1], 0x0073u), __byte_perm(acc[4 * q + 2], acc[4 * q + 3], 0x0073u), 0x5410u);
                __stcs(reinterpret_cast<uint4*>(out + (size_t)y * p.nw + x), make_uint4(o[0], o[1], o[2], o[3]));
            }
        }
    } else {  // any width: one column per thread, byte stores (a warp still writes 32 consecutive bytes)
        const int x = x0 + tid;
        if (x < p.nw) {
            for (int y = y0; y < y1; ++y) {
                const int2 m = __ldg(p.ymeta + y);
                const int* w = p.yw + (size_t)y * p.ky;
                const uint8_t* col = s_tmp + (m.x - r0) * kDecTX + tid;
                int acc = 1 << (kPrecBits - 1);
                for (int j = 0; j < m.y; ++j) acc += (int)col[j * kDecTX] * __ldg(w + j);
                out[(size_t)y * p.nw + x] = (uint8_t)clip8(acc);
            }
        }
    }
}

struct DecodeLayout {
    size_t xmeta_off, xw_off, ymeta_off, yw_off, bytes;
    int kx, ky, ty, rmax, cstride;
};

// Resample.c: ksize = (int)ceil(support) * 2 + 1 with support = max(in / out, 1) for the triangle filter
static int decode_ksize(int in_size, int out_size) {
    const double scale = (double)in_size / (double)out_size;
    return (int)ceil(scale < 1.0 ? 1.0 : scale) * 2 + 1;
}

static DecodeLayout decode_layout(int ch, int cw, int nh, int nw) {
    DecodeLayout l;
    l.kx = decode_ksize(cw, nw);
    l.ky = decode_ksize(ch, nh);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) / 256 * 256;
        return at;
    };
    l.xmeta_off = take((size_t)nw * sizeof(int2));
    l.xw_off = take((size_t)nw * l.kx * sizeof(int));
    l.ymeta_off = take((size_t)nh * sizeof(int2));
    l.yw_off = take((size_t)nh * l.ky * sizeof(int));
    l.bytes = off;
    // input rows under ty output rows: the centres span (ty - 1) * scale, the taps reach `support` either side
    const double sy = (double)ch / (double)nh;
    const double sx = (double)cw / (double)nw;
    l.cstride = (((int)ceil((kDecTX - 1) * sx) + l.kx + 2 + 8) + 15) / 16 * 16;  // + 8: widening to whole words
    l.ty = kDecTY;
    for (;;) {
        l.rmax = (int)ceil((l.ty - 1) * sy) + l.ky + 2;
        if ((size_t)l.rmax * (kDecTX + l.cstride) <= 96 * 1024 || l.ty <= 4) break;
        l.ty >>= 1;
    }
    return l;
}

}  // namespace mrcnn

using namespace mrcnn;

extern "C" {

size_t mrcnn_full_masks_workspace_bytes(int D, int mask_h, int mask_w, int H, int W) {
    (void)mask_w;
    if (D <= 0 || mask_h <= 0 || H <= 0 || W <= 0) return 256;
    return paste_layout(mask_h, H, W).det_bytes * (size_t)D;
}

int mrcnn_full_masks(const int64_t* class_ids, const float* boxes, const float* masks, int D, int NC, int mask_h, int mask_w,
                     int H, int W, uint8_t* out, void* workspace, size_t workspace_bytes, mrcnn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRCNN_REQUIRE(D >= 0 && NC > 0 && H > 0 && W > 0, "mrcnn_full_masks: bad sizes");
    MRCNN_REQUIRE(mask_h > 0 && mask_w > 0 && mask_h <= kMaxMaskSide && mask_w <= kMaxMaskSide,
                  "mrcnn_full_masks: mask side must be in [1, 64]");
    if (D == 0) return MRCNN_OK;
    MRCNN_REQUIRE_DEV(class_ids);
    MRCNN_REQUIRE_DEV(boxes);
    MRCNN_REQUIRE_DEV(masks);
    MRCNN_REQUIRE_DEV(out);
    MRCNN_REQUIRE_DEV(workspace);
    MRCNN_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15u) == 0, "mrcnn_full_masks: boxes must be 16-byte aligned");
    if (workspace_bytes < mrcnn_full_masks_workspace_bytes(D, mask_h, mask_w, H, W) || (reinterpret_cast<uintptr_t>(workspace) & 255u))
        return fail(MRCNN_E_WORKSPACE, "mrcnn_full_masks: workspace too small or not 256-byte aligned");
    PasteParams p;
    p.class_ids = class_ids; p.boxes = boxes; p.masks = masks;
    p.D = D; p.NC = NC; p.mh = mask_h; p.mw = mask_w; p.H = H; p.W = W;
    p.groups = (H + kGroupRows - 1) / kGroupRows;
    p.ws = reinterpret_cast<unsigned char*>(workspace);
    p.out = out;
    p.err = device_error_word();
    MRCNN_REQUIRE(p.err != nullptr, "cannot allocate device error word");
    MRCNN_REQUIRE((long long)D * (p.groups > kPrepParts ? p.groups : kPrepParts) < (1ll << 31), "mrcnn_full_masks: too many detections");
    paste_prepare_kernel<<<D * kPrepParts, kMaskThreads, 0, stream>>>(p);
    MRCNN_LAUNCH_CHECK();
    const bool vec = (W % 16 == 0) && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    if (vec)
        full_masks_kernel<true><<<D * p.groups, kMaskThreads, 0, stream>>>(p);
    else
        full_masks_kernel<false><<<D * p.groups, kMaskThreads, 0, stream>>>(p);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

size_t mrcnn_decode_masks_workspace_bytes(int crop_h, int crop_w, int out_h, int out_w) {
    if (crop_h <= 0 || crop_w <= 0 || out_h <= 0 || out_w <= 0) return 256;
    return decode_layout(crop_h, crop_w, out_h, out_w).bytes;
}

int mrcnn_decode_masks(const uint8_t* masks, int src_is_bool, int D, int H, int W, int top, int left, int crop_h, int crop_w,
                       int out_h, int out_w, uint8_t* out, void* workspace, size_t workspace_bytes, mrcnn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRCNN_REQUIRE(D >= 0 && H > 0 && W > 0, "mrcnn_decode_masks: bad sizes");
    MRCNN_REQUIRE(crop_h > 0 && crop_w > 0 && top >= 0 && left >= 0 && top + crop_h <= H && left + crop_w <= W,
                  "mrcnn_decode_masks: the crop window must lie inside the mask");
    MRCNN_REQUIRE(out_h > 0 && out_w > 0, "mrcnn_decode_masks: height and width must be > 0");
    MRCNN_REQUIRE(D <= 65535, "mrcnn_decode_masks: at most 65535 masks per call");
    if (D == 0) return MRCNN_OK;
    MRCNN_REQUIRE_DEV(masks);
    MRCNN_REQUIRE_DEV(out);
    MRCNN_REQUIRE_DEV(workspace);
    const DecodeLayout l = decode_layout(crop_h, crop_w, out_h, out_w);
    if (workspace_bytes < l.bytes || (reinterpret_cast<uintptr_t>(workspace) & 255u))
        return fail(MRCNN_E_WORKSPACE, "mrcnn_decode_masks: workspace too small or not 256-byte aligned");
    const size_t smem = (size_t)l.rmax * (kDecTX + l.cstride);
    MRCNN_REQUIRE(smem <= 200 * 1024, "mrcnn_decode_masks: downscale factor too large (%d intermediate rows per tile)", l.rmax);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    DecodeParams p;
    p.masks = masks; p.src_bool = src_is_bool ? 1 : 0;
    p.src_vec4 = (W % 4 == 0 && (reinterpret_cast<uintptr_t>(masks) & 3u) == 0) ? 1 : 0;
    p.D = D; p.H = H; p.W = W;
    p.top = top; p.left = left; p.ch = crop_h; p.cw = crop_w;
    p.nh = out_h; p.nw = out_w;
    p.kx = l.kx; p.ky = l.ky; p.ty = l.ty; p.rmax = l.rmax; p.cstride = l.cstride;
    p.xmeta = reinterpret_cast<const int2*>(ws + l.xmeta_off);
    p.xw = reinterpret_cast<const int*>(ws + l.xw_off);
    p.ymeta = reinterpret_cast<const int2*>(ws + l.ymeta_off);
    p.yw = reinterpret_cast<const int*>(ws + l.yw_off);
    p.out = out;
    decode_taps_kernel<<<(out_w + out_h + 255) / 256, 256, 0, stream>>>(crop_w, out_w, l.kx, reinterpret_cast<int2*>(ws + l.xmeta_off),
                                                                    reinterpret_cast<int*>(ws + l.xw_off), crop_h, out_h, l.ky,
                                                                    reinterpret_cast<int2*>(ws + l.ymeta_off),
                                                                    reinterpret_cast<int*>(ws + l.yw_off));
    MRCNN_LAUNCH_CHECK();
    const dim3 grid((out_w + kDecTX - 1) / kDecTX, (out_h + l.ty - 1) / l.ty, D);
    MRCNN_REQUIRE(grid.y <= 65535, "mrcnn_decode_masks: target too tall");
    const bool vec = (out_w % 16 == 0) && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    if (vec) {
        if (smem > 48 * 1024) MRCNN_CUDA(cudaFuncSetAttribute(decode_masks_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        decode_masks_kernel<true><<<grid, kDecTX, smem, stream>>>(p);
    } else {
        if (smem > 48 * 1024) MRCNN_CUDA(cudaFuncSetAttribute(decode_masks_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        decode_masks_kernel<false><<<grid, kDecTX, smem, stream>>>(p);
    }
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

}  // extern "C"
