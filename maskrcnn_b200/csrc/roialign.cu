// roialign.cu — crop_and_resize / PyramidROIAlign forward and scatter-add backward for sm_100a.
//
// Replaces (reference, /root/reference):
//   c++ext/maskrcnn/csrc/cpu/crop_cpu.cpp:13-164 / cuda/crop_cuda.cu:17-88     crop forward
//   c++ext/maskrcnn/csrc/cpu/crop_cpu.cpp:167-265 / cuda/crop_cuda.cu:90-170   crop backward
//   model.py:276-393 roi_align (level assignment + 4 per-level crops + cat/sort/gather) and its autograd.
//
// Two families of kernels:
//   * channels-last ("nhwc") kernels — the fast path.  One CTA = one RoI x 64 channels.  The box, its
//     pyramid level and the p+p axis taps are computed once and staged in shared memory; 16 lanes cover
//     the 64 channels with one 128-bit load per bilinear tap (a 256 B contiguous segment per tap), 16
//     bins are in flight per CTA.  NCHW outputs / output-gradients are transposed through a shared
//     memory tile so that global traffic is fully coalesced on both sides.  Backward aggregates, in
//     shared memory, all bins of a RoI that fall on the same feature-map column pair before issuing
//     128-bit vector reductions (red.global.add.v4.f32) to the channels-last gradient pyramid.
//   * strided ("generic") kernels — any layout, any C, one thread per output scalar.  Used for NCHW
//     feature maps and the C=1 mask-target crop (model.py:501-502).
//
// HBM-bound: algorithmic bytes per RoI = C*p*p*4 (output) + unique taps*C*4, see DESIGN.md.
#include <limits.h>

#include "api_util.h"
#include "nms_core.cuh"

namespace mrcnn {

struct PyrLevel {
    float* ptr;  // [B, H, W, C] (nhwc kernels) or layout-dependent (generic kernels)
    int H, W;
};

struct RoiParams {
    PyrLevel lv[4];
    int pyramid;  // 1: level chosen per box (P2..P5); 0: single image tensor lv[0]
    LevelRule rule;
    int B, C;
    const float* boxes;        // [N,4]
    const int32_t* box_index;  // [N] or null (all zero)
    int N;
    int ph, pw;
    float extrap;
    float* crops;  // forward: output; backward: incoming gradient (read-only)
    int32_t* levels_out;
    int* err;
};

constexpr int kChunk = 64;     // channels per CTA
constexpr int kLanes = 16;     // float4 lanes covering a chunk
constexpr int kSlots = 16;     // bins in flight per CTA
constexpr int kThreads = 256;  // kLanes * kSlots

struct RoiCtx {
    float* base;  // start of the selected image in the selected level
    int H, W;
    bool ok;
};

// Axis tap staged in shared memory with element offsets pre-multiplied (rows: y*W*C, cols: x*C).
struct __align__(16) TapS {
    int lo, hi;
    float lerp;
    int valid;
};

__device__ __forceinline__ RoiCtx select_level(const RoiParams& p, int n, float4& box) {
    box.x = __ldg(p.boxes + 4 * n + 0);
    box.y = __ldg(p.boxes + 4 * n + 1);
    box.z = __ldg(p.boxes + 4 * n + 2);
    box.w = __ldg(p.boxes + 4 * n + 3);
    const int bi = p.box_index ? __ldg(p.box_index + n) : 0;
    int l = 0;
    if (p.pyramid) l = roi_level(box.x, box.y, box.z, box.w, p.rule) - 2;
    const PyrLevel L = (l == 0) ? p.lv[0] : (l == 1) ? p.lv[1] : (l == 2) ? p.lv[2] : p.lv[3];
    RoiCtx c;
    c.H = L.H;
    c.W = L.W;
    c.ok = (unsigned)bi < (unsigned)p.B;
    c.base = L.ptr + (size_t)(c.ok ? bi : 0) * L.H * L.W * p.C;
    if (threadIdx.x == 0 && blockIdx.y == 0) {
        if (!c.ok) atomicOr(p.err, 1);
        if (p.levels_out) p.levels_out[n] = l + 2;
    }
    return c;
}

// Stage the ph + pw taps of this RoI.  Threads [0,ph) do rows, threads [64,64+pw) do columns.
__device__ __forceinline__ void stage_taps(const RoiParams& p, const RoiCtx& ctx, const float4 box, int ph, int pw,
                                           TapS* s_ty, TapS* s_tx) {
    const int tid = threadIdx.x;
    if (tid < ph) {
        const AxisTap t = axis_tap(box.x, box.z, ctx.H, ph, tid);
        TapS o;
        o.valid = (t.lo >= 0) && ctx.ok;
        o.lo = o.valid ? t.lo * ctx.W * p.C : 0;
        o.hi = o.valid ? t.hi * ctx.W * p.C : 0;
        o.lerp = t.lerp;
        s_ty[tid] = o;
    } else if (tid >= 64 && tid < 64 + pw) {
        const AxisTap t = axis_tap(box.y, box.w, ctx.W, pw, tid - 64);
        TapS o;
        o.valid = (t.lo >= 0) && ctx.ok;
        o.lo = o.valid ? t.lo * p.C : 0;
        o.hi = o.valid ? t.hi * p.C : 0;
        o.lerp = t.lerp;
        s_tx[tid - 64] = o;
    }
}

// Linear copy between a contiguous global [rows][P2] block and the padded shared tile [rows][P2pad],
// four consecutive elements per thread and iteration; (row, col) advance incrementally (one division
// per thread, not per element).  kToShared: global -> tile (backward), else tile -> global (forward).
template <bool kToShared>
__device__ __forceinline__ void tile_copy(float* tile, float* g, int total, int P2, int P2pad) {
    const int tid = threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(g) & 15u) == 0);
    const int step = kThreads * 4;
    const int step_r = step / P2, step_c = step - step_r * P2;
    int e = tid * 4;
    int r = e / P2, c = e - r * P2;
    for (; e < total; e += step) {
        int rr[4], cc[4];
        rr[0] = r;
        cc[0] = c;
#pragma unroll
        for (int j = 1; j < 4; ++j) {
            cc[j] = cc[j - 1] + 1;
            rr[j] = rr[j - 1];
            if (cc[j] == P2) {
                cc[j] = 0;
                rr[j] += 1;
            }
        }
        if (vec && e + 3 < total) {
            if (kToShared) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(g + e));
                tile[rr[0] * P2pad + cc[0]] = v.x;
                tile[rr[1] * P2pad + cc[1]] = v.y;
                tile[rr[2] * P2pad + cc[2]] = v.z;
                tile[rr[3] * P2pad + cc[3]] = v.w;
            } else {
                float4 v;
                v.x = tile[rr[0] * P2pad + cc[0]];
                v.y = tile[rr[1] * P2pad + cc[1]];
                v.z = tile[rr[2] * P2pad + cc[2]];
                v.w = tile[rr[3] * P2pad + cc[3]];
                __stcs(reinterpret_cast<float4*>(g + e), v);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (e + j < total) {
                    if (kToShared) tile[rr[j] * P2pad + cc[j]] = __ldcs(g + e + j);
                    else __stcs(g + e + j, tile[rr[j] * P2pad + cc[j]]);
                }
            }
        }
        r += step_r;
        c += step_c;
        if (c >= P2) {
            c -= P2;
            r += 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Forward, channels-last input.  grid = (N, ceil(C/64)), block = 256, dyn smem = 64*(P2|1)*4 (NCHW out).
// POOL > 0: compile-time pool size (7, 14); POOL == 0: runtime ph x pw.
// ------------------------------------------------------------------------------------------------
template <int POOL, bool kOutNHWC>
__global__ void __launch_bounds__(kThreads, 4) roialign_fwd_nhwc_kernel(const RoiParams p) {
    extern __shared__ __align__(16) float tile[];
    __shared__ TapS s_ty[kMaxPool];
    __shared__ TapS s_tx[kMaxPool];

    const int ph = POOL ? POOL : p.ph;
    const int pw = POOL ? POOL : p.pw;
    const int n = blockIdx.x;
    const int c0 = blockIdx.y * kChunk;
    const int tid = threadIdx.x;
    const int P2 = ph * pw;
    const int P2pad = P2 | 1;
    const int C = p.C;

    float4 box;
    const RoiCtx ctx = select_level(p, n, box);
    stage_taps(p, ctx, box, ph, pw, s_ty, s_tx);
    __syncthreads();

    const int lane = tid & (kLanes - 1);
    const int slot = tid >> 4;
    const int c = c0 + 4 * lane;
    const bool c_ok = c < C;  // C % 4 == 0 is guaranteed by the launcher
    const float* src = ctx.base + c;
    float* out_nhwc = p.crops + (size_t)n * P2 * C + c;

    // All per-bin addressing is 32-bit element offsets from `src` / `out_nhwc` (the launcher guarantees
    // H*W*C and P2*C fit in an int): one IMAD.WIDE per access instead of 64-bit pointer chains.
    if (c_ok) {
#pragma unroll 1
        for (int b0 = slot; b0 < P2; b0 += 2 * kSlots) {
            float4 tl[2], tr[2], bl[2], br[2];
            float xl[2], yl[2];
            bool inside[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int b = b0 + u * kSlots;
                inside[u] = false;
                if (b < P2) {
                    const int y = b / pw;
                    const int x = b - y * pw;
                    const TapS ty = s_ty[y];
                    const TapS tx = s_tx[x];
                    inside[u] = ty.valid && tx.valid;
                    if (inside[u]) {
                        tl[u] = ldg_f4(src + (ty.lo + tx.lo));
                        tr[u] = ldg_f4(src + (ty.lo + tx.hi));
                        bl[u] = ldg_f4(src + (ty.hi + tx.lo));
                        br[u] = ldg_f4(src + (ty.hi + tx.hi));
                        xl[u] = tx.lerp;
                        yl[u] = ty.lerp;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int b = b0 + u * kSlots;
                if (b >= P2) continue;
                float4 v;
                if (inside[u]) {
                    v.x = bilerp(tl[u].x, tr[u].x, bl[u].x, br[u].x, xl[u], yl[u]);
                    v.y = bilerp(tl[u].y, tr[u].y, bl[u].y, br[u].y, xl[u], yl[u]);
                    v.z = bilerp(tl[u].z, tr[u].z, bl[u].z, br[u].z, xl[u], yl[u]);
                    v.w = bilerp(tl[u].w, tr[u].w, bl[u].w, br[u].w, xl[u], yl[u]);
                } else {
                    v = make_float4(p.extrap, p.extrap, p.extrap, p.extrap);
                }
                if (kOutNHWC) {
                    stg_f4_stream(out_nhwc + b * C, v);
                } else {
                    float* t = tile + ((4 * lane) * P2pad + b);
                    t[0] = v.x;
                    t[P2pad] = v.y;
                    t[2 * P2pad] = v.z;
                    t[3 * P2pad] = v.w;
                }
            }
        }
    }
    if (!kOutNHWC) {
        __syncthreads();
        const int cc = min(kChunk, C - c0);
        // contiguous [cc][P2] block of the NCHW output
        tile_copy<false>(tile, p.crops + ((size_t)n * C + c0) * P2, cc * P2, P2, P2pad);
    }
}

// ------------------------------------------------------------------------------------------------
// Backward, channels-last gradient pyramid.  Same CTA shape as the forward.
//
// Column aggregation: for one output row y of the RoI, all pw bins share (y_lo, y_hi, y_lerp) and hit
// columns x_lo(x), x_hi(x), which are non-decreasing in x.  Each (slot, lane) owner walks the bins of
// its row (segment) in x order and keeps a running float4 sum per feature-map column; it flushes one
// red.v4 per DISTINCT column (times two rows) instead of four per bin.  When the RoI is up-sampled
// (fewer feature columns than bins, the common case for the 14x14 mask head) this removes most atomics.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void flush_col(float* row0, float* row1, int col_off, float4 acc, float yl, bool two_rows) {
    // d(top) = (1 - yl) * s, d(bottom) = yl * s with s = sum_x w_x * g  (crop_cpu.cpp:254-260, regrouped)
    const float w0 = __fsub_rn(1.0f, yl);
    red_add_f4(row0 + col_off, make_float4(acc.x * w0, acc.y * w0, acc.z * w0, acc.w * w0));
    if (two_rows) red_add_f4(row1 + col_off, make_float4(acc.x * yl, acc.y * yl, acc.z * yl, acc.w * yl));
}

template <int POOL, bool kGradNHWC>
__global__ void __launch_bounds__(kThreads) roialign_bwd_nhwc_kernel(const RoiParams p) {
    extern __shared__ __align__(16) float tile[];
    __shared__ TapS s_ty[kMaxPool];
    __shared__ TapS s_tx[kMaxPool];

    const int ph = POOL ? POOL : p.ph;
    const int pw = POOL ? POOL : p.pw;
    const int n = blockIdx.x;
    const int c0 = blockIdx.y * kChunk;
    const int tid = threadIdx.x;
    const int P2 = ph * pw;
    const int P2pad = P2 | 1;
    const int C = p.C;

    float4 box;
    const RoiCtx ctx = select_level(p, n, box);
    stage_taps(p, ctx, box, ph, pw, s_ty, s_tx);
    if (!kGradNHWC) {
        const int cc = min(kChunk, C - c0);
        tile_copy<true>(tile, p.crops + ((size_t)n * C + c0) * P2, cc * P2, P2, P2pad);
    }
    __syncthreads();
    if (!ctx.ok) return;

    const int lane = tid & (kLanes - 1);
    const int slot = tid >> 4;
    const int c = c0 + 4 * lane;
    if (c >= C) return;
    float* dst = ctx.base + c;
    const float* g_nhwc = p.crops + (size_t)n * P2 * C + c;

    // Work items = (output row, column segment).  With few rows (7x7) each row is split into segments so
    // that all 16 slots have work; aggregation then happens within a segment.
    const int segs = (ph >= kSlots) ? 1 : min(pw, kSlots / ph);
    const int items = ph * segs;
    for (int it = slot; it < items; it += kSlots) {
        const int y = it / segs;
        const int seg = it - y * segs;
        const int xbeg = (seg * pw) / segs;
        const int xend = ((seg + 1) * pw) / segs;
        const TapS ty = s_ty[y];
        if (!ty.valid) continue;
        float* row0 = dst + ty.lo;
        float* row1 = dst + ty.hi;
        const bool two_rows = ty.lerp != 0.0f;  // lerp == 0 <=> hi == lo: nothing goes to a second row
        int cur = -1;                           // element offset of the column accumulated in acc; nxt is the column after it
        bool has_nxt = false;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* tl_ = tile + ((4 * lane) * P2pad + y * pw);
        const float* gl_ = g_nhwc + (y * pw) * C;
        // software pipeline: the gradient of bin x + 1 is in flight while bin x is accumulated
        float4 g_next = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kGradNHWC && xbeg < xend) g_next = ldg_f4_stream(gl_ + xbeg * C);
        for (int x = xbeg; x < xend; ++x) {
            const TapS tx = s_tx[x];
            float4 g;
            if (kGradNHWC) {
                g = g_next;
                if (x + 1 < xend) g_next = ldg_f4_stream(gl_ + (x + 1) * C);
            } else {
                const float* t = tl_ + x;
                g = make_float4(t[0], t[P2pad], t[2 * P2pad], t[3 * P2pad]);
            }
            if (!tx.valid) continue;
            if (tx.lo != cur) {
                if (cur >= 0) {
                    flush_col(row0, row1, cur, acc, ty.lerp, two_rows);
                    if (has_nxt && tx.lo != cur + C) flush_col(row0, row1, cur + C, nxt, ty.lerp, two_rows);
                }
                acc = (cur >= 0 && has_nxt && tx.lo == cur + C) ? nxt : make_float4(0.f, 0.f, 0.f, 0.f);
                nxt = make_float4(0.f, 0.f, 0.f, 0.f);
                has_nxt = false;
                cur = tx.lo;
            }
            const float wl = __fsub_rn(1.0f, tx.lerp);
            acc.x += wl * g.x;
            acc.y += wl * g.y;
            acc.z += wl * g.z;
            acc.w += wl * g.w;
            if (tx.lerp != 0.0f) {  // <=> hi == lo + 1 column
                nxt.x += tx.lerp * g.x;
                nxt.y += tx.lerp * g.y;
                nxt.z += tx.lerp * g.z;
                nxt.w += tx.lerp * g.w;
                has_nxt = true;
            }
        }
        if (cur >= 0) {
            flush_col(row0, row1, cur, acc, ty.lerp, two_rows);
            if (has_nxt) flush_col(row0, row1, cur + C, nxt, ty.lerp, two_rows);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward as a GATHER (tile-owner) — channels-last gradients in, channels-last pyramid gradient out.
//
// Every 8x8 tile of every (image, level) gradient map is owned by one CTA, which sums the contributions
// of all RoIs of that (image, level) that reach the tile and writes each pixel exactly once.  No atomics,
// no separate zero fill, no read-modify-write: DRAM traffic is the algorithmic minimum (upstream gradient
// read once, gradient pyramid written once), and the summation order is fixed -> bit-reproducible.
//
//   pass 1a roi_bin_kernel (one CTA): key = image*4 + level per RoI, sorted (key, index) -> CSR lists.
//   pass 1b roi_taps_kernel: per RoI the ph + pw axis taps, the pixel footprint, the affine bin-position model.
//   pass 2  roialign_bwd_gather_kernel: grid = tiles (coarse levels first); see the kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kGTile = 8;       // tile side in feature-map pixels
constexpr int kGThreads = 256;
constexpr int kBinMaxN = 8192;  // RoIs per call (uint16 positions, per-tile hit list in shared memory)
constexpr size_t kBinSmemMax = 200 * 1024;

struct __align__(8) GTap {
    int lo;      // floor tap index; the ceil tap is lo + 1 iff lerp != 0.  Invalid: a large negative number.
    float lerp;
};

struct GatherParams {
    PyrLevel lv[4];
    int tiles_x[4];
    int tiles_n[4];   // tiles per image of each level
    int blk_base[4];  // blockIdx layout: level 3 first ... level 0 last; blk_base[i] = first block of the i-th group
    int B, C, N;
    int ph, pw;
    const float* grads;      // [N][ph*pw][C]
    const int32_t* rid;      // [N] RoI index at sorted position j ((image, level) major, index minor); -1 = skipped
    const int4* bbox;        // [N] {ylo, yhi, xlo, xhi} pixel footprint of position j (empty if ylo > yhi)
    const float4* rp;        // [N] {y p0, 1/y scale, x p0, 1/x scale}: sample position of bin b ~ p0 + b*scale
    const GTap* taps;        // [N][ph + pw]
    const int32_t* offsets;  // [4B + 1] positions of each (image, level) segment
};

// pass 1a: stable counting sort of the RoIs by key = image*4 + level (one CTA, 32 warps): warp w owns the
// contiguous index range [w*chunk, (w+1)*chunk) and walks it 32 RoIs at a time; __match_any_sync ranks equal
// keys inside a step, a per-(key, warp) histogram in shared memory carries the rank across steps, and one
// block scan over the (key-major, warp-minor) histogram turns ranks into positions - so rid lists every
// (image, level) segment in ascending RoI index.  RoIs with a bad box_index are left out (rid tail = -1).
constexpr int kBinWarps = 32;

__global__ void __launch_bounds__(1024) roi_bin_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ box_index,
                                                       int N, int B, LevelRule rule, int32_t* __restrict__ rid,
                                                       int32_t* __restrict__ offsets, int* err) {
    extern __shared__ __align__(16) unsigned char bin_smem[];
    const int K = 4 * B;
    int* hist = reinterpret_cast<int*>(bin_smem);                 // [K][32]
    uint32_t* packed = reinterpret_cast<uint32_t*>(hist + (size_t)K * kBinWarps);  // [N] key << 16 | rank in (key, warp)
    __shared__ int s_warp_tot[kBinWarps];
    const int tid = threadIdx.x, warp = tid >> 5, wl = tid & 31;
    for (int i = tid; i < K * kBinWarps; i += blockDim.x) hist[i] = 0;
    __syncthreads();

    const int chunk = ((N + kBinWarps - 1) / kBinWarps + 31) & ~31;
    const int nbeg = warp * chunk, nend = min(N, nbeg + chunk);
    for (int n0 = nbeg; n0 < nend; n0 += 32) {
        const int n = n0 + wl;
        int k = -1;
        if (n < nend) {
            const float y1 = __ldg(boxes + 4 * n), x1 = __ldg(boxes + 4 * n + 1);
            const float y2 = __ldg(boxes + 4 * n + 2), x2 = __ldg(boxes + 4 * n + 3);
            const int bi = box_index ? __ldg(box_index + n) : 0;
            if ((unsigned)bi < (unsigned)B) k = bi * 4 + roi_level(y1, x1, y2, x2, rule) - 2;
            else atomicOr(err, 1);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (k >= 0 && wl == leader) {
            base = hist[k * kBinWarps + warp];
            hist[k * kBinWarps + warp] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        if (n < nend) packed[n] = (k >= 0) ? (((uint32_t)k << 16) | (uint32_t)(base + __popc(peers & ((1u << wl) - 1u)))) : 0xffffffffu;
        __syncwarp();
    }
    __syncthreads();

    // exclusive scan of hist in (key, warp) order
    const int E = K * kBinWarps;
    const int per = (E + (int)blockDim.x - 1) / (int)blockDim.x;
    const int ebeg = min(E, tid * per), eend = min(E, ebeg + per);
    int local = 0;
    for (int e = ebeg; e < eend; ++e) local += hist[e];
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (wl >= o) incl += u;
    }
    if (wl == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int v = s_warp_tot[wl];
        int inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc2, o);
            if (wl >= o) inc2 += u;
        }
        s_warp_tot[wl] = inc2 - v;
    }
    __syncthreads();
    int run = s_warp_tot[warp] + incl - local;
    for (int e = ebeg; e < eend; ++e) {
        const int v = hist[e];
        hist[e] = run;
        run += v;
    }
    __syncthreads();
    const int total = s_warp_tot[kBinWarps - 1] + __shfl_sync(0xffffffffu, incl, 31);  // valid only in the last warp
    if (tid == blockDim.x - 1) offsets[K] = total;
    for (int k = tid; k < K; k += blockDim.x) offsets[k] = hist[k * kBinWarps];
    __shared__ int s_total;
    if (tid == blockDim.x - 1) s_total = total;
    __syncthreads();
    for (int n = tid; n < N; n += blockDim.x) {
        const uint32_t pk = packed[n];
        if (pk != 0xffffffffu) rid[hist[(pk >> 16) * kBinWarps + n / chunk] + (int)(pk & 0xffffu)] = n;
    }
    for (int j = s_total + tid; j < N; j += blockDim.x) rid[j] = -1;
}

// pass 1b: one 128-thread CTA per sorted position: the ph + pw taps, the pixel footprint and the
// affine sample-position model used to bound the bin search.
__global__ void __launch_bounds__(128) roi_taps_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ rid,
                                                       int N, int ph, int pw, LevelRule rule, PyrLevel l0, PyrLevel l1,
                                                       PyrLevel l2, PyrLevel l3, int4* __restrict__ bbox,
                                                       float4* __restrict__ rp, GTap* __restrict__ taps) {
    __shared__ int s_min[2], s_max[2];
    const int j = blockIdx.x;
    const int tid = threadIdx.x;
    const int r = __ldg(rid + j);
    if (tid < 2) {
        s_min[tid] = INT_MAX;
        s_max[tid] = INT_MIN;
    }
    __syncthreads();
    GTap* out = taps + (size_t)j * (ph + pw);
    if (r < 0) {
        if (tid == 0) {
            bbox[j] = make_int4(1, 0, 1, 0);
            rp[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int b = tid; b < ph + pw; b += blockDim.x) out[b] = GTap{-(1 << 30), 0.f};
        return;
    }
    const float y1 = __ldg(boxes + 4 * r), x1 = __ldg(boxes + 4 * r + 1);
    const float y2 = __ldg(boxes + 4 * r + 2), x2 = __ldg(boxes + 4 * r + 3);
    const int l = roi_level(y1, x1, y2, x2, rule) - 2;
    const PyrLevel L = (l == 0) ? l0 : (l == 1) ? l1 : (l == 2) ? l2 : l3;
    for (int b = tid; b < ph + pw; b += blockDim.x) {
        const bool is_row = b < ph;
        const AxisTap t = is_row ? axis_tap(y1, y2, L.H, ph, b) : axis_tap(x1, x2, L.W, pw, b - ph);
        GTap g;
        g.lo = (t.lo >= 0) ? t.lo : -(1 << 30);
        g.lerp = t.lerp;
        out[b] = g;
        if (t.lo >= 0) {
            atomicMin(&s_min[is_row ? 0 : 1], t.lo);
            atomicMax(&s_max[is_row ? 0 : 1], t.hi);
        }
    }
    __syncthreads();
    if (tid == 0) {
        bbox[j] = make_int4(s_min[0], s_max[0], s_min[1], s_max[1]);
        // position of bin b along an axis ~ a1*(size-1) + b * ((a2-a1)*(size-1)/(crop-1))  (crop_cpu.cpp:52-60)
        const float sy = (ph > 1) ? ((y2 - y1) * (float)(L.H - 1)) / (float)(ph - 1) : 0.f;
        const float sx = (pw > 1) ? ((x2 - x1) * (float)(L.W - 1)) / (float)(pw - 1) : 0.f;
        float4 q;
        q.x = y1 * (float)(L.H - 1);
        q.y = (ph > 1 && fabsf(sy) >= 0.01f) ? 1.0f / sy : 0.f;  // 0 => search all bins
        q.z = x1 * (float)(L.W - 1);
        q.w = (pw > 1 && fabsf(sx) >= 0.01f) ? 1.0f / sx : 0.f;
        rp[j] = q;
    }
}

// Candidate bins [b0, b1] whose sample can touch a pixel of [pix_lo, pix_hi] (position within 1 of it), with
// slack; every candidate is verified exactly against its tap afterwards, so over-inclusion is harmless.
__device__ __forceinline__ void bin_range(float p0, float inv, int pix_lo, int pix_hi, int nb, int& b0, int& b1) {
    b0 = 0;
    b1 = nb - 1;
    if (inv != 0.f && inv == inv) {
        const float u = ((float)(pix_lo - 1) - p0) * inv;
        const float v = ((float)(pix_hi + 1) - p0) * inv;
        const float lo = fminf(u, v) - 0.02f, hi = fmaxf(u, v) + 0.02f;
        if (lo == lo && hi == hi) {
            b0 = max(0, (int)ceilf(fmaxf(lo, -1.0f)));
            b1 = min(nb - 1, (int)floorf(fminf(hi, (float)nb)));
        }
    }
}

__device__ __forceinline__ GTap ldg_tap(const GTap* q) {
    const int2 v = __ldg(reinterpret_cast<const int2*>(q));
    GTap t;
    t.lo = v.x;
    t.lerp = __int_as_float(v.y);
    return t;
}

__device__ __forceinline__ float tap_weight(const GTap t, int pix) {
    // crop_cpu.cpp:254-260: (1 - lerp) goes to the floor tap, lerp to the ceil tap
    if (t.lo == pix) return __fsub_rn(1.0f, t.lerp);
    if (t.lo + 1 == pix) return t.lerp;  // lerp == 0 when the ceil tap coincides with the floor tap
    return 0.f;
}

// pass 2: one CTA per 8x8 tile of one (image, level) map.  256 threads = 8 warps; warp w owns tile ROW w, a lane
// carries 4 NV channels of all 8 pixels of that row in registers, so a warp covers 128 NV channels and every
// gradient load is a 512-byte contiguous warp access.  More channels take ceil(C / (128 NV)) passes.
//
// Work is enumerated from the RoI side, never searched from the pixel side:
//   plan      (CTA, once per batch of <= 32 RoIs reaching the tile) stage the row taps and the column items
//             {bin column offset, tile column, weights} and trim, per tile row, the range of bin rows whose taps
//             land on it and, for the tile, the range of bin columns whose taps land in its 8 columns;
//   generate  (warp, lane = RoI) expand (RoI, bin row, bin column) into a flat per-warp queue of work items
//             {gradient offset, tile column, wy * (1 - xl), wy * xl}; a warp scan places the items in RoI order;
//   consume   (warp) a branch-light pipelined loop: four items, four 128-bit loads in flight, then a warp-uniform
//             switch on the tile column adds the bin into the one or two accumulators it touches (no dynamic
//             register indexing, no wasted FMAs).
constexpr int kOBatchMax = 32;   // RoIs per plan = lanes of the generating warp
constexpr int kOQueue = 256;     // work items per warp queue; needs ph * pw <= kOQueue
constexpr int kOMaxPool = 16;
constexpr int kONoop = 9;

struct __align__(16) ColItem {
    int off;  // bin column * C (element offset inside a bin row)
    int idx;  // floor column relative to the tile + 1, 0..8 (kONoop: no contribution)
    float wa, wb;
};

struct __align__(16) QItem {
    int off;  // element offset of the bin in grads (32-bit by eligibility)
    int idx;  // as ColItem.idx
    float wa, wb;
};

template <int NV>
struct AccRow {
    float4 a[kGTile][NV];
};

template <int NV>
__device__ __forceinline__ void fma_px(float4 (&a)[NV], float w, const float4 (&v)[NV]) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        a[k].x = fmaf(w, v[k].x, a[k].x);
        a[k].y = fmaf(w, v[k].y, a[k].y);
        a[k].z = fmaf(w, v[k].z, a[k].z);
        a[k].w = fmaf(w, v[k].w, a[k].w);
    }
}

// Adds one bin into the row accumulators: tile column idx - 1 gets wa, column idx gets wb.
template <int NV>
__device__ __forceinline__ void owner_accumulate(AccRow<NV>& r, int idx, float wa, float wb, const float4 (&v)[NV]) {
    switch (idx) {
        case 0: fma_px<NV>(r.a[0], wb, v); break;
        case 1: fma_px<NV>(r.a[0], wa, v); fma_px<NV>(r.a[1], wb, v); break;
        case 2: fma_px<NV>(r.a[1], wa, v); fma_px<NV>(r.a[2], wb, v); break;
        case 3: fma_px<NV>(r.a[2], wa, v); fma_px<NV>(r.a[3], wb, v); break;
        case 4: fma_px<NV>(r.a[3], wa, v); fma_px<NV>(r.a[4], wb, v); break;
        case 5: fma_px<NV>(r.a[4], wa, v); fma_px<NV>(r.a[5], wb, v); break;
        case 6: fma_px<NV>(r.a[5], wa, v); fma_px<NV>(r.a[6], wb, v); break;
        case 7: fma_px<NV>(r.a[6], wa, v); fma_px<NV>(r.a[7], wb, v); break;
        case 8: fma_px<NV>(r.a[7], wa, v); break;
        default: break;
    }
}

// dynamic shared memory layout of the gather kernel
struct GatherSmem {
    QItem queue[kGThreads / 32][kOQueue + 4];
    unsigned char taps[kOBatchMax * (kOMaxPool * (sizeof(GTap) + sizeof(ColItem)))];  // per RoI: row taps, column items
    uint16_t hits[kBinMaxN];  // sorted positions of the RoIs that reach this tile, ascending
    short2 rng[kOBatchMax][kGTile + 1];  // per RoI: bin-row range per tile row, then the bin-column range
    int goff[kOBatchMax];                // rid * P2 * C
    int wcnt[kGThreads / 32];
};

template <int POOL, int NV, bool kAccumulate>  // NV float4 per lane and pixel
__global__ void __launch_bounds__(kGThreads, (NV == 1) ? 3 : 2) roialign_bwd_gather_kernel(const GatherParams p) {
    extern __shared__ __align__(16) unsigned char gather_smem_raw[];
    GatherSmem& S = *reinterpret_cast<GatherSmem*>(gather_smem_raw);

    const int ph = POOL ? POOL : p.ph;
    const int pw = POOL ? POOL : p.pw;
    const int P2 = ph * pw;
    const int ntap = ph + pw;
    const int row_bytes = kOMaxPool * (int)sizeof(GTap);
    const int roi_bytes = kOMaxPool * (int)(sizeof(GTap) + sizeof(ColItem));
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int wl = tid & 31;

    // which tile: coarse levels first (their tiles collect the most RoIs), image-major inside a level
    int b = blockIdx.x;
    const int l = (b < p.blk_base[1]) ? 3 : (b < p.blk_base[2]) ? 2 : (b < p.blk_base[3]) ? 1 : 0;
    b -= (l == 3) ? 0 : (l == 2) ? p.blk_base[1] : (l == 1) ? p.blk_base[2] : p.blk_base[3];
    const PyrLevel L = (l == 0) ? p.lv[0] : (l == 1) ? p.lv[1] : (l == 2) ? p.lv[2] : p.lv[3];
    const int ntx = (l == 0) ? p.tiles_x[0] : (l == 1) ? p.tiles_x[1] : (l == 2) ? p.tiles_x[2] : p.tiles_x[3];
    const int tiles_l = (l == 0) ? p.tiles_n[0] : (l == 1) ? p.tiles_n[1] : (l == 2) ? p.tiles_n[2] : p.tiles_n[3];
    const int img = b / tiles_l;
    const int t = b - img * tiles_l;
    const int tyi = t / ntx;
    const int ty0 = tyi * kGTile;
    const int tx0 = (t - tyi * ntx) * kGTile;
    const int H = L.H, W = L.W, C = p.C;

    const int lbeg = __ldg(p.offsets + img * 4 + l);
    const int lend = __ldg(p.offsets + img * 4 + l + 1);

    // ---- cull: which RoIs of this (image, level) reach the tile?  Ordered compaction keeps the ascending
    //      RoI index, i.e. a fixed summation order. ----
    int nh = 0;
    for (int next = lbeg; next < lend; next += kGThreads) {
        const int j = next + tid;
        bool hit = false;
        if (j < lend) {
            const int4 bb = __ldg(p.bbox + j);
            hit = bb.y >= ty0 && bb.x <= ty0 + kGTile - 1 && bb.w >= tx0 && bb.z <= tx0 + kGTile - 1;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (wl == 0) S.wcnt[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kGThreads / 32; ++w) {
            const int v = S.wcnt[w];
            if (w < warp) before += v;
            total += v;
        }
        if (hit) S.hits[nh + before + __popc(m & ((1u << wl) - 1u))] = (uint16_t)j;
        nh += total;
        __syncthreads();
    }

    const int y = ty0 + warp;  // this warp's feature-map row
    const bool replan = nh > kOBatchMax;
    bool planned = false;
    QItem* queue = S.queue[warp];
    for (int cbase = 0; cbase < C; cbase += 128 * NV) {
        // lane wl owns channels cbase + 4 wl + 128 k, k < NV: each k is one 512-byte warp access
        const int c = cbase + 4 * wl;
        AccRow<NV> acc;
#pragma unroll
        for (int j = 0; j < kGTile; ++j)
#pragma unroll
            for (int k = 0; k < NV; ++k) acc.a[j][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        bool live[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) live[k] = (c + 128 * k) < C;
        const float* gbase = p.grads + (live[0] ? c : 0);
        const bool row_ok = y < H && cbase < C;  // warp-uniform

        for (int h0 = 0; h0 < nh; h0 += kOBatchMax) {
            const int nb = min(kOBatchMax, nh - h0);
            if (replan || !planned) {
                // ---- plan ----
                if (planned) __syncthreads();  // everyone is done with the previous plan
                for (int i = tid; i < nb * ntap; i += kGThreads) {
                    const int h = i / ntap, bq = i - h * ntap;
                    const GTap tp = ldg_tap(p.taps + (size_t)S.hits[h0 + h] * ntap + bq);
                    unsigned char* base = S.taps + h * roi_bytes;
                    if (bq < ph) {
                        reinterpret_cast<GTap*>(base)[bq] = tp;
                    } else {
                        ColItem it;
                        const int j0 = tp.lo - tx0;
                        it.off = (bq - ph) * C;
                        it.idx = ((unsigned)(j0 + 1) <= (unsigned)kGTile) ? j0 + 1 : kONoop;
                        it.wa = __fsub_rn(1.0f, tp.lerp);
                        it.wb = tp.lerp;
                        reinterpret_cast<ColItem*>(base + row_bytes)[bq - ph] = it;
                    }
                }
                for (int i = tid; i < nb * (kGTile + 1); i += kGThreads) {
                    const int h = i / (kGTile + 1), line = i - h * (kGTile + 1);
                    const int j = S.hits[h0 + h];
                    const int4 bb = __ldg(p.bbox + j);
                    const float4 rp = __ldg(p.rp + j);
                    const GTap* tp = p.taps + (size_t)j * ntap;
                    int b0 = 1, b1 = 0;
                    if (line < kGTile) {
                        const int yy = ty0 + line;
                        if (yy >= bb.x && yy <= bb.y) {
                            bin_range(rp.x, rp.y, yy, yy, ph, b0, b1);
                            while (b0 <= b1 && tap_weight(ldg_tap(tp + b0), yy) == 0.f) ++b0;
                            while (b1 >= b0 && tap_weight(ldg_tap(tp + b1), yy) == 0.f) --b1;
                        }
                    } else {
                        bin_range(rp.z, rp.w, tx0, tx0 + kGTile - 1, pw, b0, b1);
                        // a column tap reaches the tile iff its floor column is in [tx0 - 1, tx0 + 7]
                        while (b0 <= b1 && (unsigned)(ldg_tap(tp + ph + b0).lo - tx0 + 1) > (unsigned)kGTile) ++b0;
                        while (b1 >= b0 && (unsigned)(ldg_tap(tp + ph + b1).lo - tx0 + 1) > (unsigned)kGTile) --b1;
                        S.goff[h] = __ldg(p.rid + j) * P2 * C;
                    }
                    S.rng[h][line] = make_short2((short)b0, (short)b1);
                }
                __syncthreads();
                planned = true;
            }
            if (!row_ok) continue;

            // ---- generate + consume, in rounds of as many RoIs as fit the queue (RoIs in ascending index order) ----
            int hdone = 0;
            while (hdone < nb) {
                const int h = hdone + wl;
                short2 ry = make_short2(1, 0), rx = make_short2(1, 0);
                if (h < nb) {
                    ry = S.rng[h][warp];
                    rx = S.rng[h][kGTile];
                }
                const int nby = max(0, ry.y - ry.x + 1), nbx = max(0, rx.y - rx.x + 1);
                const int cnt = nby * nbx;
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, incl, o);
                    if (wl >= o) incl += u;
                }
                // lanes whose items still fit; incl is non-decreasing, cnt <= kOQueue, so lane 0 always fits
                const unsigned fit = __ballot_sync(0xffffffffu, incl <= kOQueue);
                const int m = min(nb - hdone, __ffs(~fit) ? __ffs(~fit) - 1 : 32);
                const int n = __shfl_sync(0xffffffffu, incl, m - 1);
                if (wl < m && cnt > 0) {
                    int pos = incl - cnt;
                    const GTap* ty = reinterpret_cast<const GTap*>(S.taps + h * roi_bytes);
                    const ColItem* cols = reinterpret_cast<const ColItem*>(S.taps + h * roi_bytes + row_bytes);
                    const int goff = S.goff[h];
                    for (int by = ry.x; by <= ry.y; ++by) {
                        const float wy = tap_weight(ty[by], y);
                        const int rowoff = goff + (by * pw) * C;
                        for (int bx = rx.x; bx <= rx.y; ++bx) {
                            const ColItem ci = cols[bx];
                            QItem q;
                            q.off = rowoff + ci.off;
                            q.idx = (wy != 0.f) ? ci.idx : kONoop;
                            q.wa = wy * ci.wa;
                            q.wb = wy * ci.wb;
                            queue[pos++] = q;
                        }
                    }
                }
                if (wl < 4) {  // pad to a multiple of four with no-ops (offset 0 is a valid address)
                    QItem q;
                    q.off = 0; q.idx = kONoop; q.wa = 0.f; q.wb = 0.f;
                    queue[n + wl] = q;
                }
                __syncwarp();
#pragma unroll 1
                for (int i = 0; i < n; i += 4) {
                    QItem it[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) it[u] = queue[i + u];
                    float4 v[4][NV];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int k = 0; k < NV; ++k)
                            v[u][k] = live[k] ? ldg_f4(gbase + it[u].off + 128 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 4; ++u) owner_accumulate<NV>(acc, it[u].idx, it[u].wa, it[u].wb, v[u]);
                }
                __syncwarp();
                hdone += m;
            }
        }

        // ---- every pixel of the tile is written exactly once ----
        if (row_ok) {
            float* o = L.ptr + (((size_t)img * H + y) * W + tx0) * C + c;
#pragma unroll
            for (int j = 0; j < kGTile; ++j) {
                if (tx0 + j < W) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        if (!live[k]) continue;
                        float4 a = acc.a[j][k];
                        float* q = o + j * C + 128 * k;
                        if (kAccumulate) {
                            const float4 o0 = *reinterpret_cast<const float4*>(q);
                            a.x += o0.x; a.y += o0.y; a.z += o0.z; a.w += o0.w;
                        }
                        stg_f4_stream(q, a);
                    }
                }
            }
        }
    }
}

// Zero-fills up to four buffers in one launch (per-image slices of the gradient pyramid).
struct ZeroParams {
    float4* ptr[4];
    long long n4[4];  // float4 elements per buffer
};

__global__ void __launch_bounds__(256) zero_levels_kernel(const ZeroParams z) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        float4* q = z.ptr[l];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < z.n4[l]; i += stride) q[i] = zero;
    }
}

// ------------------------------------------------------------------------------------------------
// Generic strided kernels: one thread per output scalar, any layout.  Element strides in floats.
// ------------------------------------------------------------------------------------------------
struct Strides4 {
    long long n, c, h, w;
};

__device__ __forceinline__ Strides4 strides_of(int layout, int C, int H, int W) {
    Strides4 s;
    if (layout == MRCNN_NHWC) {
        s.n = (long long)H * W * C;
        s.c = 1;
        s.h = (long long)W * C;
        s.w = C;
    } else {
        s.n = (long long)C * H * W;
        s.c = (long long)H * W;
        s.h = W;
        s.w = 1;
    }
    return s;
}

struct GenericParams {
    RoiParams r;
    int image_layout;
    int crops_layout;
};

template <bool kBackward>
__global__ void __launch_bounds__(256) crop_generic_kernel(const GenericParams gp) {
    const RoiParams& p = gp.r;
    const long long total = (long long)p.N * p.C * p.ph * p.pw;
    const Strides4 so = strides_of(gp.crops_layout, p.C, p.ph, p.pw);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % p.pw);
        long long t = idx / p.pw;
        const int y = (int)(t % p.ph);
        t /= p.ph;
        const int c = (int)(t % p.C);
        const int n = (int)(t / p.C);
        const float y1 = __ldg(p.boxes + 4 * n + 0), x1 = __ldg(p.boxes + 4 * n + 1);
        const float y2 = __ldg(p.boxes + 4 * n + 2), x2 = __ldg(p.boxes + 4 * n + 3);
        const int bi = p.box_index ? __ldg(p.box_index + n) : 0;
        int l = 0;
        if (p.pyramid) l = roi_level(y1, x1, y2, x2, p.rule) - 2;
        const PyrLevel L = (l == 0) ? p.lv[0] : (l == 1) ? p.lv[1] : (l == 2) ? p.lv[2] : p.lv[3];
        float* cp = p.crops + n * so.n + c * so.c + y * so.h + x * so.w;
        const bool ok = (unsigned)bi < (unsigned)p.B;
        if (!ok) {
            if (c == 0 && y == 0 && x == 0) atomicOr(p.err, 1);
            if (!kBackward) *cp = p.extrap;
            continue;
        }
        if (!kBackward && p.levels_out && c == 0 && y == 0 && x == 0) p.levels_out[n] = l + 2;
        const AxisTap ty = axis_tap(y1, y2, L.H, p.ph, y);
        const AxisTap tx = axis_tap(x1, x2, L.W, p.pw, x);
        if (ty.lo < 0 || tx.lo < 0) {
            if (!kBackward) *cp = p.extrap;
            continue;
        }
        const Strides4 si = strides_of(gp.image_layout, p.C, L.H, L.W);
        float* img = L.ptr + bi * si.n + c * si.c;
        float* ptl = img + ty.lo * si.h + tx.lo * si.w;
        float* ptr = img + ty.lo * si.h + tx.hi * si.w;
        float* pbl = img + ty.hi * si.h + tx.lo * si.w;
        float* pbr = img + ty.hi * si.h + tx.hi * si.w;
        if (!kBackward) {
            *cp = bilerp(__ldg(ptl), __ldg(ptr), __ldg(pbl), __ldg(pbr), tx.lerp, ty.lerp);
        } else {
            // crop_cpu.cpp:254-260
            const float g = __ldg(cp);
            const float dtop = __fmul_rn(__fsub_rn(1.0f, ty.lerp), g);
            atomicAdd(ptl, __fmul_rn(__fsub_rn(1.0f, tx.lerp), dtop));
            if (tx.lerp != 0.0f) atomicAdd(ptr, __fmul_rn(tx.lerp, dtop));
            if (ty.lerp != 0.0f) {
                const float dbot = __fmul_rn(ty.lerp, g);
                atomicAdd(pbl, __fmul_rn(__fsub_rn(1.0f, tx.lerp), dbot));
                if (tx.lerp != 0.0f) atomicAdd(pbr, __fmul_rn(tx.lerp, dbot));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------

// The reference's level formula (model.py:331-338) as a function of q, with correctly rounded log2.
static int level_of_q(float q) {
    const float v = 4.0f + (float)log2((double)q);
    if (!(v == v) || isinf(v)) return 2;
    int lv = (int)nearbyintf(v);
    return lv < 2 ? 2 : (lv > 5 ? 5 : lv);
}

// Smallest positive float q with level_of_q(q) >= k (level_of_q is non-decreasing in q).
static float level_threshold(int k) {
    uint32_t lo = 0x00000001u, hi = 0x7f7fffffu;  // positive finite floats are ordered like their bits
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        float q;
        memcpy(&q, &mid, 4);
        if (level_of_q(q) >= k) hi = mid; else lo = mid + 1;
    }
    float q;
    memcpy(&q, &lo, 4);
    return q;
}

static LevelRule make_level_rule(float image_area) {
    static float t3 = 0.f, t4 = 0.f, t5 = 0.f;
    if (t3 == 0.f) {
        t3 = level_threshold(3);
        t4 = level_threshold(4);
        t5 = level_threshold(5);
    }
    LevelRule r;
    r.denom = 224.0f / sqrtf(image_area);  // model.py:335-336: 224.0 / torch.sqrt(image_area), fp32
    r.t3 = t3;
    r.t4 = t4;
    r.t5 = t5;
    return r;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int launch_roi(const RoiParams& p, int image_layout, int crops_layout, bool backward, cudaStream_t stream) {
    if (p.N == 0) return MRCNN_OK;
    bool fast = image_layout == MRCNN_NHWC && (p.C % 4) == 0 && p.ph <= kMaxPool && p.pw <= kMaxPool &&
                aligned16(p.crops);
    for (int l = 0; l < (p.pyramid ? 4 : 1); ++l) fast = fast && aligned16(p.lv[l].ptr);
    const int P2 = p.ph * p.pw;
    const size_t smem = (crops_layout == MRCNN_NHWC) ? 0 : sizeof(float) * kChunk * (size_t)(P2 | 1);
    if (fast && smem > 200 * 1024) fast = false;
    for (int l = 0; l < (p.pyramid ? 4 : 1); ++l)  // the vectorised kernels use 32-bit element offsets inside one image
        if ((long long)p.lv[l].H * p.lv[l].W * p.C >= (1ll << 31)) fast = false;
    if ((long long)P2 * p.C >= (1ll << 31)) fast = false;
    if (fast) {
        const dim3 grid(p.N, (p.C + kChunk - 1) / kChunk);
        if (grid.y > 65535) return fail(MRCNN_E_INVALID_ARG, "C too large");
#define MRCNN_LAUNCH_NHWC(KERNEL)                                                                          \
    do {                                                                                                   \
        if (smem > 48 * 1024)                                                                              \
            MRCNN_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        KERNEL<<<grid, kThreads, smem, stream>>>(p);                                                       \
    } while (0)
#define MRCNN_DISPATCH_POOL(NAME, FLAG)                                             \
    do {                                                                            \
        if (p.ph == 7 && p.pw == 7) MRCNN_LAUNCH_NHWC((NAME<7, FLAG>));             \
        else if (p.ph == 14 && p.pw == 14) MRCNN_LAUNCH_NHWC((NAME<14, FLAG>));     \
        else MRCNN_LAUNCH_NHWC((NAME<0, FLAG>));                                    \
    } while (0)
        if (!backward) {
            if (crops_layout == MRCNN_NHWC) MRCNN_DISPATCH_POOL(roialign_fwd_nhwc_kernel, true);
            else MRCNN_DISPATCH_POOL(roialign_fwd_nhwc_kernel, false);
        } else {
            if (crops_layout == MRCNN_NHWC) MRCNN_DISPATCH_POOL(roialign_bwd_nhwc_kernel, true);
            else MRCNN_DISPATCH_POOL(roialign_bwd_nhwc_kernel, false);
        }
#undef MRCNN_DISPATCH_POOL
#undef MRCNN_LAUNCH_NHWC
    } else {
        GenericParams gp;
        gp.r = p;
        gp.image_layout = image_layout;
        gp.crops_layout = crops_layout;
        const long long total = (long long)p.N * p.C * P2;
        long long blocks = (total + 255) / 256;
        const long long cap = (long long)sm_count() * 64;
        if (blocks > cap) blocks = cap;
        if (!backward) crop_generic_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(gp);
        else crop_generic_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(gp);
    }
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

static int launch_zero(float* const ptr[4], const size_t elems[4], int nbuf, cudaStream_t stream) {
    ZeroParams z = {};
    bool any = false, vec = true;
    for (int l = 0; l < nbuf; ++l) vec = vec && aligned16(ptr[l]) && (elems[l] % 4 == 0);
    if (!vec) {
        for (int l = 0; l < nbuf; ++l) MRCNN_CUDA(cudaMemsetAsync(ptr[l], 0, sizeof(float) * elems[l], stream));
        return MRCNN_OK;
    }
    for (int l = 0; l < nbuf; ++l) {
        z.ptr[l] = reinterpret_cast<float4*>(ptr[l]);
        z.n4[l] = (long long)(elems[l] / 4);
        any = any || elems[l] > 0;
    }
    if (!any) return MRCNN_OK;
    zero_levels_kernel<<<sm_count() * 8, 256, 0, stream>>>(z);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

struct GatherWorkspace {
    int32_t* rid;
    int4* bbox;
    float4* rp;
    GTap* taps;
    int32_t* offsets;
    size_t bytes;
};

static GatherWorkspace carve_gather(void* base, int B, int N, int pool) {
    GatherWorkspace w;
    const size_t n = (size_t)(N > 0 ? N : 1);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* q = base ? (void*)((char*)base + off) : nullptr;
        off += align_up(bytes, 256);
        return q;
    };
    w.rid = (int32_t*)take(n * 4);
    w.bbox = (int4*)take(n * 16);
    w.rp = (float4*)take(n * 16);
    w.taps = (GTap*)take(n * 2 * (size_t)pool * sizeof(GTap));
    w.offsets = (int32_t*)take((size_t)(4 * (size_t)B + 1) * 4);
    w.bytes = off;
    return w;
}

// true if the gather backward can serve this call
static bool gather_eligible(int B, int C, int N, int pool, int gfm_layout, int grads_layout, const float* grads,
                            float* const gfm[4], const void* workspace, size_t workspace_bytes) {
    if (gfm_layout != MRCNN_NHWC || grads_layout != MRCNN_NHWC) return false;
    if ((C % 4) != 0 || pool > kOMaxPool || N > kBinMaxN || N <= 0) return false;
    if ((long long)N * pool * pool * C >= (1ll << 31)) return false;  // 32-bit element offsets into grads
    if (workspace == nullptr || workspace_bytes < carve_gather(nullptr, B, N, pool).bytes) return false;
    if ((size_t)4 * B * kBinWarps * 4 + (size_t)N * 4 > kBinSmemMax || 4 * (long long)B >= 65536) return false;
    if (!aligned16(grads) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return false;
    for (int l = 0; l < 4; ++l)
        if (!aligned16(gfm[l])) return false;
    return true;
}

static int launch_bwd_gather(const float* grads, const int H[4], const int W[4], int B, int C, const float* boxes,
                             const int32_t* box_index, int N, int pool, float image_area, float* const gfm[4],
                             int accumulate, void* workspace, cudaStream_t stream) {
    const GatherWorkspace ws = carve_gather(workspace, B, N, pool);
    const LevelRule rule = make_level_rule(image_area);
    const size_t bin_smem = (size_t)4 * B * kBinWarps * 4 + (size_t)N * 4;
    MRCNN_CUDA(cudaFuncSetAttribute(roi_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBinSmemMax));
    roi_bin_kernel<<<1, 1024, bin_smem, stream>>>(boxes, box_index, N, B, rule, ws.rid, ws.offsets, device_error_word());
    MRCNN_LAUNCH_CHECK();
    GatherParams g = {};
    long long total = 0;
    for (int l = 0; l < 4; ++l) {
        g.lv[l] = {gfm[l], H[l], W[l]};
        g.tiles_x[l] = (W[l] + kGTile - 1) / kGTile;
        g.tiles_n[l] = g.tiles_x[l] * ((H[l] + kGTile - 1) / kGTile);
        total += (long long)g.tiles_n[l] * B;
    }
    MRCNN_REQUIRE(total < (1ll << 31), "mrcnn_pyramid_roi_align_backward: too many tiles");
    g.blk_base[0] = 0;  // groups: level 3, 2, 1, 0
    for (int i = 1; i < 4; ++i) g.blk_base[i] = g.blk_base[i - 1] + g.tiles_n[4 - i] * B;
    roi_taps_kernel<<<N, 128, 0, stream>>>(boxes, ws.rid, N, pool, pool, rule, g.lv[0], g.lv[1], g.lv[2], g.lv[3], ws.bbox,
                                          ws.rp, ws.taps);
    MRCNN_LAUNCH_CHECK();
    g.B = B; g.C = C; g.N = N;
    g.ph = pool; g.pw = pool;
    g.grads = grads; g.rid = ws.rid; g.bbox = ws.bbox; g.rp = ws.rp; g.taps = ws.taps; g.offsets = ws.offsets;
    const unsigned grid = (unsigned)total;
    const bool wide = C > 128;  // 8 channels per lane: one pass covers 256 channels
    const size_t smem = sizeof(GatherSmem);
#define MRCNN_LAUNCH_GATHER_K(KERNEL)                                                                       \
    do {                                                                                                    \
        MRCNN_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        KERNEL<<<grid, kGThreads, smem, stream>>>(g);                                                       \
    } while (0)
#define MRCNN_LAUNCH_GATHER(POOL)                                                                           \
    do {                                                                                                    \
        if (wide && accumulate) MRCNN_LAUNCH_GATHER_K((roialign_bwd_gather_kernel<POOL, 2, true>));         \
        else if (wide) MRCNN_LAUNCH_GATHER_K((roialign_bwd_gather_kernel<POOL, 2, false>));                 \
        else if (accumulate) MRCNN_LAUNCH_GATHER_K((roialign_bwd_gather_kernel<POOL, 1, true>));            \
        else MRCNN_LAUNCH_GATHER_K((roialign_bwd_gather_kernel<POOL, 1, false>));                           \
    } while (0)
    if (pool == 7) MRCNN_LAUNCH_GATHER(7);
    else if (pool == 14) MRCNN_LAUNCH_GATHER(14);
    else MRCNN_LAUNCH_GATHER(0);
#undef MRCNN_LAUNCH_GATHER_K
#undef MRCNN_LAUNCH_GATHER
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

static int check_layout(int v, const char* what) {
    if (v != MRCNN_NCHW && v != MRCNN_NHWC) return fail(MRCNN_E_INVALID_ARG, "%s must be MRCNN_NCHW or MRCNN_NHWC", what);
    return MRCNN_OK;
}

}  // namespace mrcnn

using namespace mrcnn;

extern "C" {

size_t mrcnn_pyramid_roi_align_backward_workspace_bytes(int B, int N, int pool) {
    if (B <= 0 || N < 0 || pool <= 0) return 256;
    return carve_gather(nullptr, B, N, pool).bytes;
}

int mrcnn_crop_forward(const float* image, int B, int C, int H, int W, int image_layout, const float* boxes,
                       const int32_t* box_index, int N, float extrapolation_value, int crop_h, int crop_w,
                       float* crops, int crops_layout, mrcnn_stream_t stream) {
    MRCNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "mrcnn_crop_forward: image dims must be positive");
    MRCNN_REQUIRE(N >= 0 && crop_h > 0 && crop_w > 0, "mrcnn_crop_forward: bad N / crop size");
    if (int rc = check_layout(image_layout, "image_layout")) return rc;
    if (int rc = check_layout(crops_layout, "crops_layout")) return rc;
    if (N == 0) return MRCNN_OK;
    MRCNN_REQUIRE_DEV(image);
    MRCNN_REQUIRE_DEV(boxes);
    MRCNN_REQUIRE_DEV(box_index);
    MRCNN_REQUIRE_DEV(crops);
    RoiParams p = {};
    p.lv[0] = {const_cast<float*>(image), H, W};
    p.pyramid = 0;
    p.B = B; p.C = C;
    p.boxes = boxes; p.box_index = box_index; p.N = N;
    p.ph = crop_h; p.pw = crop_w; p.extrap = extrapolation_value;
    p.crops = crops;
    p.err = device_error_word();
    MRCNN_REQUIRE(p.err != nullptr, "cannot allocate device error word");
    return launch_roi(p, image_layout, crops_layout, false, (cudaStream_t)stream);
}

int mrcnn_crop_backward(const float* grads, int grads_layout, const float* boxes, const int32_t* box_index, int N,
                        int crop_h, int crop_w, float* grads_image, int B, int C, int H, int W, int image_layout,
                        int zero_fill, mrcnn_stream_t stream) {
    MRCNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "mrcnn_crop_backward: image dims must be positive");
    MRCNN_REQUIRE(N >= 0 && crop_h > 0 && crop_w > 0, "mrcnn_crop_backward: bad N / crop size");
    if (int rc = check_layout(image_layout, "image_layout")) return rc;
    if (int rc = check_layout(grads_layout, "grads_layout")) return rc;
    MRCNN_REQUIRE_DEV(grads_image);
    if (zero_fill)
        MRCNN_CUDA(cudaMemsetAsync(grads_image, 0, sizeof(float) * (size_t)B * C * H * W, (cudaStream_t)stream));
    if (N == 0) return MRCNN_OK;
    MRCNN_REQUIRE_DEV(grads);
    MRCNN_REQUIRE_DEV(boxes);
    MRCNN_REQUIRE_DEV(box_index);
    RoiParams p = {};
    p.lv[0] = {grads_image, H, W};
    p.pyramid = 0;
    p.B = B; p.C = C;
    p.boxes = boxes; p.box_index = box_index; p.N = N;
    p.ph = crop_h; p.pw = crop_w; p.extrap = 0.f;
    p.crops = const_cast<float*>(grads);
    p.err = device_error_word();
    MRCNN_REQUIRE(p.err != nullptr, "cannot allocate device error word");
    return launch_roi(p, image_layout, grads_layout, true, (cudaStream_t)stream);
}

int mrcnn_pyramid_roi_align_forward(const float* const fm[4], const int H[4], const int W[4], int B, int C,
                                    int fm_layout, const float* boxes, const int32_t* box_index, int N, int pool,
                                    float image_area, float* out, int out_layout, int32_t* levels_out,
                                    mrcnn_stream_t stream) {
    MRCNN_REQUIRE(fm && H && W, "mrcnn_pyramid_roi_align_forward: null level tables");
    MRCNN_REQUIRE(B > 0 && C > 0 && N >= 0 && pool > 0 && image_area > 0.f, "mrcnn_pyramid_roi_align_forward: bad sizes");
    if (int rc = check_layout(fm_layout, "fm_layout")) return rc;
    if (int rc = check_layout(out_layout, "out_layout")) return rc;
    if (N == 0) return MRCNN_OK;
    RoiParams p = {};
    for (int l = 0; l < 4; ++l) {
        MRCNN_REQUIRE(H[l] > 0 && W[l] > 0, "mrcnn_pyramid_roi_align_forward: level %d has empty shape", l);
        MRCNN_REQUIRE_DEV(fm[l]);
        p.lv[l] = {const_cast<float*>(fm[l]), H[l], W[l]};
    }
    MRCNN_REQUIRE_DEV(boxes);
    MRCNN_REQUIRE_DEV(out);
    if (box_index) MRCNN_REQUIRE_DEV(box_index);
    if (levels_out) MRCNN_REQUIRE_DEV(levels_out);
    p.pyramid = 1;
    p.rule = make_level_rule(image_area);
    p.B = B; p.C = C;
    p.boxes = boxes; p.box_index = box_index; p.N = N;
    p.ph = pool; p.pw = pool; p.extrap = 0.f;  // model.py:373 CropFunction(pool, pool, 0)
    p.crops = out;
    p.levels_out = levels_out;
    p.err = device_error_word();
    MRCNN_REQUIRE(p.err != nullptr, "cannot allocate device error word");
    return launch_roi(p, fm_layout, out_layout, false, (cudaStream_t)stream);
}

int mrcnn_pyramid_roi_align_backward(const float* grads, int grads_layout, const int H[4], const int W[4], int B,
                                     int C, const float* boxes, const int32_t* box_index, int N, int pool,
                                     float image_area, float* const gfm[4], int gfm_layout, int zero_fill,
                                     const int32_t* image_offsets_host, int algo, void* workspace, size_t workspace_bytes,
                                     mrcnn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRCNN_REQUIRE(gfm && H && W, "mrcnn_pyramid_roi_align_backward: null level tables");
    MRCNN_REQUIRE(B > 0 && C > 0 && N >= 0 && pool > 0 && image_area > 0.f, "mrcnn_pyramid_roi_align_backward: bad sizes");
    if (int rc = check_layout(gfm_layout, "gfm_layout")) return rc;
    if (int rc = check_layout(grads_layout, "grads_layout")) return rc;
    RoiParams p = {};
    size_t per_image[4];
    for (int l = 0; l < 4; ++l) {
        MRCNN_REQUIRE(H[l] > 0 && W[l] > 0, "mrcnn_pyramid_roi_align_backward: level %d has empty shape", l);
        MRCNN_REQUIRE_DEV(gfm[l]);
        p.lv[l] = {gfm[l], H[l], W[l]};
        per_image[l] = (size_t)C * H[l] * W[l];
    }
    if (N > 0) {
        MRCNN_REQUIRE_DEV(grads);
        MRCNN_REQUIRE_DEV(boxes);
        if (box_index && !image_offsets_host) MRCNN_REQUIRE_DEV(box_index);
    }
    p.pyramid = 1;
    p.rule = make_level_rule(image_area);
    p.C = C;
    p.ph = pool; p.pw = pool; p.extrap = 0.f;
    p.err = device_error_word();
    MRCNN_REQUIRE(p.err != nullptr, "cannot allocate device error word");

    MRCNN_REQUIRE(algo == MRCNN_BWD_AUTO || algo == MRCNN_BWD_GATHER || algo == MRCNN_BWD_SCATTER,
                  "mrcnn_pyramid_roi_align_backward: unknown algo %d", algo);
    const bool can_gather = image_offsets_host == nullptr &&
                            gather_eligible(B, C, N, pool, gfm_layout, grads_layout, grads, gfm, workspace, workspace_bytes);
    if (algo == MRCNN_BWD_GATHER)
        MRCNN_REQUIRE(can_gather,
                      "mrcnn_pyramid_roi_align_backward: MRCNN_BWD_GATHER needs channels-last grads and gfm, C %% 4 == 0, "
                      "0 < N <= %d, no image_offsets_host and a 256-byte aligned workspace of "
                      "mrcnn_pyramid_roi_align_backward_workspace_bytes()", kBinMaxN);
    if (can_gather && algo == MRCNN_BWD_GATHER) {
        // tile-owner gather: writes every pixel once (zero fill included), no atomics
        return launch_bwd_gather(grads, H, W, B, C, boxes, box_index, N, pool, image_area, gfm, zero_fill ? 0 : 1, workspace,
                                 stream);
    }
    if (image_offsets_host == nullptr) {
        if (zero_fill) {
            size_t elems[4];
            for (int l = 0; l < 4; ++l) elems[l] = per_image[l] * (size_t)B;
            if (int rc = launch_zero(gfm, elems, 4, stream)) return rc;
        }
        if (N == 0) return MRCNN_OK;
        p.B = B;
        p.boxes = boxes; p.box_index = box_index; p.N = N;
        p.crops = const_cast<float*>(grads);
        return launch_roi(p, gfm_layout, grads_layout, true, stream);
    }
    // image-by-image: clear one image's pyramid slice, then scatter that image's boxes into it
    MRCNN_REQUIRE(image_offsets_host[0] == 0 && image_offsets_host[B] == N,
                  "mrcnn_pyramid_roi_align_backward: image_offsets_host must run from 0 to N");
    const size_t crop_elems = (size_t)C * pool * pool;
    for (int i = 0; i < B; ++i) {
        const int beg = image_offsets_host[i], end = image_offsets_host[i + 1];
        MRCNN_REQUIRE(beg <= end, "mrcnn_pyramid_roi_align_backward: image_offsets_host must be non-decreasing");
        float* slice[4];
        for (int l = 0; l < 4; ++l) {
            slice[l] = gfm[l] + (size_t)i * per_image[l];
            p.lv[l].ptr = slice[l];
        }
        if (zero_fill)
            if (int rc = launch_zero(slice, per_image, 4, stream)) return rc;
        if (end == beg) continue;
        p.B = 1;
        p.boxes = boxes + (size_t)4 * beg;
        p.box_index = nullptr;
        p.N = end - beg;
        p.crops = const_cast<float*>(grads) + (size_t)beg * crop_elems;
        if (int rc = launch_roi(p, gfm_layout, grads_layout, true, stream)) return rc;
    }
    return MRCNN_OK;
}

}  // extern "C"
