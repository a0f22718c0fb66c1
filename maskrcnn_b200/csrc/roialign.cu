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
#include <stdlib.h>
#include <string.h>

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
    float* crops_b;  // fused two-head forward: the second head's output (pool 7), else unused
    float negzero;  // -0.0f, opaque to ptxas (see bilerp4)
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
                    v = bilerp4(tl[u], tr[u], bl[u], br[u], xl[u], yl[u], p.negzero);
                } else {
                    v = make_float4(p.extrap, p.extrap, p.extrap, p.extrap);
                }
                if (kOutNHWC) {
                    stg_f4_stream(out_nhwc + b * C, v);
                } else if (POOL == 7 || POOL == 14) {
                    // the tile mirrors the output ([channel][bin], the four channels of a lane adjacent): it leaves with bulk
                    // copies.  Rows are P2 floats apart; at 14x14 (P2 = 196 = 4 mod 32) every group of four rows is skewed by
                    // four more floats, which keeps the stores at the 2-way bank conflict an odd pitch would give
                    float* t = tile + (4 * lane) * P2 + (POOL == 14 ? 4 * lane : 0) + b;
                    t[0] = v.x;
                    t[P2] = v.y;
                    t[2 * P2] = v.z;
                    t[3 * P2] = v.w;
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
        const int cc = min(kChunk, C - c0);
        float* dst = p.crops + ((size_t)n * C + c0) * P2;   // contiguous [cc][P2] block of the NCHW output
        if (POOL == 7 || POOL == 14) {
            // TMA epilogue: no second pass through the load/store pipe.  7x7: the whole [cc][49] block is one copy; 14x14: one copy
            // per group of four channel rows (3136 bytes each, 16-byte aligned on both sides)
            fence_proxy_async();
            __syncthreads();
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            if (POOL == 7) {
                if (tid == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(tile)),
                                 "r"((uint32_t)(cc * P2 * 4)), "l"(policy)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            } else if (tid < cc / 4) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + (size_t)tid * 4 * P2),
                             "r"(smem_u32(tile + tid * (4 * P2 + 4))), "r"((uint32_t)(4 * P2 * 4)), "l"(policy)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else {
            __syncthreads();
            tile_copy<false>(tile, dst, cc * P2, P2, P2pad);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Forward, channels-last input AND output, compile-time pool (7, 14): column-stationary threads.
// Same CTA (RoI x 64 channels) and the same arithmetic, but a (slot) owns ONE bin column and walks its rows: the column
// tap lives in registers for the whole CTA, a bin costs one shared-memory load (the row tap) instead of four, and the
// addresses are two fixed column pointers plus the row offset.  ncu on the generic kernel showed the LSU data pipe at
// 67 % with as many shared-memory wavefronts (tap loads) as global ones - this variant removes three quarters of them.
// Pool 14: 14 of the 16 slots own a column; pool 7: 2 row phases x 7 columns.
// ------------------------------------------------------------------------------------------------
// The CTA covers 4 * LANES channels with SLOTS * LANES threads (small CTAs: more of them resident, so the
// box -> taps -> barrier prologue of one overlaps the streaming of the others).
template <int POOL, int LANES, int SLOTS>
__global__ void __launch_bounds__(SLOTS * LANES, 1024 / (SLOTS * LANES)) roialign_fwd_nhwc_col_kernel(const RoiParams p) {
    __shared__ TapS s_ty[kMaxPool];
    __shared__ TapS s_tx[kMaxPool];
    constexpr int kPhases = SLOTS / POOL;  // rows handled in parallel by different slots
    static_assert(kPhases >= 1, "pool must not exceed the slot count");

    // the channel chunks of one RoI are adjacent in launch order: their 256-byte pieces of the same 1 KB output rows
    // and input pixels are in flight together (merged in L2, same DRAM pages)
    constexpr int kCh = 4 * LANES;
    const int chunks = (p.C + kCh - 1) / kCh;
    const int n = blockIdx.x / chunks;
    const int c0 = (blockIdx.x - n * chunks) * kCh;
    const int tid = threadIdx.x;
    const int C = p.C;

    float4 box;
    const RoiCtx ctx = select_level(p, n, box);
    stage_taps(p, ctx, box, POOL, POOL, s_ty, s_tx);
    __syncthreads();

    const int lane = tid % LANES;
    const int slot = tid / LANES;
    const int c = c0 + 4 * lane;
    if (c >= C || slot >= POOL * kPhases) return;  // C % 4 == 0 is guaranteed by the launcher
    const int phase = slot / POOL;
    const int x = slot - phase * POOL;
    const TapS tx = s_tx[x];
    const float* src_lo = ctx.base + c + (unsigned)tx.lo;  // tap offsets are non-negative element offsets
    const float* src_hi = ctx.base + c + (unsigned)tx.hi;
    float* out = p.crops + ((size_t)n * (POOL * POOL) + x) * C + c;
    const float4 ext = make_float4(p.extrap, p.extrap, p.extrap, p.extrap);

#pragma unroll 1
    for (int y0 = phase; y0 < POOL; y0 += 2 * kPhases) {
        const int y1 = y0 + kPhases;
        const bool has1 = y1 < POOL;
        const TapS ta = s_ty[y0];
        const TapS tb = s_ty[has1 ? y1 : y0];
        const bool in_a = ta.valid && tx.valid;
        const bool in_b = has1 && tb.valid && tx.valid;
        float4 tl0, tr0, bl0, br0, tl1, tr1, bl1, br1;
        if (in_a) {
            tl0 = ldg_f4(src_lo + (unsigned)ta.lo);
            tr0 = ldg_f4(src_hi + (unsigned)ta.lo);
            bl0 = ldg_f4(src_lo + (unsigned)ta.hi);
            br0 = ldg_f4(src_hi + (unsigned)ta.hi);
        }
        if (in_b) {
            tl1 = ldg_f4(src_lo + (unsigned)tb.lo);
            tr1 = ldg_f4(src_hi + (unsigned)tb.lo);
            bl1 = ldg_f4(src_lo + (unsigned)tb.hi);
            br1 = ldg_f4(src_hi + (unsigned)tb.hi);
        }
        const float4 va = in_a ? bilerp4(tl0, tr0, bl0, br0, tx.lerp, ta.lerp, p.negzero) : ext;
        stg_f4_stream(out + (unsigned)(y0 * POOL * C), va);
        if (has1) {
            const float4 vb = in_b ? bilerp4(tl1, tr1, bl1, br1, tx.lerp, tb.lerp, p.negzero) : ext;
            stg_f4_stream(out + (unsigned)(y1 * POOL * C), vb);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Forward of BOTH heads in one launch (training and inference pool the same RoIs twice from the same pyramid: 7x7 for the box
// head, model.py:778, and 14x14 for the mask head, model.py:889).  Same CTA (RoI x 64 channels) and the same column-stationary
// threads as above, first the 14x14 bins, then the 7x7 bins of the same RoI: the 7x7 taps lie inside the footprint the 14x14
// pass has just pulled through L1 / L2, so the second head costs its output bytes and (almost) no DRAM reads - 0.58 GB of the
// 8.29 GB a configs[3] step moves.  The arithmetic per bin is the single-head kernel's, so both outputs are bit-identical to it.
// ------------------------------------------------------------------------------------------------
template <int POOL, int SLOTS>
__device__ __forceinline__ void col_pass(const RoiParams& p, const RoiCtx& ctx, const TapS* s_ty, const TapS* s_tx, int n, int c, float* crops) {
    constexpr int kPhases = SLOTS / POOL;
    const int slot = threadIdx.x / kLanes;
    if (slot >= POOL * kPhases) return;
    const int C = p.C;
    const int phase = slot / POOL;
    const int x = slot - phase * POOL;
    const TapS tx = s_tx[x];
    const float* src_lo = ctx.base + c + (unsigned)tx.lo;
    const float* src_hi = ctx.base + c + (unsigned)tx.hi;
    float* out = crops + ((size_t)n * (POOL * POOL) + x) * C + c;
    const float4 ext = make_float4(p.extrap, p.extrap, p.extrap, p.extrap);
#pragma unroll 1
    for (int y0 = phase; y0 < POOL; y0 += 2 * kPhases) {
        const int y1 = y0 + kPhases;
        const bool has1 = y1 < POOL;
        const TapS ta = s_ty[y0];
        const TapS tb = s_ty[has1 ? y1 : y0];
        const bool in_a = ta.valid && tx.valid;
        const bool in_b = has1 && tb.valid && tx.valid;
        float4 tl0, tr0, bl0, br0, tl1, tr1, bl1, br1;
        if (in_a) {
            tl0 = ldg_f4(src_lo + (unsigned)ta.lo);
            tr0 = ldg_f4(src_hi + (unsigned)ta.lo);
            bl0 = ldg_f4(src_lo + (unsigned)ta.hi);
            br0 = ldg_f4(src_hi + (unsigned)ta.hi);
        }
        if (in_b) {
            tl1 = ldg_f4(src_lo + (unsigned)tb.lo);
            tr1 = ldg_f4(src_hi + (unsigned)tb.lo);
            bl1 = ldg_f4(src_lo + (unsigned)tb.hi);
            br1 = ldg_f4(src_hi + (unsigned)tb.hi);
        }
        const float4 va = in_a ? bilerp4(tl0, tr0, bl0, br0, tx.lerp, ta.lerp, p.negzero) : ext;
        stg_f4_stream(out + (unsigned)(y0 * POOL * C), va);
        if (has1) {
            const float4 vb = in_b ? bilerp4(tl1, tr1, bl1, br1, tx.lerp, tb.lerp, p.negzero) : ext;
            stg_f4_stream(out + (unsigned)(y1 * POOL * C), vb);
        }
    }
}

// Channels-last pyramid -> NCHW crops with the column-stationary threads of the kernels above: the blended float4 of a bin goes
// into a shared-memory tile laid out like the output block ([channel][bin]; at 14x14 every group of four rows skewed by 16 bytes,
// see roialign_fwd_nhwc_kernel) and the tile leaves with bulk copies.  Replaces the slot-strided loop of the general kernel for
// the two head sizes: fewer shared-memory tap loads per bin, the same epilogue.
// LANES = 16: 64 channels per CTA (256 threads, 50 KB tile at 14x14: 4 CTAs per SM); LANES = 8: 32 channels per CTA (128 threads, 25 KB:
// 8 per SM) - for small launches (an inference call: 1000 RoIs), where more and shorter CTAs overlap their phases better.
template <int POOL, int LANES>
__global__ void __launch_bounds__(16 * LANES) roialign_fwd_nhwc_colt_kernel(const RoiParams p) {
    constexpr int P2 = POOL * POOL;
    constexpr int kPhases = kSlots / POOL;
    constexpr int kCh = 4 * LANES;
    extern __shared__ __align__(16) float tile[];
    __shared__ TapS s_ty[kMaxPool];
    __shared__ TapS s_tx[kMaxPool];
    const int chunks = (p.C + kCh - 1) / kCh;
    const int n = blockIdx.x / chunks;
    const int c0 = (blockIdx.x - n * chunks) * kCh;
    const int tid = threadIdx.x;
    const int C = p.C;
    float4 box;
    const RoiCtx ctx = select_level(p, n, box);
    stage_taps(p, ctx, box, POOL, POOL, s_ty, s_tx);
    __syncthreads();
    const int lane = tid % LANES, slot = tid / LANES;
    const int c = c0 + 4 * lane;
    if (c < C && slot < POOL * kPhases) {
        const int phase = slot / POOL;
        const int x = slot - phase * POOL;
        const TapS tx = s_tx[x];
        const float* src_lo = ctx.base + c + (unsigned)tx.lo;
        const float* src_hi = ctx.base + c + (unsigned)tx.hi;
        float* t = tile + (4 * lane) * P2 + (POOL == 14 ? 4 * lane : 0) + x;
        const float4 ext = make_float4(p.extrap, p.extrap, p.extrap, p.extrap);
#pragma unroll 1
        for (int y0 = phase; y0 < POOL; y0 += 2 * kPhases) {
            const int y1 = y0 + kPhases;
            const bool has1 = y1 < POOL;
            const TapS ta = s_ty[y0];
            const TapS tb = s_ty[has1 ? y1 : y0];
            const bool in_a = ta.valid && tx.valid;
            const bool in_b = has1 && tb.valid && tx.valid;
            float4 tl0, tr0, bl0, br0, tl1, tr1, bl1, br1;
            if (in_a) {
                tl0 = ldg_f4(src_lo + (unsigned)ta.lo);
                tr0 = ldg_f4(src_hi + (unsigned)ta.lo);
                bl0 = ldg_f4(src_lo + (unsigned)ta.hi);
                br0 = ldg_f4(src_hi + (unsigned)ta.hi);
            }
            if (in_b) {
                tl1 = ldg_f4(src_lo + (unsigned)tb.lo);
                tr1 = ldg_f4(src_hi + (unsigned)tb.lo);
                bl1 = ldg_f4(src_lo + (unsigned)tb.hi);
                br1 = ldg_f4(src_hi + (unsigned)tb.hi);
            }
            const float4 va = in_a ? bilerp4(tl0, tr0, bl0, br0, tx.lerp, ta.lerp, p.negzero) : ext;
            float* o = t + y0 * POOL;
            o[0] = va.x; o[P2] = va.y; o[2 * P2] = va.z; o[3 * P2] = va.w;
            if (has1) {
                const float4 vb = in_b ? bilerp4(tl1, tr1, bl1, br1, tx.lerp, tb.lerp, p.negzero) : ext;
                float* o1 = t + y1 * POOL;
                o1[0] = vb.x; o1[P2] = vb.y; o1[2 * P2] = vb.z; o1[3 * P2] = vb.w;
            }
        }
    }
    const int cc = min(kCh, C - c0);
    float* dst = p.crops + ((size_t)n * C + c0) * P2;
    fence_proxy_async();
    __syncthreads();
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (POOL == 7) {
        if (tid == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(tile)),
                         "r"((uint32_t)(cc * P2 * 4)), "l"(policy)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else if (tid < cc / 4) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + (size_t)tid * 4 * P2),
                     "r"(smem_u32(tile + tid * (4 * P2 + 4))), "r"((uint32_t)(4 * P2 * 4)), "l"(policy)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

__global__ void __launch_bounds__(kThreads, 4) roialign_fwd_nhwc_pair_kernel(const RoiParams p) {
    __shared__ TapS s_ty14[14], s_tx14[14], s_ty7[8], s_tx7[8];
    const int chunks = (p.C + kChunk - 1) / kChunk;
    const int n = blockIdx.x / chunks;
    const int c0 = (blockIdx.x - n * chunks) * kChunk;
    float4 box;
    const RoiCtx ctx = select_level(p, n, box);
    stage_taps(p, ctx, box, 14, 14, s_ty14, s_tx14);
    // the 7x7 taps: threads [32, 39) rows, [96, 103) columns (stage_taps uses [0, ph) and [64, 64 + pw))
    {
        const int tid = threadIdx.x;
        if (tid >= 32 && tid < 39) {
            const AxisTap t = axis_tap(box.x, box.z, ctx.H, 7, tid - 32);
            TapS o;
            o.valid = (t.lo >= 0) && ctx.ok;
            o.lo = o.valid ? t.lo * ctx.W * p.C : 0;
            o.hi = o.valid ? t.hi * ctx.W * p.C : 0;
            o.lerp = t.lerp;
            s_ty7[tid - 32] = o;
        } else if (tid >= 96 && tid < 103) {
            const AxisTap t = axis_tap(box.y, box.w, ctx.W, 7, tid - 96);
            TapS o;
            o.valid = (t.lo >= 0) && ctx.ok;
            o.lo = o.valid ? t.lo * p.C : 0;
            o.hi = o.valid ? t.hi * p.C : 0;
            o.lerp = t.lerp;
            s_tx7[tid - 96] = o;
        }
    }
    __syncthreads();
    const int c = c0 + 4 * (threadIdx.x % kLanes);
    if (c >= p.C) return;  // C % 4 == 0 is guaranteed by the launcher
    col_pass<14, kSlots>(p, ctx, s_ty14, s_tx14, n, c, p.crops);
    col_pass<7, kSlots>(p, ctx, s_ty7, s_tx7, n, c, p.crops_b);
}

// ------------------------------------------------------------------------------------------------
// Forward, channels-last in and out, row-walking (the mask head's 14x14): a thread owns ONE bin column and K x 4 channels
// and walks the bin rows top to bottom.  The blend is separable exactly as crop_cpu.cpp:107-110 writes it - top = H(y_lo),
// bot = H(y_hi), H(r) = v[r][x_lo] + (v[r][x_hi] - v[r][x_lo]) * x_lerp, the same three roundings per value - so H of a feature
// row is computed once and kept in registers while the walk stays on that row: an up-sampled RoI (14 bin rows over ~10-20
// feature rows) touches ~p + 1 rows instead of 2p, 2 x K loads per NEW row instead of 4 per bin (-40 % L1/TEX wavefronts;
// the column-stationary kernel above was L1/TEX- and latency-bound at 76 % of HBM with DRAM traffic already at the algorithmic
// minimum).  Row decisions are warp-uniform (one RoI per CTA).  K float4 per thread keep 2K..4K loads in flight.
// ------------------------------------------------------------------------------------------------
template <int POOL, int K>
__global__ void __launch_bounds__(256) roialign_fwd_nhwc_row_kernel(const RoiParams p) {
    __shared__ TapS s_ty[kMaxPool];
    __shared__ TapS s_tx[kMaxPool];
    static_assert(POOL <= 16, "one slot per bin column");
    constexpr int kCh = 64 * K;  // channels per CTA
    const int chunks = (p.C + kCh - 1) / kCh;
    const int n = blockIdx.x / chunks;
    const int c0 = (blockIdx.x - n * chunks) * kCh;
    const int tid = threadIdx.x;
    const int C = p.C;

    float4 box;
    const RoiCtx ctx = select_level(p, n, box);
    stage_taps(p, ctx, box, POOL, POOL, s_ty, s_tx);
    __syncthreads();

    const int lane = tid & 15, x = tid >> 4;
    if (x >= POOL) return;
    const TapS tx = s_tx[x];
    const float4 ext = make_float4(p.extrap, p.extrap, p.extrap, p.extrap);
    bool live[K];
    const float* slo[K];
    const float* shi[K];
    float* out[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int c = c0 + 64 * k + 4 * lane;
        live[k] = c < C;                                  // C % 4 == 0 is guaranteed by the launcher
        const int cs = live[k] ? c : 0;
        slo[k] = ctx.base + cs + (unsigned)tx.lo;
        shi[k] = ctx.base + cs + (unsigned)tx.hi;
        out[k] = p.crops + ((size_t)n * (POOL * POOL) + x) * C + cs;
    }
    const unsigned long long nz = pack2(p.negzero, p.negzero), x2 = pack2(tx.lerp, tx.lerp);

    // horizontal blend of one feature row (element offset `row`) for the thread's K x 4 channels
    auto hblend = [&](unsigned row, float4 (&H)[K]) {
        float4 a[K], b[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (live[k]) {
                a[k] = ldg_f4(slo[k] + row);
                b[k] = ldg_f4(shi[k] + row);
            } else {
                a[k] = b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            unpack2(lerp2(pack2(a[k].x, a[k].y), pack2(b[k].x, b[k].y), x2, nz), H[k].x, H[k].y);
            unpack2(lerp2(pack2(a[k].z, a[k].w), pack2(b[k].z, b[k].w), x2, nz), H[k].z, H[k].w);
        }
    };

    int ra = -1, rb = -1;
    float4 Ha[K], Hb[K];
#pragma unroll
    for (int k = 0; k < K; ++k) Ha[k] = Hb[k] = make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll 1
    for (int y = 0; y < POOL; ++y) {
        const TapS ty = s_ty[y];  // warp-uniform
        const bool in = ty.valid && tx.valid;
        if (ty.valid && tx.valid) {
            if (ty.lo == rb) {
#pragma unroll
                for (int k = 0; k < K; ++k) Ha[k] = Hb[k];
                ra = rb;
                rb = -1;
            } else if (ty.lo != ra) {
                hblend((unsigned)ty.lo, Ha);
                ra = ty.lo;
            }
            if (ty.hi == ra) {
#pragma unroll
                for (int k = 0; k < K; ++k) Hb[k] = Ha[k];
                rb = ra;
            } else if (ty.hi != rb) {
                hblend((unsigned)ty.hi, Hb);
                rb = ty.hi;
            }
        }
        const unsigned long long y2 = pack2(ty.lerp, ty.lerp);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float4 v = ext;
            if (in) {
                unpack2(lerp2(pack2(Ha[k].x, Ha[k].y), pack2(Hb[k].x, Hb[k].y), y2, nz), v.x, v.y);
                unpack2(lerp2(pack2(Ha[k].z, Ha[k].w), pack2(Hb[k].z, Hb[k].w), y2, nz), v.z, v.w);
            }
            if (live[k]) stg_f4_stream(out[k] + (unsigned)(y * POOL * C), v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Forward, channels-last in and out, TMA-pipelined (the mask head's 14x14 at C * 4 bytes per pixel a multiple of 16).
//
// In channels-last memory the footprint of a RoI on ONE feature row - pixels x_min .. x_max, all C channels - is one
// contiguous run of ncols * C * 4 bytes (15 KB for a typical 15-column footprint at C = 256), and one output row of the crop
// (p bins x C channels) is one contiguous run of p * C * 4 bytes (14 KB).  So the whole data movement of a RoI is a few dozen
// large 1-D bulk copies, and the SM's load/store pipe only ever touches shared memory:
//
//   producer  one thread.  Walks the RoI's feature rows in the order the blend needs them (the row list is built once per CTA
//             by the same walk the consumers do) and issues one cp.async.bulk global -> shared per row into a ring of slots,
//             each tracked by a "full" mbarrier (expect_tx = row bytes); it re-uses a slot when the consumers' "empty"
//             mbarrier of that slot completes.  The ring is 64 KB: 2 .. 8 rows in flight per CTA, two CTAs per SM.
//   consumers 256 threads, thread = (bin column, 4-channel lane), C / 64 channel quarters each.  Wait for the row, read their two
//             taps (x_lo, x_hi) from the slot, blend horizontally - H(r) exactly as crop_cpu.cpp:107-108 - release the slot,
//             and keep H of the current and previous row in registers (the separable form of the row-walking kernel above,
//             bit-identical).  Every finished output row is staged in shared memory in the output's own order and leaves as
//             ONE bulk copy shared -> global with an L2 evict-first hint (three staging rows: the store of row y is still
//             reading while row y + 1 is blended).
//
// DRAM sees long sequential bursts in both directions instead of 256-byte pieces from four CTAs per RoI; nothing waits on a
// register load.  RoIs whose footprint is wider than the ring allows (ncols * C * 4 * 2 > ring) take the in-kernel fallback:
// the column-stationary register path, chunk by chunk.
// ------------------------------------------------------------------------------------------------
constexpr int kTmaRingBytes = 64 * 1024;
constexpr int kTmaMaxSlots = 8;
constexpr int kTmaOutBufs = 3;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int POOL>
__global__ void __launch_bounds__(288) roialign_fwd_nhwc_tma_kernel(const RoiParams p) {
    constexpr int kMaxQ = 4;  // channel quarters of 64 per thread: C <= 256
    extern __shared__ __align__(128) unsigned char s_raw[];
    float* ring = reinterpret_cast<float*>(s_raw);                                    // [slots][ncols][C]
    float* outb = reinterpret_cast<float*>(s_raw + kTmaRingBytes);                    // [kTmaOutBufs][POOL][C]
    __shared__ TapS s_ty[kMaxPool];
    __shared__ TapS s_tx[kMaxPool];
    __shared__ int s_rows[2 * POOL];     // element offsets (row * W * C) of the feature rows, in consumption order
    __shared__ int s_nrows, s_xmin, s_ncols;
    __shared__ __align__(8) uint64_t s_full[kTmaMaxSlots], s_empty[kTmaMaxSlots];

    const int n = blockIdx.x;
    const int tid = threadIdx.x;
    const int C = p.C;

    float4 box;
    const RoiCtx ctx = select_level(p, n, box);
    stage_taps(p, ctx, box, POOL, POOL, s_ty, s_tx);
    __syncthreads();
    if (tid == 0) {
        int xmin = INT_MAX, xmax = -1;
        for (int x = 0; x < POOL; ++x)
            if (s_tx[x].valid) {
                xmin = min(xmin, s_tx[x].lo);
                xmax = max(xmax, s_tx[x].hi);
            }
        int nr = 0;
        if (xmax >= 0) {  // the walk of the consumers, rows only: which feature row is fetched when
            int ra = -1, rb = -1;
            for (int y = 0; y < POOL; ++y) {
                const TapS ty = s_ty[y];
                if (!ty.valid) continue;
                if (ty.lo == rb) { ra = rb; rb = -1; }
                else if (ty.lo != ra) { s_rows[nr++] = ty.lo; ra = ty.lo; }
                if (ty.hi == ra) rb = ra;
                else if (ty.hi != rb) { s_rows[nr++] = ty.hi; rb = ty.hi; }
            }
        }
        MRCNN_DBG(nr >= 0 && nr <= 2 * POOL);
        s_nrows = nr;
        s_xmin = (xmax >= 0) ? xmin : 0;          // element offset (x * C)
        s_ncols = (xmax >= 0) ? (xmax - xmin) / C + 1 : 0;
        for (int i = 0; i < kTmaMaxSlots; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 8);            // one arrival per consumer warp
        }
        fence_barrier_init();
    }
    __syncthreads();
    const int nrows = s_nrows, ncols = s_ncols, xmin = s_xmin;
    const uint32_t row_bytes = (uint32_t)ncols * (uint32_t)C * 4u;
    int slots = row_bytes ? (int)(kTmaRingBytes / row_bytes) : kTmaMaxSlots;
    if (slots > kTmaMaxSlots) slots = kTmaMaxSlots;
    const bool piped = slots >= 2;                // else: footprint too wide for the ring -> register path below
    const int P2 = POOL * POOL;
    float* crop = p.crops + (size_t)n * P2 * C;

    if (!piped) {  // ---- fallback: column-stationary register path over all channels (rare: very wide boxes)
        if (tid >= 256) return;
        const int lane = tid & 15, slot = tid >> 4;
        const float4 ext = make_float4(p.extrap, p.extrap, p.extrap, p.extrap);
        for (int c = 4 * lane; c < C; c += 64) {
            for (int b = slot; b < P2; b += 16) {
                const int y = b / POOL, x = b - y * POOL;
                const TapS ty = s_ty[y], tx = s_tx[x];
                float4 v = ext;
                if (ty.valid && tx.valid) {
                    const float* src = ctx.base + c;
                    v = bilerp4(ldg_f4(src + (ty.lo + tx.lo)), ldg_f4(src + (ty.lo + tx.hi)), ldg_f4(src + (ty.hi + tx.lo)),
                                ldg_f4(src + (ty.hi + tx.hi)), tx.lerp, ty.lerp, p.negzero);
                }
                stg_f4_stream(crop + (size_t)b * C + c, v);
            }
        }
        return;
    }

    if (tid >= 256) {  // ---- producer
        if (tid == 256) {
            const float* src = ctx.base + xmin;
            for (int i = 0; i < nrows; ++i) {
                const int sl = i % slots;
                const int k = i / slots;
                if (k > 0) mbar_wait(&s_empty[sl], (uint32_t)((k - 1) & 1));
                MRCNN_DBG(sl >= 0 && sl < slots && (size_t)(sl + 1) * row_bytes <= (size_t)kTmaRingBytes && (row_bytes & 15u) == 0);
                MRCNN_DBG((unsigned)s_rows[i] + (unsigned)xmin + (unsigned)ncols * (unsigned)C <= (unsigned)(ctx.H * ctx.W * C));
                mbar_expect_tx(&s_full[sl], row_bytes);
                bulk_g2s(reinterpret_cast<unsigned char*>(ring) + (size_t)sl * row_bytes, src + (unsigned)s_rows[i], row_bytes, &s_full[sl]);
            }
        }
        return;
    }

    // ---- consumers
    const int lane = tid & 15, x = tid >> 4, wlane = tid & 31;
    const bool colv = x < POOL;
    const TapS tx = s_tx[colv ? x : 0];
    const bool x_in = colv && tx.valid;
    const int nq = (C + 63) / 64;                                     // <= kMaxQ by the launcher
    const int jlo = x_in ? (tx.lo - xmin) : 0, jhi = x_in ? (tx.hi - xmin) : 0;   // element offsets inside a ring row
    const unsigned long long nz = pack2(p.negzero, p.negzero), x2 = pack2(tx.lerp, tx.lerp);
    const float4 ext = make_float4(p.extrap, p.extrap, p.extrap, p.extrap);
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    int next = 0;  // next entry of s_rows to consume
    auto take_row = [&](float4 (&H)[kMaxQ]) {
        const int sl = next % slots;
        MRCNN_DBG(next < nrows && jlo >= 0 && jhi >= jlo && (unsigned)(jhi + C) * 4u <= row_bytes);
        mbar_wait(&s_full[sl], (uint32_t)((next / slots) & 1));
        const float* row = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(ring) + (size_t)sl * row_bytes);
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) {
            const int c = 64 * q + 4 * lane;
            if (q < nq && c < C && x_in) {
                const float4 a = *reinterpret_cast<const float4*>(row + jlo + c);
                const float4 b = *reinterpret_cast<const float4*>(row + jhi + c);
                unpack2(lerp2(pack2(a.x, a.y), pack2(b.x, b.y), x2, nz), H[q].x, H[q].y);
                unpack2(lerp2(pack2(a.z, a.w), pack2(b.z, b.w), x2, nz), H[q].z, H[q].w);
            }
        }
        __syncwarp();
        if (wlane == 0) mbar_arrive(&s_empty[sl]);
        ++next;
    };

    int ra = -1, rb = -1;
    float4 Ha[kMaxQ], Hb[kMaxQ];
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q) Ha[q] = Hb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool any_col = ncols > 0;

#pragma unroll 1
    for (int y = 0; y < POOL; ++y) {
        const TapS ty = s_ty[y];  // CTA-uniform
        if (ty.valid && any_col) {
            if (ty.lo == rb) {
#pragma unroll
                for (int q = 0; q < kMaxQ; ++q) Ha[q] = Hb[q];
                ra = rb;
                rb = -1;
            } else if (ty.lo != ra) {
                take_row(Ha);
                ra = ty.lo;
            }
            if (ty.hi == ra) {
#pragma unroll
                for (int q = 0; q < kMaxQ; ++q) Hb[q] = Ha[q];
                rb = ra;
            } else if (ty.hi != rb) {
                take_row(Hb);
                rb = ty.hi;
            }
        }
        float* ob = outb + (size_t)(y % kTmaOutBufs) * POOL * C;
        if (colv) {
            const bool in = ty.valid && x_in;
            const unsigned long long y2 = pack2(ty.lerp, ty.lerp);
#pragma unroll
            for (int q = 0; q < kMaxQ; ++q) {
                const int c = 64 * q + 4 * lane;
                if (q < nq && c < C) {
                    float4 v = ext;
                    if (in) {
                        unpack2(lerp2(pack2(Ha[q].x, Ha[q].y), pack2(Hb[q].x, Hb[q].y), y2, nz), v.x, v.y);
                        unpack2(lerp2(pack2(Ha[q].z, Ha[q].w), pack2(Hb[q].z, Hb[q].w), y2, nz), v.z, v.w);
                    }
                    *reinterpret_cast<float4*>(ob + x * C + c) = v;
                }
            }
        }
        fence_proxy_async();                             // this thread's row piece -> visible to the async proxy (the bulk store)
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the row is complete in shared memory
        if (tid == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(crop + (size_t)y * POOL * C),
                         "r"(smem_u32(ob)), "r"((uint32_t)(POOL * C * 4)), "l"(policy)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // at most one store still reading: the buffer written two rows from now (the one of row y - 1) is free by then
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
// Backward as a GATHER (row-owner) — channels-last gradients in, channels-last pyramid gradient out.
//
// The adjoint of bilinear sampling is a sparse-matrix product  dImage[pixels x C] = A * g[bins x C]  with four
// non-zeros per bin.  The scatter kernel walks A by columns (bins) and needs atomics, a zero-fill pass and a
// read-modify-write of the whole gradient pyramid.  This path walks A by ROWS: a "unit" is one row of 8
// consecutive pixels of one (image, level) gradient map, owned by exactly one warp, which keeps the 8 x 256
// channel sums in registers and writes every pixel exactly once - no atomics on gradient data, no zero-fill
// pass, no read-modify-write: DRAM traffic is the algorithmic minimum (upstream gradient read once, gradient
// pyramid written once).
//
//   pass 1  bwd_items_kernel<false>  every (RoI, bin row, floor|ceil tap) thread walks its bin columns and
//                                    counts the work items each unit will receive (run-aggregated atomics).
//   pass 2  bwd_alloc_kernel         a slice of the item array per unit (warp-aggregated cursor).
//   pass 3  bwd_items_kernel<true>   the same walk writes the items {gradient offset, wy(1-xl), wy xl, column}.
//   pass 4  roialign_bwd_gather_kernel  one warp per unit, coarse levels first: items are fetched 32 at a time
//                                    (one coalesced 512-byte read) and broadcast through shared memory; the
//                                    gradients stream through a 4-stage cp.async pipeline; a warp-uniform switch
//                                    on the column adds a bin into the one or two accumulators it touches.
// Passes 1-3 move ~16 bytes per non-zero of A (a few % of the gradient bytes) and replace all per-tile searching.
// ------------------------------------------------------------------------------------------------
constexpr int kGTile = 8;       // pixels per unit
constexpr int kONoop = 9;
constexpr int kOHead2 = 16;  // QItem.idx bit: the bin belongs to the second gradient tensor

struct __align__(16) QItem {
    int off;    // element offset of the bin in grads (32-bit by eligibility)
    float wa;   // weight of pixel idx - 1 of the unit
    float wb;   // weight of pixel idx
    int idx;    // floor column relative to the unit + 1, 0..8 (kONoop: no contribution)
};

struct UnitGeom {
    float* ptr;
    int H, W;
    int segs;       // units per pixel row: ceil(W / 8)
    int unit_base;  // first unit of the level; levels are laid out coarse to fine (P5 first)
};

struct GatherParams {
    UnitGeom g[4];
    LevelRule rule;
    int B, C, N;
    int ph, pw;
    int units;
    const float* boxes;
    const int32_t* box_index;
    const float* grads;   // [N][ph*pw][C]
    const float* grads2;  // second head of a fused two-head backward (items tagged kOHead2), else == grads
    int head_flag;        // 0 or kOHead2: ORed into the items the fill pass writes
    int out_nchw;         // 1: the gradient maps are NCHW (a unit = 8 consecutive pixels of every channel plane = one 32-byte sector each)
    int* cnt;             // [units] items per unit
    int* pos;            // [units] after alloc: first item; after fill: one past the last item
    unsigned int* cursor;
    QItem* items;
    int* err;
};

template <bool kFill>
__global__ void __launch_bounds__(256) bwd_items_kernel(const GatherParams p) {
    const int ph = p.ph, pw = p.pw;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)p.N * ph * 2) return;
    const int r = (int)(t & 1);
    const int by = (int)((t >> 1) % ph);
    const int n = (int)((t >> 1) / ph);
    const float y1 = __ldg(p.boxes + 4 * n), x1 = __ldg(p.boxes + 4 * n + 1);
    const float y2 = __ldg(p.boxes + 4 * n + 2), x2 = __ldg(p.boxes + 4 * n + 3);
    const int bi = p.box_index ? __ldg(p.box_index + n) : 0;
    if ((unsigned)bi >= (unsigned)p.B) {
        if (!kFill && by == 0 && r == 0) atomicOr(p.err, 1);
        return;
    }
    const int l = roi_level(y1, x1, y2, x2, p.rule) - 2;
    const UnitGeom G = (l == 0) ? p.g[0] : (l == 1) ? p.g[1] : (l == 2) ? p.g[2] : p.g[3];
    const AxisTap ty = axis_tap(y1, y2, G.H, ph, by);
    if (ty.lo < 0) return;
    if (r == 1 && ty.lerp == 0.0f) return;  // the ceil tap coincides with the floor tap (crop_cpu.cpp:254-260)
    const int y = ty.lo + r;
    const float wy = r ? ty.lerp : __fsub_rn(1.0f, ty.lerp);
    const int row_unit = G.unit_base + (bi * G.H + y) * G.segs;
    const int row_off = (n * ph * pw + by * pw) * p.C;

    // walk the bin columns; consecutive bins that fall into the same unit are claimed with one atomic
    int run_seg = -1, run_beg = 0, run_len = 0;
    for (int bx = 0; bx <= pw; ++bx) {
        AxisTap tx;
        tx.lo = -1; tx.hi = -1; tx.lerp = 0.f;
        if (bx < pw) tx = axis_tap(x1, x2, G.W, pw, bx);
        const int seg = (tx.lo >= 0) ? (tx.lo >> 3) : -1;
        if (seg != run_seg) {
            if (run_seg >= 0) {
                if (!kFill) {
                    atomicAdd(p.cnt + row_unit + run_seg, run_len);
                } else {
                    int q = atomicAdd(p.pos + row_unit + run_seg, run_len);
                    for (int b2 = run_beg; b2 < run_beg + run_len; ++b2) {
                        const AxisTap u = axis_tap(x1, x2, G.W, pw, b2);
                        const int j = u.lo & 7;
                        QItem it;
                        it.off = row_off + b2 * p.C;
                        it.wa = wy * __fsub_rn(1.0f, u.lerp);
                        it.wb = (j < 7) ? wy * u.lerp : 0.f;
                        it.idx = (j + 1) | p.head_flag;
                        p.items[q++] = it;
                    }
                }
            }
            run_seg = seg;
            run_beg = bx;
            run_len = 0;
        }
        if (seg >= 0) {
            ++run_len;
            // the ceil column of a bin whose floor column is the last of its unit belongs to the next unit
            if ((tx.lo & 7) == 7 && tx.lerp != 0.0f) {
                if (!kFill) {
                    atomicAdd(p.cnt + row_unit + seg + 1, 1);
                } else {
                    const int q = atomicAdd(p.pos + row_unit + seg + 1, 1);
                    QItem it;
                    it.off = row_off + bx * p.C;
                    it.wa = 0.f;
                    it.wb = wy * tx.lerp;
                    it.idx = p.head_flag;
                    p.items[q] = it;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) bwd_alloc_kernel(const GatherParams p) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (u < p.units) ? p.cnt[u] : 0;
    int incl = c;
    const int wl = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (wl >= o) incl += v;
    }
    unsigned int base = 0;
    if (wl == 31 && incl > 0) base = atomicAdd(p.cursor, (unsigned int)incl);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (u < p.units) p.pos[u] = (int)base + incl - c;
}

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

// Adds one bin into the unit's accumulators: pixel idx - 1 gets wa, pixel idx gets wb.
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

// One warp (= one CTA, so a finished unit frees its slot at once) per unit.  NV float4 per lane and pixel: a warp
// covers 128 NV channels per pass.  The gradients of U items per stage are copied straight into shared memory with
// cp.async (each lane copies and later reads back only its own 16-byte slots, so no cross-lane synchronisation is
// needed on the data) and ST stages are kept in flight per warp without holding them in registers.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// (No minimum-blocks bound: forcing 24 / 28 / 32 CTAs per SM - 80 / 72 / 64 registers - spills and measured 0.61 / 0.75 / 0.89 ms
// against 0.58 ms for the configs[3] 14x14 backward; U = 3 or ST = 6 are slower too.  profiles/r02_experiments.txt)
template <int NV, int U, int ST, bool kAccumulate, bool kOutNCHW = false>
__global__ void __launch_bounds__(32, 20) roialign_bwd_gather_kernel(const GatherParams p) {
    __shared__ QItem s_items[64];
    __shared__ __align__(16) float4 s_data[ST][U][NV][32];
    const int wl = threadIdx.x;
    // The default grid has one CTA per unit (one trip through this loop); a smaller, persistent grid strides over the units.
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        do {  // one unit

    const int l = (u < p.g[2].unit_base) ? 3 : (u < p.g[1].unit_base) ? 2 : (u < p.g[0].unit_base) ? 1 : 0;
    const UnitGeom G = (l == 0) ? p.g[0] : (l == 1) ? p.g[1] : (l == 2) ? p.g[2] : p.g[3];
    const int local = u - G.unit_base;
    const int rowid = local / G.segs;  // image * H + y
    const int seg = local - rowid * G.segs;
    const int C = p.C;
    const int npx = min(kGTile, G.W - seg * kGTile);
    float* out = G.ptr + ((size_t)rowid * G.W + seg * kGTile) * C;

    // NCHW gradient maps: channel ch of this unit = the 8 floats at ((img * C + ch) * H + y) * W + 8 * seg - one aligned sector
    const int img_ = rowid / G.H, y_ = rowid - img_ * G.H;
    float* out_nchw = G.ptr + (((size_t)img_ * C) * G.H + y_) * G.W + seg * kGTile;
    const size_t plane_ = (size_t)G.H * G.W;
    const bool vec8 = kOutNCHW && npx == kGTile && ((reinterpret_cast<uintptr_t>(out_nchw) | (plane_ * 4)) & 15u) == 0;
    const int n = __ldg(p.cnt + u);
    const int item_end = __ldg(p.pos + u);   // requested together with n: one round trip instead of two in front of the item loads
    if (n == 0) {  // nothing reaches this unit: it is all zeros
        if (!kAccumulate) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!kOutNCHW) {
                for (int i = 4 * wl; i < npx * C; i += 128) stg_f4_stream(out + i, z);
            } else {
                for (int ch = wl; ch < C; ch += 32) {
                    float* o = out_nchw + (size_t)ch * plane_;
                    if (vec8) {
                        stg_f4_stream(o, z);
                        stg_f4_stream(o + 4, z);
                    } else {
                        for (int j = 0; j < npx; ++j) o[j] = 0.f;
                    }
                }
            }
        }
        break;
    }
    MRCNN_DBG(n > 0 && item_end - n >= 0);
    const QItem* items = p.items + (item_end - n);
    for (int cbase = 0; cbase < C; cbase += 128 * NV) {
        const int c = cbase + 4 * wl;
        const bool live = c < C;
        AccRow<NV> acc;
#pragma unroll
        for (int j = 0; j < kGTile; ++j)
#pragma unroll
            for (int k = 0; k < NV; ++k) acc.a[j][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* gbase = p.grads + (live ? c : 0);
        const float* gbase2 = p.grads2 + (live ? c : 0);

        // The unit's items are ONE stream through a 64-entry ring in shared memory: while the groups of batch b (32 items) are
        // consumed, batch b + 1 is already staged and batch b + 2 sits in a register of every lane, so neither the item loads nor
        // a pipeline drain / refill interrupt the cp.async ring at a batch boundary (the first version staged 32 items at a time
        // and emptied the ring after each batch: two exposed round trips per 32 items for the units of the coarse levels, which
        // hold hundreds of items each).  Entries past n are no-ops (weight 0, offset 0).
        static_assert(32 % U == 0 && (ST - 1) * U <= 32, "a group never straddles a batch; the ring looks at most one batch ahead");
        auto load_batch = [&](int bch) -> QItem {
            QItem q;
            q.off = 0; q.wa = 0.f; q.wb = 0.f; q.idx = kONoop;
            const int i = 32 * bch + wl;
            if (i < n) {
                const int4 raw = __ldg(reinterpret_cast<const int4*>(items + i));
                q.off = raw.x; q.wa = __int_as_float(raw.y); q.wb = __int_as_float(raw.z); q.idx = raw.w;
            }
            return q;
        };
        auto issue = [&](int grp) {
            const int st = grp % ST;
#pragma unroll
            for (int e = 0; e < U; ++e) {
                const QItem it = s_items[(grp * U + e) & 63];
                const float* src = ((it.idx & kOHead2) ? gbase2 : gbase) + it.off;
#pragma unroll
                for (int k = 0; k < NV; ++k) cp_async16(&s_data[st][e][k][wl], src + 128 * k);
            }
        };
        {
            const QItem r0 = load_batch(0), r1 = load_batch(1);
            __syncwarp();   // every lane is done with the ring's previous contents
            s_items[wl] = r0;
            s_items[32 + wl] = r1;
        }
        QItem ahead = load_batch(2);
        __syncwarp();
        const int ngroups = (n + U - 1) / U;
        // prologue: ST - 1 stages in flight
#pragma unroll
        for (int sgi = 0; sgi < ST - 1; ++sgi) {
            if (sgi < ngroups) issue(sgi);
            cp_async_commit();
        }
#pragma unroll 1
        for (int gi = 0; gi < ngroups; ++gi) {
            if (gi > 0 && ((gi * U) & 31) == 0) {   // batch b = gi U / 32 starts: batch b - 1 is consumed, its half takes batch b + 1
                const int bch = (gi * U) >> 5;
                __syncwarp();
                s_items[((bch + 1) & 1) * 32 + wl] = ahead;
                ahead = load_batch(bch + 2);
                __syncwarp();
            }
            const int nxt = gi + ST - 1;
            if (nxt < ngroups) issue(nxt);
            cp_async_commit();
            cp_async_wait<ST - 1>();
            const int st = gi % ST;
#pragma unroll
            for (int e = 0; e < U; ++e) {
                const QItem it = s_items[(gi * U + e) & 63];
                float4 v[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k) v[k] = s_data[st][e][k][wl];
                owner_accumulate<NV>(acc, it.idx & 15, it.wa, it.wb, v);
            }
        }
        cp_async_wait<0>();

        if (kOutNCHW && live) {
            // every channel the lane holds: its 8 pixel sums are one sector of that channel's plane
#pragma unroll
            for (int k = 0; k < NV; ++k) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float* o = out_nchw + (size_t)(c + 128 * k + q) * plane_;
                    float v[kGTile];
#pragma unroll
                    for (int j = 0; j < kGTile; ++j) {
                        const float4 a = acc.a[j][k];
                        v[j] = (q == 0) ? a.x : (q == 1) ? a.y : (q == 2) ? a.z : a.w;
                    }
                    if (kAccumulate) {
#pragma unroll
                        for (int j = 0; j < kGTile; ++j)
                            if (j < npx) v[j] += o[j];
                    }
                    if (vec8) {
                        stg_f4_stream(o, make_float4(v[0], v[1], v[2], v[3]));
                        stg_f4_stream(o + 4, make_float4(v[4], v[5], v[6], v[7]));
                    } else {
#pragma unroll
                        for (int j = 0; j < kGTile; ++j)
                            if (j < npx) o[j] = v[j];
                    }
                }
            }
        } else if (live) {
            float* o = out + c;
#pragma unroll
            for (int j = 0; j < kGTile; ++j) {
                if (j < npx) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        float4 a = acc.a[j][k];
                        float* dst = o + j * C + 128 * k;
                        if (kAccumulate) {
                            const float4 o0 = *reinterpret_cast<const float4*>(dst);
                            a.x += o0.x; a.y += o0.y; a.z += o0.z; a.w += o0.w;
                        }
                        stg_f4_stream(dst, a);
                    }
                }
            }
        }
    }
        } while (0);
    }  // units
}

// ------------------------------------------------------------------------------------------------
// NCHW feature maps (what an unmodified model.py produces: conv outputs in torch's default memory format) with NCHW
// crops - the literal drop-in's layouts.  In a channel plane one bilinear tap is 4 bytes; what IS contiguous is the
// footprint row of a RoI: the p bin columns of one output row read two short runs (x_lo / x_hi of every bin) of ONE
// feature row.  The kernels below therefore put the bin COLUMNS on lanes and walk the bin ROWS:
//
//   forward   CTA = RoI x 64 channels, warp = 8 consecutive channels, lane = (channel group, bin column), K channels per
//             thread.  The blend is evaluated separably, exactly as crop_cpu.cpp:107-110 writes it: top = H(y_lo),
//             bot = H(y_hi) with H(r) = v[r][x_lo] + (v[r][x_hi] - v[r][x_lo]) * x_lerp - the same three roundings, so the
//             result is bit-identical - and H of a feature row is kept in registers while consecutive bin rows reuse it
//             (an up-sampled RoI touches ~p + 1 rows, not 2p: 2 loads per NEW row instead of 4 per bin).  All row
//             decisions are warp-uniform (one RoI per CTA).  The warp's 8 x p x p outputs are staged in shared memory in
//             the output's own order and leave with ONE bulk async copy (cp.async.bulk.global.shared::cta, TMA; SASS
//             UBLKCP): whole 16-byte-aligned runs, no partial sectors, no LSU store traffic.
//   backward  CTA = RoI x 64 channels; the CTA's [64][p*p] slice of the upstream gradient is contiguous in NCHW and
//             arrives with one bulk async copy (mbarrier complete_tx).  warp = channel, lane = FEATURE COLUMN of the
//             RoI's footprint: a lane sums the bins that touch its column (their x_lo == column, or x_hi == column), walks
//             the bin rows keeping the sums of the current and the next feature row in registers, and flushes a row when
//             the walk leaves it - one coalesced red.global.add.f32 per (channel, feature row) run of columns instead of
//             four scattered atomics per bin (crop_cuda.cu:151-168's design).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_taps_nchw(const RoiCtx& ctx, const float4 box, int ph, int pw, TapS* s_ty, TapS* s_tx) {
    const int tid = threadIdx.x;
    if (tid < ph) {
        const AxisTap t = axis_tap(box.x, box.z, ctx.H, ph, tid);
        TapS o;
        o.valid = (t.lo >= 0) && ctx.ok;
        o.lo = o.valid ? t.lo * ctx.W : 0;   // element offset of the row inside a channel plane
        o.hi = o.valid ? t.hi * ctx.W : 0;
        o.lerp = t.lerp;
        s_ty[tid] = o;
    } else if (tid >= 64 && tid < 64 + pw) {
        const AxisTap t = axis_tap(box.y, box.w, ctx.W, pw, tid - 64);
        TapS o;
        o.valid = (t.lo >= 0) && ctx.ok;
        o.lo = o.valid ? t.lo : 0;
        o.hi = o.valid ? t.hi : 0;
        o.lerp = t.lerp;
        s_tx[tid - 64] = o;
    }
}

// dst (global) <- src (shared), `bytes` a multiple of 16, both 16-byte aligned: one TMA bulk copy issued by the calling
// thread, which also waits until the source has been read (the CTA may then exit or reuse the buffer).
__device__ __forceinline__ void bulk_s2g_and_wait(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    // the crops are a write-once stream (1.6 GB per configs[3] forward): evict-first in L2, so that they do not push out the
    // feature maps other RoIs of the image are about to read
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the generic-proxy smem writes above -> visible to the async proxy
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes), "l"(policy)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// (Fetching two feature rows ahead of the walk - registers A / B alternating, 64 registers - measured slower: 7x7 0.424 ms against
// 0.335 ms, 14x14 0.881 against 0.859 at the configs[3] geometry; the simple walk below stays.)
constexpr int kNchwCW = 8;  // channels per warp

template <int POOL>
__global__ void __launch_bounds__(256) roialign_fwd_nchw_kernel(const RoiParams p) {
    constexpr int LG = (POOL > 8) ? 16 : 8;  // lanes per channel group (>= POOL)
    constexpr int G = 32 / LG;               // channel groups per warp
    constexpr int K = kNchwCW / G;           // channels per thread
    constexpr int P2 = POOL * POOL;
    static_assert(POOL <= 16, "bin columns live on the lanes of a channel group");
    extern __shared__ __align__(128) float s_out[];  // [8 warps][8 channels][P2]: the outputs in global order
    __shared__ TapS s_ty[POOL];
    __shared__ TapS s_tx[POOL];

    const int chunks = (p.C + kChunk - 1) / kChunk;
    const int n = blockIdx.x / chunks;
    const int c0 = (blockIdx.x - n * chunks) * kChunk;
    const int tid = threadIdx.x;
    const int C = p.C;

    float4 box;
    const RoiCtx ctx = select_level(p, n, box);
    stage_taps_nchw(ctx, box, POOL, POOL, s_ty, s_tx);
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int cw0 = c0 + warp * kNchwCW;              // the warp's first channel
    const int cc = min(kNchwCW, C - cw0);             // channels this warp owns (<= 0: nothing)
    if (cc <= 0) return;
    const int g = lane / LG, x = lane - g * LG;
    float* wout = s_out + warp * (kNchwCW * P2);
    const bool col = x < POOL;
    const TapS tx = s_tx[col ? x : 0];
    const bool x_in = col && tx.valid;
    // 32-bit element offsets from the image's base (the launcher guarantees C * H * W < 2^31): one IMAD.WIDE per load
    const unsigned plane = (unsigned)(ctx.H * ctx.W);
    const float* base = ctx.base;
    unsigned olo[K], ohi[K];
    bool live[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int cl = g * K + k;                     // channel inside the warp's 8: banks of the groups do not collide
        live[k] = x_in && cl < cc;
        const unsigned o = (unsigned)(cw0 + (cl < cc ? cl : 0)) * plane;
        olo[k] = o + (unsigned)tx.lo;
        ohi[k] = o + (unsigned)tx.hi;
    }

    int ra = -1, rb = -1;  // feature rows (as plane offsets) whose horizontal blends Ha / Hb hold
    float Ha[K], Hb[K];
#pragma unroll
    for (int k = 0; k < K; ++k) Ha[k] = Hb[k] = 0.f;

#pragma unroll 1
    for (int y = 0; y < POOL; ++y) {
        const TapS ty = s_ty[y];  // warp-uniform
        float v[K];
        if (ty.valid) {
            if (ty.lo == rb) {            // the walk moved down one row: yesterday's bottom is today's top
#pragma unroll
                for (int k = 0; k < K; ++k) Ha[k] = Hb[k];
                ra = rb;
                rb = -1;
            } else if (ty.lo != ra) {
                float lo[K], hi[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    MRCNN_DBG(!live[k] || ohi[k] + (unsigned)ty.lo < (unsigned)C * plane);
                    lo[k] = live[k] ? __ldg(base + (olo[k] + (unsigned)ty.lo)) : 0.f;
                    hi[k] = live[k] ? __ldg(base + (ohi[k] + (unsigned)ty.lo)) : 0.f;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) Ha[k] = __fadd_rn(lo[k], __fmul_rn(__fsub_rn(hi[k], lo[k]), tx.lerp));
                ra = ty.lo;
            }
            if (ty.hi == ra) {            // y_lerp == 0: the ceil row is the floor row
#pragma unroll
                for (int k = 0; k < K; ++k) Hb[k] = Ha[k];
                rb = ra;
            } else if (ty.hi != rb) {
                float lo[K], hi[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    lo[k] = live[k] ? __ldg(base + (olo[k] + (unsigned)ty.hi)) : 0.f;
                    hi[k] = live[k] ? __ldg(base + (ohi[k] + (unsigned)ty.hi)) : 0.f;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) Hb[k] = __fadd_rn(lo[k], __fmul_rn(__fsub_rn(hi[k], lo[k]), tx.lerp));
                rb = ty.hi;
            }
#pragma unroll
            for (int k = 0; k < K; ++k)
                v[k] = x_in ? __fadd_rn(Ha[k], __fmul_rn(__fsub_rn(Hb[k], Ha[k]), ty.lerp)) : p.extrap;
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] = p.extrap;
        }
        if (col) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                MRCNN_DBG((g * K + k) * P2 + y * POOL + x < kNchwCW * P2);
                wout[(g * K + k) * P2 + y * POOL + x] = v[k];
            }
        }
    }
    fence_proxy_async();   // every lane's staged outputs -> visible to the async proxy before lane 0 issues the bulk store
    __syncwarp();
    float* dst = p.crops + ((size_t)n * C + cw0) * P2;
    const uint32_t bytes = (uint32_t)(cc * P2 * sizeof(float));
    if (((reinterpret_cast<uintptr_t>(dst) | bytes) & 15u) == 0) {
        if (lane == 0) bulk_s2g_and_wait(dst, wout, bytes);
    } else {  // odd channel counts / unaligned outputs: plain coalesced stores
        for (int i = lane; i < cc * P2; i += 32) __stcs(dst + i, wout[i]);
    }
}

// Backward for NCHW gradient maps and NCHW upstream gradients.  grid = N x ceil(C / 64), 256 threads; dynamic smem =
// the CTA's gradient slice [64][P2] (one bulk async copy).  Same thread layout as the forward: lane = (channel group,
// bin column), K channels per thread, bin rows walked top to bottom.  The adjoint is applied separably: a thread first
// sums, down its bin column, everything that lands on one FEATURE ROW (running sums for the current row and the one below,
// flushed when the walk leaves a row - an up-sampled RoI has ~p + 1 rows for 2p taps), then spreads a finished row sum over
// its two feature columns: two red.global.add.f32 per thread and finished row, lanes on neighbouring addresses of one
// plane row.  History (configs[3] 14x14 backward, kernel time; profiles/r02_nchw_history.txt): lanes on feature columns with
// per-lane bin lists 5.3 ms (issue-bound); this version 1.19 ms (L1/TEX red requests: 112 M); a third one that regrouped a
// finished row by feature column through shared memory (39 M requests) 1.81 ms - the extra shared-memory reads and 85
// registers cost more than the saved requests.
template <int POOL>
__global__ void __launch_bounds__(256) roialign_bwd_nchw_kernel(const RoiParams p) {
    constexpr int LG = (POOL > 8) ? 16 : 8;
    constexpr int G = 32 / LG;
    constexpr int K = kNchwCW / G;
    constexpr int P2 = POOL * POOL;
    extern __shared__ __align__(128) float s_g[];  // [64][P2] upstream gradient slice, global order
    __shared__ TapS s_ty[POOL];
    __shared__ TapS s_tx[POOL];
    __shared__ __align__(8) uint64_t s_bar;

    const int chunks = (p.C + kChunk - 1) / kChunk;
    const int n = blockIdx.x / chunks;
    const int c0 = (blockIdx.x - n * chunks) * kChunk;
    const int tid = threadIdx.x;
    const int C = p.C;
    const int cc_cta = min(kChunk, C - c0);

    const float* gsrc = p.crops + ((size_t)n * C + c0) * P2;
    const uint32_t bytes = (uint32_t)(cc_cta * P2 * sizeof(float));
    const bool bulk = ((reinterpret_cast<uintptr_t>(gsrc) | bytes) & 15u) == 0;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (bulk) {
        if (tid == 0) {
            mbar_expect_tx(&s_bar, bytes);
            bulk_g2s(s_g, gsrc, bytes, &s_bar);
        }
    } else {
        for (int i = tid; i < cc_cta * P2; i += 256) s_g[i] = __ldcs(gsrc + i);
    }

    float4 box;
    const RoiCtx ctx = select_level(p, n, box);
    stage_taps_nchw(ctx, box, POOL, POOL, s_ty, s_tx);
    __syncthreads();
    if (bulk) mbar_wait(&s_bar, 0);
    if (!ctx.ok) return;

    const int warp = tid >> 5, lane = tid & 31;
    const int cw = warp * kNchwCW;                    // first channel of the warp inside the CTA's slice
    const int cc = min(kNchwCW, cc_cta - cw);
    if (cc <= 0) return;
    const int g = lane / LG, x = lane - g * LG;
    const bool col = x < POOL;
    const TapS tx = s_tx[col ? x : 0];
    const bool x_in = col && tx.valid;
    const float wl = __fsub_rn(1.0f, tx.lerp), wh = tx.lerp;
    const bool two_cols = tx.lerp != 0.0f;            // <=> hi == lo + 1
    const size_t plane = (size_t)ctx.H * ctx.W;
    const int W = ctx.W;
    float* dlo[K];
    const float* gk[K];
    bool live[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int cl = g * K + k;
        live[k] = x_in && cl < cc;
        dlo[k] = ctx.base + (size_t)(c0 + cw + (cl < cc ? cl : 0)) * plane + tx.lo;
        gk[k] = s_g + (cw + (cl < cc ? cl : 0)) * P2 + (col ? x : 0);
    }
    const int dx = tx.hi - tx.lo;

    int ra = -1;           // plane offset of the feature row summed in A; Bn is the row below it
    bool hasB = false;
    float A[K], Bn[K];
#pragma unroll
    for (int k = 0; k < K; ++k) A[k] = Bn[k] = 0.f;

    auto flush = [&](const float (&V)[K], int row) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (live[k]) {
                atomicAdd(dlo[k] + row, __fmul_rn(wl, V[k]));
                if (two_cols) atomicAdd(dlo[k] + row + dx, __fmul_rn(wh, V[k]));
            }
        }
    };

#pragma unroll 1
    for (int y = 0; y < POOL; ++y) {
        const TapS ty = s_ty[y];  // warp-uniform
        if (!ty.valid) continue;
        float gv[K];
#pragma unroll
        for (int k = 0; k < K; ++k) gv[k] = gk[k][y * POOL];
        if (ty.lo != ra) {
            bool carry = false;
            if (ra >= 0) {
                flush(A, ra);
                if (hasB) {
                    if (ty.lo == ra + W) carry = true;   // the walk moved down exactly one row: B becomes A
                    else flush(Bn, ra + W);
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                A[k] = carry ? Bn[k] : 0.f;
                Bn[k] = 0.f;
            }
            hasB = false;
            ra = ty.lo;
        }
        const float w0 = __fsub_rn(1.0f, ty.lerp);
#pragma unroll
        for (int k = 0; k < K; ++k) A[k] = fmaf(w0, gv[k], A[k]);
        if (ty.lerp != 0.0f) {  // <=> hi == lo + one row
#pragma unroll
            for (int k = 0; k < K; ++k) Bn[k] = fmaf(ty.lerp, gv[k], Bn[k]);
            hasB = true;
        }
    }
    if (ra >= 0) {
        flush(A, ra);
        if (hasB) flush(Bn, ra + W);
    }
}

// ------------------------------------------------------------------------------------------------
// Few-channel NCHW crops (the 28x28 mask targets of mrn_samples, model.py:501-502: C = 1, image = [P,1,1024,1024] gt masks):
// one CTA per (crop, channel plane).  The box and the ph + pw axis taps are computed ONCE per crop (the strided kernel below
// recomputes both divisions per output scalar); every thread then blends four bins per pass with its 16 tap loads in flight
// together.  Output runs are contiguous ([n][c][y][x]).  Latency-bound by construction - 784 outputs per crop, taps scattered
// over a 4 MB mask - so the point is to start all of a crop's loads at once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) crop_plane_fwd_kernel(const RoiParams p) {
    __shared__ TapS s_ty[kMaxPool];
    __shared__ TapS s_tx[kMaxPool];
    const int C = p.C;
    const int n = blockIdx.x / C;
    const int c = blockIdx.x - n * C;
    const int tid = threadIdx.x;
    const int ph = p.ph, pw = p.pw, P2 = ph * pw;

    float4 box;
    const RoiCtx ctx = select_level(p, n, box);
    stage_taps_nchw(ctx, box, ph, pw, s_ty, s_tx);
    __syncthreads();
    const float* plane = ctx.base + (size_t)c * ctx.H * ctx.W;
    float* out = p.crops + ((size_t)n * C + c) * P2;
    for (int b0 = tid; b0 < P2; b0 += 4 * 256) {
        float tl[4], tr[4], bl[4], br[4], xl[4], yl[4];
        bool in[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int b = b0 + u * 256;
            in[u] = false;
            if (b < P2) {
                const int y = b / pw, x = b - y * pw;
                const TapS ty = s_ty[y], tx = s_tx[x];
                in[u] = ty.valid && tx.valid;
                if (in[u]) {
                    tl[u] = __ldg(plane + ty.lo + tx.lo);
                    tr[u] = __ldg(plane + ty.lo + tx.hi);
                    bl[u] = __ldg(plane + ty.hi + tx.lo);
                    br[u] = __ldg(plane + ty.hi + tx.hi);
                    xl[u] = tx.lerp;
                    yl[u] = ty.lerp;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int b = b0 + u * 256;
            if (b < P2) out[b] = in[u] ? bilerp(tl[u], tr[u], bl[u], br[u], xl[u], yl[u]) : p.extrap;
        }
    }
}

// Upstream gradients [N][C][P2] (NCHW-contiguous, what torch's conv backward hands to an unmodified model.py) -> [N][P2][C]
// for the gather backward: one CTA per (RoI, 64 channels); the slice is contiguous on the way in (coalesced 128-bit loads into
// a padded shared tile) and 256-byte pieces on the way out.
__global__ void __launch_bounds__(kThreads) grads_nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int P2) {
    extern __shared__ __align__(16) float tile[];
    const int n = blockIdx.x, c0 = blockIdx.y * kChunk;
    const int cc = min(kChunk, C - c0);
    const int P2pad = P2 | 1;
    tile_copy<true>(tile, const_cast<float*>(src) + ((size_t)n * C + c0) * P2, cc * P2, P2, P2pad);
    __syncthreads();
    const int lane = threadIdx.x & (kLanes - 1), slot = threadIdx.x >> 4;
    const int c = 4 * lane;
    if (c >= cc) return;   // C % 4 == 0
    float* o = dst + (size_t)n * P2 * C + c0 + c;
    const float* t = tile + c * P2pad;
    for (int b = slot; b < P2; b += kSlots)
        stg_f4_stream(o + (size_t)b * C, make_float4(t[b], t[P2pad + b], t[2 * P2pad + b], t[3 * P2pad + b]));
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

// Which forward kernel serves channels-last 14x14: 0 = column-stationary (the default: fastest measured), 2 = row-walking with
// 128 channels per CTA, 9 = TMA ring pipeline.  Fixed at first use; MRCNN_FWD14 selects the alternatives for experiments and for
// the bit-identity test (results are identical).
static int fwd14_variant() {
    const char* e = getenv("MRCNN_FWD14");
    if (e == nullptr) return 0;
    if (!strcmp(e, "col")) return 0;
    if (!strcmp(e, "row")) return 2;
    if (!strcmp(e, "tma")) return 9;
    return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int launch_roi(const RoiParams& p_in, int image_layout, int crops_layout, bool backward, cudaStream_t stream) {
    if (p_in.N == 0) return MRCNN_OK;
    RoiParams p = p_in;
    p.negzero = -0.0f;
    bool fast = image_layout == MRCNN_NHWC && (p.C % 4) == 0 && p.ph <= kMaxPool && p.pw <= kMaxPool &&
                aligned16(p.crops);
    for (int l = 0; l < (p.pyramid ? 4 : 1); ++l) fast = fast && aligned16(p.lv[l].ptr);
    const int P2 = p.ph * p.pw;
    const size_t smem = (crops_layout == MRCNN_NHWC) ? 0 : sizeof(float) * kChunk * (size_t)(P2 | 1);
    if (fast && smem > 200 * 1024) fast = false;
    for (int l = 0; l < (p.pyramid ? 4 : 1); ++l)  // the vectorised kernels use 32-bit element offsets inside one image
        if ((long long)p.lv[l].H * p.lv[l].W * p.C >= (1ll << 31)) fast = false;
    if ((long long)P2 * p.C >= (1ll << 31)) fast = false;
    // NCHW feature maps with NCHW crops / gradients at the two head sizes: the row-walking kernels
    bool nchw_fast = image_layout == MRCNN_NCHW && crops_layout == MRCNN_NCHW && p.ph == p.pw && (p.ph == 7 || p.ph == 14);
    for (int l = 0; l < (p.pyramid ? 4 : 1); ++l)
        if ((long long)p.C * p.lv[l].H * p.lv[l].W >= (1ll << 31)) nchw_fast = false;   // 32-bit element offsets inside one image
    const long long nchw_grid = (long long)p.N * ((p.C + kChunk - 1) / kChunk);
    if (nchw_grid >= (1ll << 31)) nchw_fast = false;
    // few-channel NCHW forward (mask targets): one CTA per (crop, plane)
    if (!backward && !nchw_fast && image_layout == MRCNN_NCHW && crops_layout == MRCNN_NCHW && p.C <= 4 && p.ph <= kMaxPool &&
        p.pw <= kMaxPool && (long long)p.N * p.C < (1ll << 31)) {
        bool fits = true;
        for (int l = 0; l < (p.pyramid ? 4 : 1); ++l)
            if ((long long)p.lv[l].H * p.lv[l].W >= (1ll << 31)) fits = false;
        if (fits) {
            crop_plane_fwd_kernel<<<(unsigned)(p.N * p.C), 256, 0, stream>>>(p);
            MRCNN_LAUNCH_CHECK();
            return MRCNN_OK;
        }
    }
    if (nchw_fast) {
        const size_t smem_n = sizeof(float) * kChunk * (size_t)P2;
#define MRCNN_LAUNCH_NCHW(KERNEL)                                                                                \
    do {                                                                                                         \
        if (smem_n > 48 * 1024)                                                                                  \
            MRCNN_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_n));  \
        KERNEL<<<(unsigned)nchw_grid, 256, smem_n, stream>>>(p);                                                 \
    } while (0)
        if (!backward) {
            if (p.ph == 7) MRCNN_LAUNCH_NCHW(roialign_fwd_nchw_kernel<7>);
            else MRCNN_LAUNCH_NCHW(roialign_fwd_nchw_kernel<14>);
        } else {
            if (p.ph == 7) MRCNN_LAUNCH_NCHW(roialign_bwd_nchw_kernel<7>);
            else MRCNN_LAUNCH_NCHW(roialign_bwd_nchw_kernel<14>);
        }
#undef MRCNN_LAUNCH_NCHW
        MRCNN_LAUNCH_CHECK();
        return MRCNN_OK;
    }
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
            if (crops_layout == MRCNN_NHWC) {
                // 16 lanes (64 channels) per CTA measured best on B200: 8 -> 504 us, 16 -> 468 us, 32 -> 490 us, 64 -> 509 us (14x14)
                const long long flat = (long long)grid.x * grid.y;
                const bool ok = flat < (1ll << 31);
                // pool 7: 128-thread CTAs (7 of 8 slots own a column): 210 -> 178 us, more CTAs resident hide the prologue;
                // pool 14: 256 threads, 14 of 16 slots; splitting its columns (rows) over two CTAs measured 474 (543) vs 468 us
                if (ok && p.ph == 7 && p.pw == 7) roialign_fwd_nhwc_col_kernel<7, kLanes, 8><<<(unsigned)flat, 8 * kLanes, 0, stream>>>(p);
                else if (ok && p.ph == 14 && p.pw == 14) {
                    // MRCNN_FWD14 = col (default) | row | tma: the measured alternatives (profiles/r02_experiments.txt), bit-identical
                    static const int variant = fwd14_variant();
                    const int K = (variant == 2) ? 2 : 0;
                    const unsigned grid_k = K ? (unsigned)((long long)p.N * ((p.C + 64 * K - 1) / (64 * K))) : 0;
                    if (variant == 9 && p.C <= 256 && (long long)p.N < (1ll << 31)) {
                        const size_t smem_t = kTmaRingBytes + (size_t)kTmaOutBufs * 14 * p.C * sizeof(float);
                        MRCNN_CUDA(cudaFuncSetAttribute(roialign_fwd_nhwc_tma_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
                        roialign_fwd_nhwc_tma_kernel<14><<<(unsigned)p.N, 288, smem_t, stream>>>(p);
                    } else
                    if (K == 2) roialign_fwd_nhwc_row_kernel<14, 2><<<grid_k, 256, 0, stream>>>(p);
                    else roialign_fwd_nhwc_col_kernel<14, kLanes, kSlots><<<(unsigned)flat, kThreads, 0, stream>>>(p);
                }
                else MRCNN_LAUNCH_NHWC((roialign_fwd_nhwc_kernel<0, true>));
            } else if ((p.ph == 7 && p.pw == 7) || (p.ph == 14 && p.pw == 14)) {
                const long long flat = (long long)grid.x * grid.y;
                if (flat >= (1ll << 31)) return fail(MRCNN_E_INVALID_ARG, "too many RoIs");
                // stage_taps uses threads [0, pool) and [64, 64 + pool): 128-thread CTAs are enough
                const bool narrow = p.N <= 4096;     // small launches: 32-channel CTAs (twice as many, half the tile)
                const size_t smem_t = narrow ? smem / 2 + 64 : smem;
                const unsigned grid_t = narrow ? (unsigned)((long long)p.N * ((p.C + 31) / 32)) : (unsigned)flat;
                if (p.ph == 7) {
                    if (narrow) roialign_fwd_nhwc_colt_kernel<7, 8><<<grid_t, 128, smem_t, stream>>>(p);
                    else roialign_fwd_nhwc_colt_kernel<7, 16><<<grid_t, 256, smem_t, stream>>>(p);
                } else if (narrow) {
                    roialign_fwd_nhwc_colt_kernel<14, 8><<<grid_t, 128, smem_t, stream>>>(p);
                } else {
                    MRCNN_CUDA(cudaFuncSetAttribute(roialign_fwd_nhwc_colt_kernel<14, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
                    roialign_fwd_nhwc_colt_kernel<14, 16><<<grid_t, 256, smem_t, stream>>>(p);
                }
            } else MRCNN_DISPATCH_POOL(roialign_fwd_nhwc_kernel, false);
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
    int* cnt;
    int* pos;
    unsigned int* cursor;
    QItem* items;
    QItem* sorted;       // deterministic plans only: scratch of the per-unit item sort, same capacity as items
    int* heavy;          // deterministic plans only: [units] queue of the units sorted by a whole CTA (its length: cursor[1])
    size_t clear_bytes;  // cnt + cursor are cleared before every call (they are adjacent)
    size_t bytes;
};

// Deterministic plans (mrcnn_set_deterministic / MRCNN_DETERMINISTIC=1): the fill pass places a unit's items through atomic
// cursors, so their order - the fp32 summation order of the gather - depends on scheduling.  With the switch on, every plan ends
// with a pass that rewrites each unit's items in ascending (gradient offset, column) order: two plans of the same boxes then give
// bit-identical gradients.  It costs a second item array in the workspace and one short kernel per plan.
static int g_deterministic = -1;
static bool deterministic_plans() {
    if (g_deterministic < 0) {
        const char* e = getenv("MRCNN_DETERMINISTIC");
        g_deterministic = (e && e[0] == '1') ? 1 : 0;
    }
    return g_deterministic == 1;
}

static long long gather_units(const int H[4], const int W[4], int B) {
    long long units = 0;
    for (int l = 0; l < 4; ++l) units += (long long)B * H[l] * ((W[l] + kGTile - 1) / kGTile);
    return units;
}

static GatherWorkspace carve_gather(void* base, const int H[4], const int W[4], int B, int N, long long bins) {
    GatherWorkspace w;
    const size_t units = (size_t)gather_units(H, W, B);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* q = base ? (void*)((char*)base + off) : nullptr;
        off += align_up(bytes, 256);
        return q;
    };
    w.cursor = (unsigned int*)take(8);   // [0] the fill cursor, [1] the length of the heavy-unit queue (deterministic plans)
    w.cnt = (int*)take(units * 4);
    w.clear_bytes = off;
    w.pos = (int*)take(units * 4);
    // every bin yields at most four work items (two rows x a unit boundary); bins = sum of pool^2 over the heads
    const size_t item_bytes = (size_t)4 * (size_t)(N > 0 ? N : 1) * (size_t)bins * sizeof(QItem);
    w.items = (QItem*)take(item_bytes);
    w.sorted = deterministic_plans() ? (QItem*)take(item_bytes) : nullptr;
    w.heavy = deterministic_plans() ? (int*)take(units * 4) : nullptr;
    w.bytes = off;
    return w;
}

// true if the gather backward can serve this call (bins = sum of pool^2 over the heads, max_pool the largest pool)
static bool gather_eligible(const int H[4], const int W[4], int B, int C, int N, long long bins, int max_pool, int gfm_layout,
                            int grads_layout, const float* grads, float* const gfm[4], const void* workspace,
                            size_t workspace_bytes) {
    // NCHW gradient maps are written sector-wise by the gather itself; NCHW upstream gradients are transposed into the tail of the
    // workspace first (carve_gather_ex).  Both need W % 8 == 0 for aligned sectors to be the common case (any W is handled).
    (void)gfm_layout;
    if (grads_layout == MRCNN_NCHW && sizeof(float) * kChunk * (size_t)((max_pool * max_pool) | 1) > 200 * 1024) return false;
    if ((C % 4) != 0 || N <= 0) return false;
    if ((long long)N * max_pool * max_pool * C >= (1ll << 31)) return false;  // 32-bit element offsets into grads
    if ((long long)4 * N * bins >= (1ll << 31)) return false;                 // 32-bit item positions
    if (gather_units(H, W, B) >= (1ll << 31) - 8) return false;
    if ((long long)N * max_pool * 2 >= (1ll << 31) * 256) return false;
    {
        size_t need = carve_gather(nullptr, H, W, B, N, bins).bytes;
        if (grads_layout == MRCNN_NCHW) need += align_up((size_t)N * C * (size_t)bins * sizeof(float), 256);
        if (workspace == nullptr || workspace_bytes < need) return false;
    }
    if (!aligned16(grads) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return false;
    for (int l = 0; l < 4; ++l)
        if (!aligned16(gfm[l])) return false;
    return true;
}

// Geometry + workspace pointers of a gather backward (everything but the tensors).
static GatherParams gather_params(const GatherWorkspace& ws, const int H[4], const int W[4], int B, int C, int N, float image_area) {
    GatherParams g = {};
    int base = 0;
    for (int l = 3; l >= 0; --l) {  // coarse levels first: their units collect the most work
        g.g[l].ptr = nullptr;
        g.g[l].H = H[l];
        g.g[l].W = W[l];
        g.g[l].segs = (W[l] + kGTile - 1) / kGTile;
        g.g[l].unit_base = base;
        base += B * H[l] * g.g[l].segs;
    }
    g.units = base;
    g.rule = make_level_rule(image_area);
    g.B = B; g.C = C; g.N = N;
    g.cnt = ws.cnt; g.pos = ws.pos; g.cursor = ws.cursor; g.items = ws.items;
    g.err = device_error_word();
    return g;
}

// Deterministic plans: every unit's items rewritten in ascending (gradient offset, column | head) order, through the scratch array
// and back.  Keys are unique inside a unit (a bin reaches a unit at most once as a main item and never also as a spill item); the
// index tie-break keeps the light pass a permutation even if they were not.
//   light units (<= kSortLight items: nearly all of them, ~30 items on average): one warp per unit, rank sort (every item counts
//     the items that precede it);
//   heavy units (the coarse levels collect thousands of items per unit): queued by the light pass, then one CTA per unit -
//     a bitonic network over (key, position) pairs in shared memory up to kSortCap items, a CTA-wide rank sort beyond.
constexpr int kSortLight = 64;
constexpr int kSortCap = 4096;

__device__ __forceinline__ uint64_t item_key(const QItem& it) { return ((uint64_t)(uint32_t)it.off << 32) | (uint32_t)it.idx; }

__global__ void __launch_bounds__(256) bwd_sort_items_kernel(const GatherParams p, QItem* __restrict__ scratch, int* __restrict__ heavy,
                                                             unsigned int* __restrict__ heavy_count) {
    const int u = (int)(((long long)blockIdx.x * 256 + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (u >= p.units) return;
    const int n = p.cnt[u];
    if (n < 2) return;
    if (n > kSortLight) {
        if (lane == 0) heavy[atomicAdd(heavy_count, 1u)] = u;   // the ORDER of this queue does not matter: units are independent
        return;
    }
    const int beg = p.pos[u] - n;
    MRCNN_DBG(beg >= 0);
    QItem* src = p.items + beg;
    QItem* tmp = scratch + beg;
    for (int i = lane; i < n; i += 32) {
        const QItem it = src[i];
        const uint64_t key = item_key(it);
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const uint64_t kj = item_key(src[j]);   // the same address for every lane: one broadcast load
            rank += (kj < key || (kj == key && j < i)) ? 1 : 0;
        }
        MRCNN_DBG(rank >= 0 && rank < n);
        tmp[rank] = it;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) src[i] = tmp[i];
}

__global__ void __launch_bounds__(256) bwd_sort_heavy_kernel(const GatherParams p, QItem* __restrict__ scratch, const int* __restrict__ heavy,
                                                             const unsigned int* __restrict__ heavy_count) {
    __shared__ uint64_t s_key[kSortCap];
    __shared__ uint16_t s_idx[kSortCap];
    const int tid = threadIdx.x;
    const int count = (int)*heavy_count;
    for (int q = blockIdx.x; q < count; q += gridDim.x) {
        const int u = heavy[q];
        const int n = p.cnt[u];
        const int beg = p.pos[u] - n;
        MRCNN_DBG(u >= 0 && u < p.units && n > kSortLight && beg >= 0);
        QItem* src = p.items + beg;
        QItem* tmp = scratch + beg;
        if (n <= kSortCap) {
            int P = 128;
            while (P < n) P <<= 1;
            for (int i = tid; i < P; i += 256) {
                s_key[i] = i < n ? item_key(src[i]) : ~0ull;   // padding sorts to the end (a real key never has all bits set: idx <= 31)
                s_idx[i] = (uint16_t)i;
            }
            __syncthreads();
            for (int k = 2; k <= P; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < (P >> 1); t += 256) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        const int l = i | j;
                        const bool up = (i & k) == 0;
                        const uint64_t a = s_key[i], b = s_key[l];
                        if ((a > b) == up) {
                            s_key[i] = b;
                            s_key[l] = a;
                            const uint16_t x = s_idx[i];
                            s_idx[i] = s_idx[l];
                            s_idx[l] = x;
                        }
                    }
                    __syncthreads();
                }
            }
            for (int i = tid; i < n; i += 256) {
                MRCNN_DBG(s_idx[i] < n);
                tmp[i] = src[s_idx[i]];
            }
        } else {   // beyond the shared-memory network: every item counts its predecessors (n^2 / 256 steps per thread)
            for (int i = tid; i < n; i += 256) {
                const QItem it = src[i];
                const uint64_t key = item_key(it);
                int rank = 0;
                for (int j = 0; j < n; ++j) {
                    const uint64_t kj = item_key(src[j]);
                    rank += (kj < key || (kj == key && j < i)) ? 1 : 0;
                }
                tmp[rank] = it;
            }
        }
        __syncthreads();
        for (int i = tid; i < n; i += 256) src[i] = tmp[i];
        __syncthreads();
    }
}

// Passes 1-3: the work-item queues of every unit.  They depend on the boxes, the pool sizes, C and the pyramid geometry
// only - not on the gradients - so a training step can build them while the forward runs (mrcnn_..._backward_plan).
static int launch_gather_plan(GatherParams g, const GatherWorkspace& ws, void* workspace, int heads, const int pools[2],
                              const float* boxes, const int32_t* box_index, cudaStream_t stream) {
    g.boxes = boxes; g.box_index = box_index;
    MRCNN_CUDA(cudaMemsetAsync(workspace, 0, ws.clear_bytes, stream));
    const unsigned ugrid = (unsigned)((g.units + 255) / 256);
    for (int pass = 0; pass < 2; ++pass) {  // count every head, allocate, fill every head
        for (int h = 0; h < heads; ++h) {
            g.ph = pools[h]; g.pw = pools[h];
            g.head_flag = h ? kOHead2 : 0;
            const long long walkers = (long long)g.N * pools[h] * 2;
            const unsigned wgrid = (unsigned)((walkers + 255) / 256);
            if (pass == 0) bwd_items_kernel<false><<<wgrid, 256, 0, stream>>>(g);
            else bwd_items_kernel<true><<<wgrid, 256, 0, stream>>>(g);
            MRCNN_LAUNCH_CHECK();
        }
        if (pass == 0) {
            bwd_alloc_kernel<<<ugrid, 256, 0, stream>>>(g);
            MRCNN_LAUNCH_CHECK();
        }
    }
    if (ws.sorted != nullptr) {
        bwd_sort_items_kernel<<<(unsigned)(((long long)g.units * 32 + 255) / 256), 256, 0, stream>>>(g, ws.sorted, ws.heavy, ws.cursor + 1);
        MRCNN_LAUNCH_CHECK();
        bwd_sort_heavy_kernel<<<sm_count() * 4, 256, 0, stream>>>(g, ws.sorted, ws.heavy, ws.cursor + 1);
        MRCNN_LAUNCH_CHECK();
    }
    return MRCNN_OK;
}

// Pass 4: one warp per unit sums the bins its queue lists and writes its 8 pixels once.
static int launch_gather_run(GatherParams g, int heads, const float* const grads[2], float* const gfm[4], int accumulate,
                             cudaStream_t stream, int out_nchw = 0) {
    g.out_nchw = out_nchw;
    for (int l = 0; l < 4; ++l) g.g[l].ptr = gfm[l];
    g.grads = grads[0]; g.grads2 = grads[heads - 1];
    const bool wide = (g.C % 256) == 0;  // 8 channels per lane: one pass covers 256 channels
    // One one-warp CTA per unit: the block scheduler balances the uneven units (coarse levels collect most items) better than a
    // persistent grid striding over them - MRCNN_GATHER_CTAS = 16 / 24 / 32 / 64 CTAs per SM measured 0.84 / 0.82 / 0.73 / 0.65 ms
    // against 0.58 ms for the configs[3] 14x14 backward, and taking units from an atomic counter 0.66 ms (0.55 ms for the 7x7 head
    // against 0.31: 174,080 atomics on one word) - profiles/r02_experiments.txt.  The kernel loops over units all the same.
    static const int per_sm = getenv("MRCNN_GATHER_CTAS") ? atoi(getenv("MRCNN_GATHER_CTAS")) : 0;
    const long long want = per_sm > 0 ? (long long)sm_count() * per_sm : (long long)g.units;
    const unsigned grid = (unsigned)(want < (long long)g.units ? want : (long long)g.units);
    // U = 2 items per stage, ST = 4 stages: best of the (U, ST) grid measured on B200 (profiles/r01_*gather*)
    if (out_nchw) {
        if (wide && accumulate) roialign_bwd_gather_kernel<2, 2, 4, true, true><<<grid, 32, 0, stream>>>(g);
        else if (wide) roialign_bwd_gather_kernel<2, 2, 4, false, true><<<grid, 32, 0, stream>>>(g);
        else if (accumulate) roialign_bwd_gather_kernel<1, 2, 4, true, true><<<grid, 32, 0, stream>>>(g);
        else roialign_bwd_gather_kernel<1, 2, 4, false, true><<<grid, 32, 0, stream>>>(g);
    } else if (wide && accumulate) roialign_bwd_gather_kernel<2, 2, 4, true><<<grid, 32, 0, stream>>>(g);
    else if (wide) roialign_bwd_gather_kernel<2, 2, 4, false><<<grid, 32, 0, stream>>>(g);
    else if (accumulate) roialign_bwd_gather_kernel<1, 2, 4, true><<<grid, 32, 0, stream>>>(g);
    else roialign_bwd_gather_kernel<1, 2, 4, false><<<grid, 32, 0, stream>>>(g);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

// One or two heads (the same boxes pooled at pools[h] x pools[h], upstream gradients grads[h]) into one gradient pyramid.
static int launch_bwd_gather(int heads, const float* const grads[2], const int pools[2], const int H[4], const int W[4], int B,
                             int C, const float* boxes, const int32_t* box_index, int N, float image_area, float* const gfm[4],
                             int accumulate, void* workspace, cudaStream_t stream, int grads_layout = MRCNN_NHWC,
                             int gfm_layout = MRCNN_NHWC) {
    long long bins = 0;
    for (int h = 0; h < heads; ++h) bins += (long long)pools[h] * pools[h];
    const GatherWorkspace ws = carve_gather(workspace, H, W, B, N, bins);
    const GatherParams g = gather_params(ws, H, W, B, C, N, image_area);
    const float* gr[2] = {grads[0], grads[1]};
    if (grads_layout == MRCNN_NCHW) {
        // NCHW upstream gradients: transposed to [N][bins][C] in the tail of the workspace, one head after the other
        float* t = reinterpret_cast<float*>(static_cast<char*>(workspace) + ws.bytes);
        for (int h = 0; h < heads; ++h) {
            const int P2 = pools[h] * pools[h];
            const size_t smem = sizeof(float) * kChunk * (size_t)(P2 | 1);
            if (smem > 48 * 1024)
                MRCNN_CUDA(cudaFuncSetAttribute(grads_nchw_to_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            grads_nchw_to_nhwc_kernel<<<dim3(N, (C + kChunk - 1) / kChunk), kThreads, smem, stream>>>(grads[h], t, C, P2);
            MRCNN_LAUNCH_CHECK();
            gr[h] = t;
            t += (size_t)N * C * P2;
        }
        if (heads == 1) gr[1] = gr[0];
    }
    if (int rc = launch_gather_plan(g, ws, workspace, heads, pools, boxes, box_index, stream)) return rc;
    return launch_gather_run(g, heads, gr, gfm, accumulate, stream, gfm_layout == MRCNN_NCHW ? 1 : 0);
}

static int check_layout(int v, const char* what) {
    if (v != MRCNN_NCHW && v != MRCNN_NHWC) return fail(MRCNN_E_INVALID_ARG, "%s must be MRCNN_NCHW or MRCNN_NHWC", what);
    return MRCNN_OK;
}

}  // namespace mrcnn

using namespace mrcnn;

extern "C" {

size_t mrcnn_pyramid_roi_align_backward_workspace_bytes_ex(const int H[4], const int W[4], int B, int C, int N, int pool, int grads_layout) {
    if (!H || !W || B <= 0 || C <= 0 || N < 0 || pool <= 0) return 256;
    for (int l = 0; l < 4; ++l)
        if (H[l] <= 0 || W[l] <= 0) return 256;
    size_t b = carve_gather(nullptr, H, W, B, N, (long long)pool * pool).bytes;
    if (grads_layout == MRCNN_NCHW) b += align_up((size_t)(N > 0 ? N : 1) * C * (size_t)pool * pool * sizeof(float), 256);
    return b;
}

size_t mrcnn_pyramid_roi_align_backward_workspace_bytes(const int H[4], const int W[4], int B, int N, int pool) {
    if (!H || !W || B <= 0 || N < 0 || pool <= 0) return 256;
    for (int l = 0; l < 4; ++l)
        if (H[l] <= 0 || W[l] <= 0) return 256;
    return carve_gather(nullptr, H, W, B, N, (long long)pool * pool).bytes;
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

int mrcnn_pyramid_roi_align_forward_pair(const float* const fm[4], const int H[4], const int W[4], int B, int C,
                                         const float* boxes, const int32_t* box_index, int N, float image_area, float* out7,
                                         float* out14, mrcnn_stream_t stream) {
    MRCNN_REQUIRE(fm && H && W, "mrcnn_pyramid_roi_align_forward_pair: null level tables");
    MRCNN_REQUIRE(B > 0 && C > 0 && N >= 0 && image_area > 0.f, "mrcnn_pyramid_roi_align_forward_pair: bad sizes");
    MRCNN_REQUIRE((C % 4) == 0, "mrcnn_pyramid_roi_align_forward_pair: channels-last tensors with C %% 4 == 0");
    if (N == 0) return MRCNN_OK;
    RoiParams p = {};
    for (int l = 0; l < 4; ++l) {
        MRCNN_REQUIRE(H[l] > 0 && W[l] > 0, "mrcnn_pyramid_roi_align_forward_pair: level %d has empty shape", l);
        MRCNN_REQUIRE_DEV(fm[l]);
        MRCNN_REQUIRE(aligned16(fm[l]), "mrcnn_pyramid_roi_align_forward_pair: level %d is not 16-byte aligned", l);
        MRCNN_REQUIRE((long long)H[l] * W[l] * C < (1ll << 31), "mrcnn_pyramid_roi_align_forward_pair: level %d too large for 32-bit offsets", l);
        p.lv[l] = {const_cast<float*>(fm[l]), H[l], W[l]};
    }
    MRCNN_REQUIRE_DEV(boxes);
    MRCNN_REQUIRE_DEV(out7);
    MRCNN_REQUIRE_DEV(out14);
    if (box_index) MRCNN_REQUIRE_DEV(box_index);
    MRCNN_REQUIRE(aligned16(out7) && aligned16(out14), "mrcnn_pyramid_roi_align_forward_pair: outputs must be 16-byte aligned");
    const long long grid = (long long)N * ((C + kChunk - 1) / kChunk);
    MRCNN_REQUIRE(grid < (1ll << 31) && (long long)196 * C < (1ll << 31), "mrcnn_pyramid_roi_align_forward_pair: too many RoIs");
    p.pyramid = 1;
    p.rule = make_level_rule(image_area);
    p.B = B; p.C = C;
    p.boxes = boxes; p.box_index = box_index; p.N = N;
    p.ph = 14; p.pw = 14; p.extrap = 0.f;
    p.crops = out14;
    p.crops_b = out7;
    p.negzero = -0.0f;
    p.err = device_error_word();
    MRCNN_REQUIRE(p.err != nullptr, "cannot allocate device error word");
    roialign_fwd_nhwc_pair_kernel<<<(unsigned)grid, kThreads, 0, (cudaStream_t)stream>>>(p);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
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
                            gather_eligible(H, W, B, C, N, (long long)pool * pool, pool, gfm_layout, grads_layout, grads, gfm, workspace, workspace_bytes);
    if (algo == MRCNN_BWD_GATHER)
        MRCNN_REQUIRE(can_gather,
                      "mrcnn_pyramid_roi_align_backward: MRCNN_BWD_GATHER needs C %% 4 == 0, N > 0, N * pool^2 * C < 2^31, no "
                      "image_offsets_host and a 256-byte aligned workspace of mrcnn_pyramid_roi_align_backward_workspace_bytes_ex()");
    // With an NCHW gradient pyramid or NCHW upstream gradients the gather pays a transposed copy of the gradients and
    // sector-scattered pyramid writes: measured on B200 at the configs[3] geometry it wins for the 7x7 head (0.81 ms against 1.10 ms
    // for clear + scatter) and loses for the 14x14 head (1.64 against 1.44).  MRCNN_BWD_AUTO takes it when the upstream gradients
    // are smaller than the pyramid they scatter into.
    bool prefer_gather = true;
    if (gfm_layout == MRCNN_NCHW || grads_layout == MRCNN_NCHW) {
        double pyr = 0;
        for (int l = 0; l < 4; ++l) pyr += (double)per_image[l] * B;
        prefer_gather = (double)N * C * pool * pool < pyr;
    }
    if (can_gather && algo != MRCNN_BWD_SCATTER && (prefer_gather || algo == MRCNN_BWD_GATHER)) {
        // tile-owner gather: writes every pixel once (zero fill included), no atomics
        const float* const gr[2] = {grads, grads};
        const int pools[2] = {pool, pool};
        return launch_bwd_gather(1, gr, pools, H, W, B, C, boxes, box_index, N, image_area, gfm, zero_fill ? 0 : 1, workspace,
                                 stream, grads_layout, gfm_layout);
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

int mrcnn_pyramid_roi_align_backward_plan(const int H[4], const int W[4], int B, int C, const float* boxes,
                                          const int32_t* box_index, int N, int pool, float image_area, void* workspace,
                                          size_t workspace_bytes, mrcnn_stream_t stream) {
    MRCNN_REQUIRE(H && W, "mrcnn_pyramid_roi_align_backward_plan: null level tables");
    MRCNN_REQUIRE(B > 0 && C > 0 && N > 0 && pool > 0 && image_area > 0.f, "mrcnn_pyramid_roi_align_backward_plan: bad sizes");
    for (int l = 0; l < 4; ++l) MRCNN_REQUIRE(H[l] > 0 && W[l] > 0, "mrcnn_pyramid_roi_align_backward_plan: level %d has empty shape", l);
    MRCNN_REQUIRE_DEV(boxes);
    if (box_index) MRCNN_REQUIRE_DEV(box_index);
    MRCNN_REQUIRE_DEV(workspace);
    const long long bins = (long long)pool * pool;
    float* const none[4] = {nullptr, nullptr, nullptr, nullptr};
    MRCNN_REQUIRE(gather_eligible(H, W, B, C, N, bins, pool, MRCNN_NHWC, MRCNN_NHWC, nullptr, none, workspace, workspace_bytes),
                  "mrcnn_pyramid_roi_align_backward_plan: needs C %% 4 == 0, N * pool^2 * C < 2^31 and a 256-byte aligned workspace "
                  "of mrcnn_pyramid_roi_align_backward_workspace_bytes()");
    MRCNN_REQUIRE(device_error_word() != nullptr, "cannot allocate device error word");
    const GatherWorkspace ws = carve_gather(workspace, H, W, B, N, bins);
    const int pools[2] = {pool, pool};
    return launch_gather_plan(gather_params(ws, H, W, B, C, N, image_area), ws, workspace, 1, pools, boxes, box_index,
                              (cudaStream_t)stream);
}

int mrcnn_pyramid_roi_align_backward_planned(const float* grads, const int H[4], const int W[4], int B, int C, int N, int pool,
                                             float* const gfm[4], int zero_fill, const void* workspace, size_t workspace_bytes,
                                             mrcnn_stream_t stream) {
    MRCNN_REQUIRE(gfm && H && W, "mrcnn_pyramid_roi_align_backward_planned: null level tables");
    MRCNN_REQUIRE(B > 0 && C > 0 && N > 0 && pool > 0, "mrcnn_pyramid_roi_align_backward_planned: bad sizes");
    for (int l = 0; l < 4; ++l) {
        MRCNN_REQUIRE(H[l] > 0 && W[l] > 0, "mrcnn_pyramid_roi_align_backward_planned: level %d has empty shape", l);
        MRCNN_REQUIRE_DEV(gfm[l]);
    }
    MRCNN_REQUIRE_DEV(grads);
    MRCNN_REQUIRE_DEV(workspace);
    const long long bins = (long long)pool * pool;
    MRCNN_REQUIRE(gather_eligible(H, W, B, C, N, bins, pool, MRCNN_NHWC, MRCNN_NHWC, grads, gfm, workspace, workspace_bytes),
                  "mrcnn_pyramid_roi_align_backward_planned: needs channels-last grads and gfm, C %% 4 == 0, N * pool^2 * C < 2^31 "
                  "and the workspace mrcnn_pyramid_roi_align_backward_plan() filled");
    const GatherWorkspace ws = carve_gather(const_cast<void*>(workspace), H, W, B, N, bins);
    const float* const gr[2] = {grads, grads};
    return launch_gather_run(gather_params(ws, H, W, B, C, N, 1.0f), 1, gr, gfm, zero_fill ? 0 : 1, (cudaStream_t)stream);
}

size_t mrcnn_pyramid_roi_align_backward_pair_workspace_bytes(const int H[4], const int W[4], int B, int N, int pool_a,
                                                             int pool_b) {
    if (!H || !W || B <= 0 || N < 0 || pool_a <= 0 || pool_b <= 0) return 256;
    for (int l = 0; l < 4; ++l)
        if (H[l] <= 0 || W[l] <= 0) return 256;
    return carve_gather(nullptr, H, W, B, N, (long long)pool_a * pool_a + (long long)pool_b * pool_b).bytes;
}

int mrcnn_pyramid_roi_align_backward_pair(const float* grads_a, int pool_a, const float* grads_b, int pool_b, const int H[4],
                                          const int W[4], int B, int C, const float* boxes, const int32_t* box_index, int N,
                                          float image_area, float* const gfm[4], int zero_fill, void* workspace,
                                          size_t workspace_bytes, mrcnn_stream_t stream) {
    MRCNN_REQUIRE(gfm && H && W, "mrcnn_pyramid_roi_align_backward_pair: null level tables");
    MRCNN_REQUIRE(B > 0 && C > 0 && N > 0 && pool_a > 0 && pool_b > 0 && image_area > 0.f,
                  "mrcnn_pyramid_roi_align_backward_pair: bad sizes");
    for (int l = 0; l < 4; ++l) {
        MRCNN_REQUIRE(H[l] > 0 && W[l] > 0, "mrcnn_pyramid_roi_align_backward_pair: level %d has empty shape", l);
        MRCNN_REQUIRE_DEV(gfm[l]);
    }
    MRCNN_REQUIRE_DEV(grads_a);
    MRCNN_REQUIRE_DEV(grads_b);
    MRCNN_REQUIRE_DEV(boxes);
    if (box_index) MRCNN_REQUIRE_DEV(box_index);
    const long long bins = (long long)pool_a * pool_a + (long long)pool_b * pool_b;
    const int max_pool = pool_a > pool_b ? pool_a : pool_b;
    MRCNN_REQUIRE(gather_eligible(H, W, B, C, N, bins, max_pool, MRCNN_NHWC, MRCNN_NHWC, grads_a, gfm, workspace, workspace_bytes) &&
                      aligned16(grads_b),
                  "mrcnn_pyramid_roi_align_backward_pair: needs channels-last tensors, C %% 4 == 0, N * pool^2 * C < 2^31 and a "
                  "256-byte aligned workspace of mrcnn_pyramid_roi_align_backward_pair_workspace_bytes()");
    MRCNN_REQUIRE(device_error_word() != nullptr, "cannot allocate device error word");
    const float* const gr[2] = {grads_a, grads_b};
    const int pools[2] = {pool_a, pool_b};
    return launch_bwd_gather(2, gr, pools, H, W, B, C, boxes, box_index, N, image_area, gfm, zero_fill ? 0 : 1, workspace,
                             (cudaStream_t)stream);
}

int mrcnn_set_deterministic(int on) {
    g_deterministic = on ? 1 : 0;
    return MRCNN_OK;
}

}  // extern "C"
