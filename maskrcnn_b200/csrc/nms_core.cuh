// nms_core.cuh — building blocks shared by nms.cu, proposal.cu and detection.cu:
//   * block-wide bitonic sort of 64-bit composite keys (score key << 32 | ~index) in shared memory,
//   * the 64x64 IoU suppression-word tile (upper triangle only),
//   * the sequential greedy sweep over suppression words, run by ONE CTA entirely on the device
//     (the reference copies the N x N/64 mask to the host and sweeps on the CPU, nms_cuda.cu:107-131).
//     Row blocks of the mask are streamed into shared memory with 1-D bulk async copies (TMA,
//     cp.async.bulk + mbarrier), double-buffered, so the sweep never waits on L2 latency.
#pragma once
#include "common.cuh"

namespace mrcnn {

// ---------------------------------------------------------------------------------------------
// composite sort key: larger score first, ties -> smaller index first (a stable descending sort)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_sort_key(float score, uint32_t idx) {
    return ((uint64_t)float_to_key(score) << 32) | (uint64_t)(0xffffffffu - idx);
}
__device__ __forceinline__ uint32_t sort_key_index(uint64_t k) { return 0xffffffffu - (uint32_t)(k & 0xffffffffu); }
__device__ __forceinline__ float sort_key_score(uint64_t k) { return key_to_float((uint32_t)(k >> 32)); }

// Sorts s[0..P) (P a power of two) descending.  `gbase` is the global index of s[0] when the tile is
// part of a larger bitonic network (direction bit taken from the global index); stages k in
// [k_first, k_last], and for k == k_first only j <= j_first (used by the multi-tile merge).
__device__ __forceinline__ void block_bitonic_desc(uint64_t* s, int P, unsigned gbase, unsigned k_first,
                                                   unsigned j_first, unsigned k_last) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (unsigned k = k_first; k <= k_last; k <<= 1) {
        for (unsigned j = (k == k_first) ? j_first : (k >> 1); j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += nt) {
                const unsigned i = (((unsigned)t & ~(j - 1)) << 1) | ((unsigned)t & (j - 1));
                const unsigned l = i | j;
                const bool desc = (((gbase + i) & k) == 0);
                const uint64_t a = s[i], b = s[l];
                if ((a < b) == desc && a != b) {
                    s[i] = b;
                    s[l] = a;
                }
            }
            __syncthreads();
        }
    }
}

// 1024 keys, 1024 threads, ONE key per thread kept in a register: a stage of the network with stride j exchanges with
// lane ^ j by shuffle when j < 32 (no barrier: 40 of the 55 stages) and through a double-buffered shared-memory array
// otherwise (one __syncthreads per stage).  Fully unrolled, so k and j are constants and a shuffle stage is ~10
// instructions: with 32 warps on one SM the sort is issue-bound (the rolled shared-memory form above costs ~380 cycles per
// stage, this one ~90).  Returns the key of descending rank threadIdx.x.  buf: 2 x 1024 keys of shared memory.
__device__ __forceinline__ uint64_t block_bitonic_desc_1024_reg(uint64_t v, uint64_t* buf) {
    const unsigned tid = threadIdx.x;
    int b = 0;
#pragma unroll
    for (int lk = 1; lk <= 10; ++lk) {
        const unsigned k = 1u << lk;
        const bool desc = (tid & k) == 0u;
#pragma unroll
        for (int lj = lk - 1; lj >= 0; --lj) {
            const unsigned j = 1u << lj;
            uint64_t pv;
            if (j >= 32u) {
                buf[b * 1024 + tid] = v;
                __syncthreads();
                pv = buf[b * 1024 + (tid ^ j)];
                b ^= 1;
            } else {
                pv = __shfl_xor_sync(0xffffffffu, v, (int)j);
            }
            const bool want_max = (((tid & j) == 0u) == desc);
            v = want_max ? (v > pv ? v : pv) : (v < pv ? v : pv);
        }
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// suppression word: bit c set iff box (col0 + c) is suppressed by `rowbox` (IoU >= thr) and comes
// later in the order (col index > row index).  Column boxes/areas are read from shared memory.
// `same_class` optionally restricts suppression to equal class ids (detection layer).
// ---------------------------------------------------------------------------------------------
template <bool kClassAware>
__device__ __forceinline__ uint64_t suppression_word(const float4 rowbox, float rowarea, int rowcls, int row,
                                                     const float4* cbox, const float* carea, const int* ccls,
                                                     int col0, int ncols, float thr) {
    uint64_t w = 0;
    const int start = (row >= col0) ? (row - col0 + 1) : 0;
    const float margin = __fadd_rn(__fmul_rn(fabsf(thr), 1e-6f), 1e-37f);
    if (kClassAware) {
        for (int c = start; c < ncols; ++c) {
            bool s = iou_ge(rowbox, rowarea, cbox[c], carea[c], thr);
            s = s && (ccls[c] == rowcls);
            if (s) w |= (1ull << c);
        }
        return w;
    }
    if (start == 0 && ncols == 64) {  // a full off-diagonal tile: constant bit positions, smem offsets folded into the loads
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            if (iou_ge_m(rowbox, rowarea, cbox[c], carea[c], thr, margin)) lo |= (1u << c);
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            if (iou_ge_m(rowbox, rowarea, cbox[32 + c], carea[32 + c], thr, margin)) hi |= (1u << c);
        }
        return ((uint64_t)hi << 32) | lo;
    }
    for (int c = start; c < ncols; ++c)
        if (iou_ge_m(rowbox, rowarea, cbox[c], carea[c], thr, margin)) w |= (1ull << c);
    return w;
}

// ---------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (sm_90+; SASS: SYNCS.*, UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Greedy sweep.  mask: [rows padded to 64][W] suppression words in GLOBAL memory (only words at or
// right of the diagonal are meaningful).  n boxes in score order, W = ceil(n/64).
// Shared memory supplied by the caller:
//   remv  [W] words   — suppressed-so-far bitmap (working state)
//   kept  [W] words   — OUT: bit set = box survives
//   stage 2 * 64 * W words (16 B aligned) if `staged`, else unused
//   bars  2 mbarriers
// Stops early once max_keep survivors are found (later words of kept[] are zero).
// Returns the number of survivors found (>= max_keep possible by up to 63).  All threads must call.
// ---------------------------------------------------------------------------------------------
struct SweepSmem {
    uint64_t* remv;
    uint64_t* kept;
    uint64_t* stage;
    uint64_t* bars;
    int* total;  // one int
};

// diag_t (optional): [rows] words, bit j of diag_t[i] set iff box j of i's chunk (j < i within the chunk) suppresses box i - the
// diagonal tiles transposed, written by the mask kernel; without it warp 0 transposes each diagonal tile itself (64 broadcast loads).
__device__ __forceinline__ int block_nms_sweep(const uint64_t* mask, int n, int W, const SweepSmem sm,
                                               bool staged, int max_keep, const uint64_t* diag_t = nullptr) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int w = tid; w < W; w += nt) {
        sm.remv[w] = 0;
        sm.kept[w] = 0;
    }
    const uint32_t block_bytes = (uint32_t)(64u * (unsigned)W * 8u);
    if (tid == 0) {
        *sm.total = 0;
        if (staged) {
            mbar_init(&sm.bars[0], 1);
            mbar_init(&sm.bars[1], 1);
            fence_barrier_init();
        }
    }
    __syncthreads();
    if (staged && tid == 0 && W > 0) {
        mbar_expect_tx(&sm.bars[0], block_bytes);
        bulk_g2s(sm.stage, mask, block_bytes, &sm.bars[0]);
    }
    int total = 0;
    // warp 0 keeps the transposed diagonal words of the NEXT chunk in registers: the two global loads are off the critical path
    uint64_t pa = 0ull, pb = 0ull;
    if (diag_t != nullptr && tid < 32 && W > 0) {
        if (tid < n) pa = __ldg(diag_t + tid);
        if (tid + 32 < n) pb = __ldg(diag_t + tid + 32);
    }
    // threads per word in the propagate step: a power of two <= 16, fixed for the sweep (W - 1 words at most)
    int lp = 0;
    while (lp < 4 && (2 << lp) * max(W - 1, 1) <= nt) ++lp;
    const int parts = 1 << lp, span = 64 >> lp;
    for (int c = 0; c < W; ++c) {
        const uint64_t* rows;  // 64 rows x W words of chunk c
        if (staged) {
            const int buf = c & 1;
            if (tid == 0 && c + 1 < W) {  // prefetch the next row block into the other buffer
                fence_proxy_async();
                mbar_expect_tx(&sm.bars[buf ^ 1], block_bytes);
                bulk_g2s(sm.stage + (size_t)(buf ^ 1) * 64 * W, mask + (size_t)(c + 1) * 64 * W, block_bytes,
                         &sm.bars[buf ^ 1]);
            }
            mbar_wait(&sm.bars[buf], (uint32_t)((c >> 1) & 1));
            MRCNN_DBG(buf == 0 || buf == 1);
            MRCNN_DBG((size_t)block_bytes == (size_t)64 * W * 8 && (block_bytes & 15u) == 0);
            rows = sm.stage + (size_t)buf * 64 * W;
        } else {
            rows = mask + (size_t)c * 64 * W;
        }
        // --- resolve the 64 boxes of this chunk against each other (warp 0; lane l owns boxes l and l + 32).  Box l survives
        //     iff it is a candidate and no SURVIVING earlier box of the chunk suppresses it: a triangular system, solved by
        //     fixed-point iteration over the whole chunk at once (every pass fixes at least the next box in order; real inputs
        //     settle in a handful of passes).  The first version walked the survivors one at a time with two 64-bit shuffles
        //     per survivor - ~130 dependent cycles each, 565 us for 6000 boxes with 3545 survivors. ---
        if (tid < 32) {
            const int nrows = min(64, n - c * 64);
            uint64_t ca = 0ull, cb = 0ull;   // earlier boxes of the chunk that suppress box tid / box tid + 32 (column view)
            if (diag_t != nullptr) {
                ca = pa;
                cb = pb;
                pa = pb = 0ull;
                const int nn = n - (c + 1) * 64;   // rows of the next chunk
                if (tid < nn) pa = __ldg(diag_t + (size_t)(c + 1) * 64 + tid);
                if (tid + 32 < nn) pb = __ldg(diag_t + (size_t)(c + 1) * 64 + tid + 32);
            } else {
                for (int j = 0; j < nrows; ++j) {
                    const uint64_t d = rows[(size_t)j * W + c];   // row j of the diagonal tile (bits above j only), same word for all lanes
                    ca |= ((d >> tid) & 1ull) << j;
                    cb |= ((d >> (tid + 32)) & 1ull) << j;
                }
            }
            uint64_t cand = ~sm.remv[c];
            if (nrows < 64) cand &= ((1ull << nrows) - 1ull);
            uint64_t alive = cand;
            for (;;) {  // warp-uniform: every lane holds the same alive
                const unsigned sa = __ballot_sync(0xffffffffu, (ca & alive) != 0ull);
                const unsigned sb = __ballot_sync(0xffffffffu, (cb & alive) != 0ull);
                const uint64_t next = cand & ~((uint64_t)sa | ((uint64_t)sb << 32));
                if (next == alive) break;
                alive = next;
            }
            MRCNN_DBG((alive & ~cand) == 0ull);
            if (tid == 0) {
                sm.kept[c] = alive;
                *sm.total += __popcll(alive);
            }
        }
        __syncthreads();
        const uint64_t kept = sm.kept[c];
        total = *sm.total;
        if (total >= max_keep) {
            // never leave a bulk copy in flight behind us: drain the prefetch of chunk c + 1
            if (staged && c + 1 < W) mbar_wait(&sm.bars[(c + 1) & 1], (uint32_t)(((c + 1) >> 1) & 1));
            break;
        }
        // --- propagate: every later word ORs in the rows of the survivors of this chunk: `parts` threads per word, each owning a
        //     contiguous slice of the chunk's 64 rows (four independent loads per round); partial ORs meet in shared memory ---
        {
            const int later = W - (c + 1);
            if (later * parts <= nt) {
                const int part = tid & (parts - 1), wi = tid >> lp;
                if (wi < later) {
                    const int w = c + 1 + wi;
                    MRCNN_DBG(w > c && w < W);
                    const uint64_t k = (kept >> (part * span)) & (span == 64 ? ~0ull : ((1ull << span) - 1ull));
                    const uint64_t* base = rows + (size_t)(part * span) * W + w;
                    uint64_t acc = 0;
                    for (int r0 = 0; r0 < span && (k >> r0) != 0ull; r0 += 4) {
                        const unsigned m4 = (unsigned)(k >> r0) & 15u;
                        MRCNN_DBG(part * span + r0 + 3 < 64);
                        const uint64_t v0 = (m4 & 1u) ? base[(size_t)(r0 + 0) * W] : 0ull;
                        const uint64_t v1 = (m4 & 2u) ? base[(size_t)(r0 + 1) * W] : 0ull;
                        const uint64_t v2 = (m4 & 4u) ? base[(size_t)(r0 + 2) * W] : 0ull;
                        const uint64_t v3 = (m4 & 8u) ? base[(size_t)(r0 + 3) * W] : 0ull;
                        acc |= (v0 | v1) | (v2 | v3);
                    }
                    if (acc) {
                        if (parts == 1) sm.remv[w] |= acc;
                        else atomicOr(reinterpret_cast<unsigned long long*>(&sm.remv[w]), (unsigned long long)acc);
                    }
                }
            } else {
                for (int w = c + 1 + tid; w < W; w += nt) {   // more words than threads: one thread per word, all survivors
                    uint64_t acc = 0;
                    uint64_t k = kept;
                    while (k) {
                        const int r = __ffsll((long long)k) - 1;
                        k &= k - 1;
                        acc |= rows[(size_t)r * W + w];
                    }
                    sm.remv[w] |= acc;
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    return total;
}

// ---------------------------------------------------------------------------------------------
// IoU >= thr, "certainly not" decided without the division: for positive areas and thr > 0,
//   inter / (S - inter) >= thr  <=>  inter >= thr / (1 + thr) S,   S = area_a + area_b,
// and the right side splits into a share per box, so a pair costs one add and one compare after the intersection.  A pair below
// (1 - e) of that bound, e = 4e-6, is below the threshold in the reference's own expression too: e is ~8 times every rounding
// on either side put together (the few fp32 operations here; the sum, the difference and the division there: ~2^-22 relative).
// Everything else - the real overlaps, the borderline pairs, and non-positive or non-finite areas or thr <= 0, whose shares
// are NaN so that the compare fails - goes through iou_ge_m, the reference's expression, in a second loop that is nearly
// always short (an overlap above the threshold is one pair in thousands).
__device__ __forceinline__ float no_share(float area, float thr) {
    const bool ok = area > 0.0f && area < 1e37f && thr > 0.0f && thr < 1e6f;
    return ok ? __fmul_rn(__fdiv_rn(__fmul_rn(thr, 0.999996f), __fadd_rn(1.0f, thr)), area) : __int_as_float(0x7fc00000);
}
// false: certainly below the threshold.  One clamp is enough: with w clamped at 0 a negative h gives a product <= 0.
__device__ __forceinline__ bool iou_maybe(const float4 a, float share_a, const float4 b, float share_b) {
    const float yy1 = fmaxf(a.x, b.x);
    const float xx1 = fmaxf(a.y, b.y);
    const float yy2 = fminf(a.z, b.z);
    const float xx2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
    const float h = __fadd_rn(__fsub_rn(yy2, yy1), 1.0f);
    return !(__fmul_rn(w, h) < __fadd_rn(share_a, share_b));
}


// ---------------------------------------------------------------------------------------------
// Lower-triangle suppression tiles ("who suppresses me"), tile-major: tile q = rb (rb + 1) / 2 + cb (cb <= rb) holds, for each of
// the 64 row boxes of block rb, the EARLIER boxes of column block cb with IoU >= thr.  Shared by nms.cu and proposal.cu.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lower_tile_coords(int q, int& rb, int& cb) {
    rb = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
    while ((rb + 1) * (rb + 2) / 2 <= q) ++rb;
    while (rb * (rb + 1) / 2 > q) --rb;
    cb = q - rb * (rb + 1) / 2;
}

// The word of row box `row` = rb * 64 + t (box b, area a; row < n) of tile (rb, cb); the 64 column boxes of block cb are staged in
// shared memory (cbox / carea / cshare = no_share of the column).  Off the diagonal every column is earlier and exists.
__device__ __forceinline__ uint64_t lower_tile_word(const float4 b, float a, bool diagonal, int t, const float4* cbox, const float* carea,
                                                    const float* cshare, float thr) {
    const float share = no_share(a, thr);
    const float margin = __fadd_rn(__fmul_rn(fabsf(thr), 1e-6f), 1e-37f);
    uint64_t w = 0ull;
    if (!diagonal) {   // constant bit positions
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (iou_maybe(cbox[c], cshare[c], b, share)) lo |= (1u << c);
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (iou_maybe(cbox[32 + c], cshare[32 + c], b, share)) hi |= (1u << c);
        uint64_t maybe = ((uint64_t)hi << 32) | lo;
        while (maybe) {   // the reference's expression decides
            const int c = __ffsll((long long)maybe) - 1;
            maybe &= maybe - 1;
            if (iou_ge_m(cbox[c], carea[c], b, a, thr, margin)) w |= (1ull << c);
        }
    } else {
        for (int c = 0; c < t; ++c)
            if (iou_ge_m(cbox[c], carea[c], b, a, thr, margin)) w |= (1ull << c);
    }
    return w;
}

// .gpu-scope accesses for values exchanged between CTAs of one grid (they bypass L1)
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Largest W for which the double-buffered staging (2 * 64 * W * 8 bytes) fits next to the rest.
constexpr int kSweepStageMaxW = 160;  // 160 KB of staging

}  // namespace mrcnn
