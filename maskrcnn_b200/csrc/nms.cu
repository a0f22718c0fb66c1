// nms.cu — greedy NMS for sm_100a, fully on the device.
//
// Replaces c++ext/maskrcnn/csrc/nms.h:15-30 -> cpu/nms_cpu.cpp:11-70 (and cuda/nms_cuda.cu:29-137).
// Semantics follow the CPU implementation, which is the parity oracle: order = scores descending,
// suppress iff IoU >= threshold with the (+1) pixel convention, result = ASCENDING original indices.
//
// Pipeline (all on `stream`, no host round trip — the reference's CUDA path copies the whole mask to
// the host and sweeps there):
//   1. prepare : composite keys (score, index) -> tile sort + rank merge (bitonic above 8192 keys) -> boxes/areas gathered in
//                score order
//   2 + 3, from 8 chunks of 64 boxes on (nms_mask_lower_kernel, nms_fixpoint_pub_kernel / nms_fixpoint_kernel):
//                lower-triangle tile-major mask ("who suppresses me") and a GRID-WIDE fixed-point iteration - every chunk owned by
//                a CTA, all chunks re-evaluated at once from the previous pass's survivor words until a pass changes nothing;
//                the fixed point is the greedy answer.  Cooperative launch; CTA 0 emits the ascending list.
//   2 + 3, below that (nms_mask_kernel, nms_sweep_kernel): upper-triangle mask, one CTA sweeps it chunk by chunk (row blocks
//                streamed through shared memory by bulk async copies), then an in-CTA prefix scan emits the list.
#include <algorithm>
#include <limits.h>
#include <stdlib.h>

#include "api_util.h"
#include "nms_core.cuh"

namespace mrcnn {

constexpr int kSortTile = 8192;  // elements sorted per CTA in shared memory (64 KB)
constexpr int kFixpointMinW = 8;  // chunks of 64 boxes from which the grid-wide fixed-point route is taken

struct NmsWorkspace {
    uint64_t* sortbuf;  // [P]
    float4* sbox;       // [N64] boxes in score order
    float* sarea;       // [N64]
    int32_t* order;     // [N64] original index of the i-th best box
    uint64_t* mask;     // [N64][W]
    uint8_t* flags;     // [N64] survivor flag per ORIGINAL index
    uint64_t* diag_t;   // [N64] diagonal tiles transposed: which earlier boxes of its chunk suppress box i
    // fixed-point path (nms_fixpoint_kernel): `mask` then holds the LOWER triangle, tile-major
    uint64_t* keepw;    // [W] survivor words in score order (the iterate)
    uint32_t* obits;    // [N64 / 32] survivor bits by ORIGINAL index
    unsigned long long* gbar;  // grid barrier word: arrivals | changes of odd passes << 20 | changes of even passes << 42
    uint64_t* pub;      // [2][W][2] published survivor half-words, each with its pass tag (nms_fixpoint_pub_kernel)
    size_t bytes;
};

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

static NmsWorkspace carve_nms(void* base, int N) {
    NmsWorkspace w;
    const size_t N64 = align_up((size_t)(N > 0 ? N : 1), 64);
    const size_t W = N64 / 64;
    const size_t P = (size_t)next_pow2(N > 0 ? N : 1);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? (void*)((char*)base + off) : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    w.sortbuf = (uint64_t*)take(P * 8);
    w.sbox = (float4*)take(N64 * 16);
    w.sarea = (float*)take(N64 * 4);
    w.order = (int32_t*)take(N64 * 4);
    w.mask = (uint64_t*)take(N64 * W * 8);
    w.flags = (uint8_t*)take(N64);
    w.diag_t = (uint64_t*)take(N64 * 8);
    w.keepw = (uint64_t*)take(W * 8);
    w.obits = (uint32_t*)take(N64 / 8);
    w.gbar = (unsigned long long*)take(8);
    w.pub = (uint64_t*)take(W * 32);
    w.bytes = off;
    return w;
}

// ---- 1. prepare ---------------------------------------------------------------------------------

__device__ __forceinline__ void gather_sorted(const float* dets, uint64_t key, int i, float4* sbox, float* sarea,
                                              int32_t* order) {
    const uint32_t src = sort_key_index(key);
    const float* d = dets + (size_t)src * 5;
    const float4 b = make_float4(__ldg(d), __ldg(d + 1), __ldg(d + 2), __ldg(d + 3));
    sbox[i] = b;
    sarea[i] = box_area_p1(b);
    order[i] = (int32_t)src;
}

// N <= kSortTile: keys, sort and gather in one CTA.
__global__ void __launch_bounds__(1024) nms_prepare_small_kernel(const float* __restrict__ dets, int N, int P, float4* sbox,
                                                                 float* sarea, int32_t* order) {
    extern __shared__ __align__(16) uint64_t skeys[];
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        skeys[i] = (i < N) ? make_sort_key(__ldg(dets + (size_t)i * 5 + 4), (uint32_t)i) : 0ull;
    __syncthreads();
    block_bitonic_desc(skeys, P, 0u, 2u, 1u, (unsigned)P);
    for (int i = threadIdx.x; i < N; i += blockDim.x) gather_sorted(dets, skeys[i], i, sbox, sarea, order);
}

// 1024 < P <= kSortTile: P / 1024 CTAs sort one 1024-key tile each with the one-key-per-thread register / shuffle network
// (~3 us, all tiles at once), then every element finds its global rank as its rank inside its tile plus, for every other tile, the
// number of keys there that beat it (a binary search over the tile in shared memory; keys are unique: score, then index).  Two short
// launches instead of one CTA grinding through the 91 stages of an 8192-key network (75 us).
__global__ void __launch_bounds__(1024) nms_tile_sort_kernel(const float* __restrict__ dets, int N, uint64_t* __restrict__ tiles) {
    __shared__ uint64_t buf[2 * 1024];
    const int i = blockIdx.x * 1024 + threadIdx.x;
    const uint64_t key = (i < N) ? make_sort_key(__ldg(dets + (size_t)i * 5 + 4), (uint32_t)i) : 0ull;   // 0 = padding: below every real key
    tiles[i] = block_bitonic_desc_1024_reg(key, buf);
}

// kRankSplit CTAs per tile.  The tiles arrive in shared memory by one bulk copy (a per-thread copy loop was 8 dependent round trips
// for 64 KB), and a key's searches in the other tiles advance in lockstep - T - 1 independent shared-memory reads per step instead
// of (T - 1) x 10 dependent ones.
constexpr int kRankSplit = 4;
constexpr int kRankMaxT = kSortTile / 1024;
__global__ void __launch_bounds__(1024 / kRankSplit) nms_rank_gather_kernel(const float* __restrict__ dets, const uint64_t* __restrict__ tiles,
                                                                            int N, int T, float4* sbox, float* sarea, int32_t* order) {
    extern __shared__ __align__(128) uint64_t skeys[];   // all T tiles
    __shared__ uint64_t bar;
    constexpr int kThreads = 1024 / kRankSplit;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_expect_tx(&bar, (uint32_t)T * 8192u);
        for (int t = 0; t < T; ++t) bulk_g2s(skeys + t * 1024, tiles + t * 1024, 8192u, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0u);
    const int mine = blockIdx.x / kRankSplit, in_tile = (blockIdx.x % kRankSplit) * kThreads + threadIdx.x;
    const uint64_t key = skeys[mine * 1024 + in_tile];
    if (key == 0ull) return;   // padding
    // lo[t]: first position of tile t (descending) whose key is below `key` = how many keys of tile t beat it
    int lo[kRankMaxT];
#pragma unroll
    for (int t = 0; t < kRankMaxT; ++t) lo[t] = 0;
#pragma unroll
    for (int half = 512; half >= 1; half >>= 1) {   // branch-free lower bound over 1024 = 2^10 entries, all tiles at once
#pragma unroll
        for (int t = 0; t < kRankMaxT; ++t)
            if (t < T && skeys[t * 1024 + lo[t] + half - 1] > key) lo[t] += half;
    }
    int rank = in_tile;
#pragma unroll
    for (int t = 0; t < kRankMaxT; ++t)
        if (t < T && t != mine) rank += lo[t] + ((lo[t] == 1023 && skeys[t * 1024 + 1023] > key) ? 1 : 0);
    MRCNN_DBG(rank >= 0 && rank < N && (int)sort_key_index(key) < N);
    gather_sorted(dets, key, rank, sbox, sarea, order);
}

// N > kSortTile: multi-CTA bitonic network over a global key buffer.
__global__ void nms_fill_keys_kernel(const float* __restrict__ dets, int N, int P, uint64_t* keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) keys[i] = (i < N) ? make_sort_key(__ldg(dets + (size_t)i * 5 + 4), (uint32_t)i) : 0ull;
}

// sorts each tile completely (k = 2..tile), direction taken from the global index
__global__ void __launch_bounds__(1024) bitonic_tile_sort_kernel(uint64_t* keys) {
    extern __shared__ __align__(16) uint64_t skeys[];
    uint64_t* g = keys + (size_t)blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) skeys[i] = g[i];
    __syncthreads();
    block_bitonic_desc(skeys, kSortTile, blockIdx.x * (unsigned)kSortTile, 2u, 1u, (unsigned)kSortTile);
    for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) g[i] = skeys[i];
}

// one compare-exchange stage (j >= tile) of merge size k over the global buffer
__global__ void bitonic_global_step_kernel(uint64_t* keys, unsigned P, unsigned j, unsigned k) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (P >> 1)) return;
    const unsigned i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const unsigned l = i | j;
    const bool desc = ((i & k) == 0);
    const uint64_t a = keys[i], b = keys[l];
    if ((a < b) == desc && a != b) {
        keys[i] = b;
        keys[l] = a;
    }
}

// finishes merge size k inside each tile (j = tile/2 .. 1)
__global__ void __launch_bounds__(1024) bitonic_tile_merge_kernel(uint64_t* keys, unsigned k) {
    extern __shared__ __align__(16) uint64_t skeys[];
    uint64_t* g = keys + (size_t)blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) skeys[i] = g[i];
    __syncthreads();
    block_bitonic_desc(skeys, kSortTile, blockIdx.x * (unsigned)kSortTile, k, (unsigned)kSortTile >> 1, k);
    for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) g[i] = skeys[i];
}

__global__ void nms_gather_kernel(const float* __restrict__ dets, const uint64_t* __restrict__ keys, int N, float4* sbox,
                                  float* sarea, int32_t* order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) gather_sorted(dets, keys[i], i, sbox, sarea, order);
}

// ---- 2. mask ------------------------------------------------------------------------------------

// Upper triangle only, folded so that every CTA has a tile: row block r has W - r tiles, row block W - 1 - r has r + 1, together
// W + 1.  grid = (W + 1, ceil(W / 2)), block = 64 threads = the 64 row boxes of the tile.  (The W x W grid of round 1 launched
// 8836 CTAs for 6000 boxes and sent half of them home at once.)
__global__ void __launch_bounds__(64) nms_mask_kernel(const float4* __restrict__ sbox, const float* __restrict__ sarea, int N,
                                                      int W, float thr, uint64_t* __restrict__ mask, uint64_t* __restrict__ diag_t) {
    const int r = blockIdx.y, i = blockIdx.x;
    int rb, cb;
    if (i < W - r) {
        rb = r;
        cb = r + i;
    } else {
        rb = W - 1 - r;
        if (rb == r) return;  // odd W: the middle row block was served by the first branch
        cb = rb + (i - (W - r));
    }
    MRCNN_DBG(rb >= 0 && rb < W && cb >= rb && cb < W);
    __shared__ float4 cbox[64];
    __shared__ float carea[64];
    const int t = threadIdx.x;
    const int col0 = cb * 64;
    const int ncols = min(64, N - col0);
    if (t < ncols) {
        cbox[t] = sbox[col0 + t];
        carea[t] = sarea[col0 + t];
    }
    __syncthreads();
    const int row = rb * 64 + t;
    uint64_t w = 0ull;
    if (row < N) {
        w = suppression_word<false>(sbox[row], sarea[row], 0, row, cbox, carea, nullptr, col0, ncols, thr);
        mask[(size_t)row * W + cb] = w;
    }
    if (cb == rb) {  // diagonal tile: also store it transposed (the sweep's chunk resolve wants "who suppresses me")
        __shared__ uint64_t s_w[64];
        s_w[t] = w;
        __syncthreads();
        uint64_t col = 0ull;
#pragma unroll 8
        for (int j = 0; j < 64; ++j) col |= ((s_w[j] >> t) & 1ull) << j;
        diag_t[row] = col;   // rows up to N64 - 1: the workspace holds N64 words
    }
}

// ---- 3. sweep + ascending emit ------------------------------------------------------------------

__global__ void __launch_bounds__(1024) nms_sweep_kernel(const uint64_t* __restrict__ mask, const int32_t* __restrict__ order,
                                                         int N, int W, int staged, uint8_t* __restrict__ flags,
                                                         int64_t* __restrict__ keep_out, int32_t* __restrict__ count_out,
                                                         const uint64_t* __restrict__ diag_t) {
    extern __shared__ __align__(16) uint64_t sm_raw[];
    __shared__ uint64_t bars[2];
    __shared__ int s_total;
    __shared__ int s_warp_sums[32];
    SweepSmem sm;
    sm.stage = sm_raw;  // [2*64*W] when staged
    sm.remv = sm_raw + (staged ? (size_t)2 * 64 * W : 0);
    sm.kept = sm.remv + W;
    sm.bars = bars;
    sm.total = &s_total;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < N; i += nt) flags[i] = 0;
    block_nms_sweep(mask, N, W, sm, staged != 0, INT_MAX, diag_t);
    // survivors (score order) -> flags by original index
    for (int i = tid; i < N; i += nt)
        if ((sm.kept[i >> 6] >> (i & 63)) & 1ull) flags[order[i]] = 1;
    __syncthreads();
    // ascending compaction: thread t owns the contiguous range [t*per, (t+1)*per)
    const int per = (N + nt - 1) / nt;
    const int beg = min(N, tid * per), end = min(N, beg + per);
    int cnt = 0;
    for (int i = beg; i < end; ++i) cnt += flags[i];
    int incl = cnt;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = (lane < (nt >> 5)) ? s_warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        s_warp_sums[lane] = v;  // inclusive over warps
    }
    __syncthreads();
    int pos = incl - cnt + (warp > 0 ? s_warp_sums[warp - 1] : 0);
    for (int i = beg; i < end; ++i)
        if (flags[i]) keep_out[pos++] = (int64_t)i;
    if (tid == nt - 1) *count_out = pos;  // the last thread's end position is the total
}

// ---- 2b + 3b. the parallel route: lower-triangle mask + a grid-wide fixed-point iteration ------------------------------
//
// Greedy NMS is the unique solution of a triangular system: keep[i] = no KEPT earlier box suppresses box i.  The serial sweep
// above walks it chunk by chunk on one SM (~1.7 us per chunk of 64 boxes, all of it barrier latency: 160 us at 6000 boxes).
// Here every chunk is owned by a CTA and all chunks are re-evaluated AT ONCE from the previous pass's survivor words
// (Jacobi across chunks, exact inside a chunk), with a grid barrier per pass, until a pass changes nothing.  Chunk c is final
// after pass c + 1 at the latest (its predecessors are), so W + 1 passes bound the worst case - a chain in which every box only
// overlaps its successor - and real inputs settle in 6-10 passes (tools/sim_nms_fixpoint.py).  The fixed point IS the greedy
// answer, so the result is bit-identical to the sweep.
//
// The pull form wants "who suppresses me": word (rb, cb <= rb) of row box i holds the EARLIER boxes of column block cb with
// IoU >= thr (iou_ge_m is symmetric in its two boxes: fmax / fmin / a + b commute).  Stored tile-major - tile q = rb (rb + 1) / 2 +
// cb, 64 consecutive words - so that the rows of chunk c are one contiguous block [c + 1][64] (the diagonal tile last).


// kMaskTiles tiles per CTA (64 threads each): 4465 two-warp CTAs for 6000 boxes spent a third of the kernel ramping up
// (sm__cycles_active 39.8 k against 22.4 k issue cycles per scheduler).
constexpr int kMaskTiles = 4;
__global__ void __launch_bounds__(64 * kMaskTiles) nms_mask_lower_kernel(const float4* __restrict__ sbox, const float* __restrict__ sarea,
                                                                         int N, int W, float thr, uint64_t* __restrict__ lower,
                                                                         uint64_t* __restrict__ keepw, uint32_t* __restrict__ obits,
                                                                         unsigned long long* __restrict__ gbar, uint64_t* __restrict__ pub) {
    const int g = threadIdx.x >> 6, t = threadIdx.x & 63;
    const int tiles = W * (W + 1) / 2;
    const int q = min(blockIdx.x * kMaskTiles + g, tiles - 1);   // a spare group repeats the last tile (same values, same addresses)
    int rb, cb;
    lower_tile_coords(q, rb, cb);
    MRCNN_DBG(rb >= 0 && rb < W && cb >= 0 && cb <= rb);
    __shared__ float4 cbox_[kMaskTiles][64];
    __shared__ float carea_[kMaskTiles][64];
    __shared__ float cshare_[kMaskTiles][64];   // the column's share of the "certainly not" bound
    float4* cbox = cbox_[g];
    float* carea = carea_[g];
    float* cshare = cshare_[g];
    const int col0 = cb * 64;
    if (col0 + t < N) {
        cbox[t] = sbox[col0 + t];
        const float a = sarea[col0 + t];
        carea[t] = a;
        cshare[t] = no_share(a, thr);
    }
    __syncthreads();
    const int row = rb * 64 + t;
    uint64_t w = 0ull;
    if (row < N) w = lower_tile_word(sbox[row], sarea[row], cb == rb, t, cbox, carea, cshare, thr);
    lower[(size_t)q * 64 + t] = w;
    // the state the fixed-point kernel starts from (a kernel boundary orders it): every box kept, no survivor bits, barrier at 0
    if (cb == rb && t == 0) {
        const int nrows = min(64, N - rb * 64);
        const uint64_t valid = nrows == 64 ? ~0ull : ((1ull << nrows) - 1ull);
        keepw[rb] = valid;
        // pass 0 of the published form: tag = pass * 2 + changed
        pub[2 * rb] = (1ull << 32) | (valid & 0xffffffffull);
        pub[2 * rb + 1] = (1ull << 32) | (valid >> 32);
        pub[2 * W + 2 * rb] = 0ull;
        pub[2 * W + 2 * rb + 1] = 0ull;
    }
    const int owords = W * 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < owords; i += gridDim.x * blockDim.x) obits[i] = 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) *gbar = 0ull;
}

// All CTAs of the (cooperative) grid meet; `changed` is folded into the barrier word, in the field of this pass's parity (a
// CTA that is already one pass ahead adds to the OTHER field, never to the one a straggler is about to read).  Returns the
// barrier word as read after everybody arrived.  Thread 0 arrives for its CTA; the __syncthreads pair orders the CTA's
// global writes before the arrival (fence + relaxed atomic = release) and its later reads after the acquire.
__device__ __forceinline__ unsigned long long grid_meet(unsigned long long* gbar, unsigned long long* s_word, unsigned pass,
                                                        bool changed) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(gbar, 1ull | (changed ? ((pass & 1u) ? (1ull << 20) : (1ull << 42)) : 0ull));
        const unsigned long long target = (unsigned long long)gridDim.x * pass;
        unsigned long long v;
        do {
            v = ld_acquire_u64(gbar);
        } while ((v & 0xfffffull) < target);
        *s_word = v;
    }
    __syncthreads();
    return *s_word;
}

__global__ void __launch_bounds__(1024) nms_fixpoint_kernel(const uint64_t* __restrict__ lower, const int32_t* __restrict__ order, int N,
                                                            int W, int rows_in_smem, uint64_t* __restrict__ keepw,
                                                            uint32_t* __restrict__ obits, unsigned long long* __restrict__ gbar,
                                                            int64_t* __restrict__ keep_out, int32_t* __restrict__ count_out) {
    extern __shared__ __align__(16) uint64_t s_dyn[];
    uint64_t* s_keep = s_dyn;        // [W] the survivor words as read at the start of a pass
    uint64_t* s_rows = s_dyn + W;    // rows_in_smem: the [c][64] words of this CTA's only chunk
    __shared__ unsigned s_sup[2];
    __shared__ unsigned long long s_word;
    __shared__ int s_warp_sums[32];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int S = nt >> 6, box = tid & 63, slice = tid >> 6;
    const int G = gridDim.x;
    if (rows_in_smem) {
        const int c = blockIdx.x;
        const uint64_t* src = lower + (size_t)c * (c + 1) / 2 * 64;
        for (int i = tid; i < c * 64; i += nt) s_rows[i] = __ldg(src + i);
    }
    unsigned seen[2] = {0u, 0u};   // the change counts of odd / even passes at the last look
    unsigned pass = 0;
    // one chunk per CTA: its diagonal words and its own survivor word stay in registers (every global read in the pass loop is
    // an L2 round trip on the critical path: ~1000 cycles each, three of them per pass before this)
    uint64_t own_ca = 0ull, own_cb = 0ull, own_word = 0ull;
    if (rows_in_smem && warp == 0) {
        const int c = blockIdx.x;
        const uint64_t* tiles = lower + (size_t)c * (c + 1) / 2 * 64;
        own_ca = __ldg(tiles + (size_t)c * 64 + lane);
        own_cb = __ldg(tiles + (size_t)c * 64 + lane + 32);
        const int nrows = min(64, N - c * 64);
        own_word = nrows == 64 ? ~0ull : ((1ull << nrows) - 1ull);   // what nms_mask_lower_kernel wrote
    }
    for (;;) {
        ++pass;
        bool changed = false;
        for (int c = blockIdx.x; c < W; c += G) {
            if (tid < 2) s_sup[tid] = 0u;
            // the survivor words of the earlier chunks, all fetched at once (one L2 round trip; read one by one inside the
            // loop below they cost a round trip each: 41 us for the kernel against ~20)
            for (int w = tid; w < c; w += nt) s_keep[w] = ld_relaxed_u64(keepw + w);
            __syncthreads();
            const uint64_t* tiles = lower + (size_t)c * (c + 1) / 2 * 64;
            const uint64_t* rows = rows_in_smem ? s_rows : tiles;
            uint64_t acc = 0ull;
            for (int w = slice; w < c; w += S) {
                MRCNN_DBG(w >= 0 && w < W && w * 64 + box < (c + 1) * 64);
                acc |= rows[w * 64 + box] & s_keep[w];
            }
            const unsigned hit = __ballot_sync(0xffffffffu, acc != 0ull);
            if (lane == 0 && hit) atomicOr(&s_sup[warp & 1], hit);   // warp -> boxes (warp & 1) * 32 ..
            __syncthreads();
            if (warp == 0) {
                const int nrows = min(64, N - c * 64);
                const uint64_t valid = nrows == 64 ? ~0ull : ((1ull << nrows) - 1ull);
                const uint64_t cand = valid & ~((uint64_t)s_sup[0] | ((uint64_t)s_sup[1] << 32));
                const uint64_t ca = rows_in_smem ? own_ca : __ldg(tiles + (size_t)c * 64 + lane);        // earlier boxes of the chunk that suppress box lane
                const uint64_t cb = rows_in_smem ? own_cb : __ldg(tiles + (size_t)c * 64 + lane + 32);   // ... box lane + 32
                uint64_t alive = cand;
                for (;;) {   // warp-uniform (same rule as block_nms_sweep's chunk resolve)
                    const unsigned sa = __ballot_sync(0xffffffffu, (ca & alive) != 0ull);
                    const unsigned sb = __ballot_sync(0xffffffffu, (cb & alive) != 0ull);
                    const uint64_t next = cand & ~((uint64_t)sa | ((uint64_t)sb << 32));
                    if (next == alive) break;
                    alive = next;
                }
                const uint64_t before = rows_in_smem ? own_word : ld_relaxed_u64(keepw + c);   // this CTA is the word's only writer
                if (lane == 0 && alive != before) {
                    atomicExch(reinterpret_cast<unsigned long long*>(keepw + c), (unsigned long long)alive);   // see nms_fixpoint_pub_kernel
                    changed = true;   // thread 0 is the one that arrives at the barrier
                }
                own_word = alive;
            }
        }
        const unsigned long long v = grid_meet(gbar, &s_word, pass, changed);
        const unsigned field = (pass & 1u) ? (unsigned)((v >> 20) & 0x3fffffu) : (unsigned)((v >> 42) & 0x3fffffu);
        if (field == seen[pass & 1u]) break;   // nobody changed a word in this pass: the fixed point
        seen[pass & 1u] = field;
    }
    // survivors (score order) -> bits by ORIGINAL index
    for (int c = blockIdx.x; c < W; c += G) {
        const uint64_t alive = ld_relaxed_u64(keepw + c);
        if (tid < 64 && ((alive >> tid) & 1ull)) {
            const int o = order[c * 64 + tid];
            MRCNN_DBG(o >= 0 && o < N);
            atomicOr(obits + (o >> 5), 1u << (o & 31));
        }
    }
    grid_meet(gbar, &s_word, pass + 1, false);
    if (blockIdx.x != 0) return;
    // ascending compaction by CTA 0: thread t owns the contiguous words [t * per, (t + 1) * per)
    const int words = (N + 31) >> 5;
    const int per = (words + nt - 1) / nt;
    const int beg = min(words, tid * per), end = min(words, beg + per);
    int cnt = 0;
    for (int i = beg; i < end; ++i) cnt += __popc(__ldcg(obits + i));
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int u = (lane < (nt >> 5)) ? s_warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, u, o);
            if (lane >= o) u += x;
        }
        s_warp_sums[lane] = u;
    }
    __syncthreads();
    int pos = incl - cnt + (warp > 0 ? s_warp_sums[warp - 1] : 0);
    for (int i = beg; i < end; ++i) {
        unsigned b = __ldcg(obits + i);
        while (b) {
            const int j = __ffs((int)b) - 1;
            b &= b - 1;
            MRCNN_DBG(pos < N);
            keep_out[pos++] = (int64_t)i * 32 + j;
        }
    }
    if (tid == nt - 1) *count_out = pos;   // the last thread's end position is the total
}

// The same iteration for W <= SM count (one chunk per CTA, up to ~9400 boxes) WITHOUT a central barrier: a CTA publishes its
// survivor word as two 64-bit values (32 survivor bits | pass tag | changed bit) into the buffer of the pass's parity, and
// every CTA polls all 2 W values of the previous pass - the data IS the flag, so a pass costs one store-to-load trip through L2
// instead of fence + arrival + poll + reload (~3.6 k cycles -> ~2 k).  Everybody reads the same published values, so everybody
// sees the same "nothing changed in the last pass" and stops together; a buffer is only overwritten two passes later, when every
// CTA has read it (it published the pass in between).  CTA 0 then holds the final words in shared memory and emits the list.
__global__ void __launch_bounds__(1024) nms_fixpoint_pub_kernel(const uint64_t* __restrict__ lower, const int32_t* __restrict__ order,
                                                                int N, int W, int keep_off, uint64_t* __restrict__ pub,
                                                                int64_t* __restrict__ keep_out, int32_t* __restrict__ count_out) {
    extern __shared__ __align__(128) uint64_t s_dyn[];
    uint64_t* s_rows = s_dyn;               // [c][64]; CTA 0 (no rows) keeps the order map and the survivor bits by original index here
    uint64_t* s_keep = s_dyn + keep_off;    // [W]
    __shared__ unsigned s_sup[2];
    __shared__ int s_any[2];   // by pass parity: did any chunk change in the pass just read
    __shared__ int s_warp_sums[32];
    __shared__ uint64_t s_bar;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int S = nt >> 6, box = tid & 63, slice = tid >> 6;
    const int c = blockIdx.x;
    const uint64_t* tiles = lower + (size_t)c * (c + 1) / 2 * 64;
    if (tid == 0 && c > 0) {   // the chunk's rows: one bulk copy, waited for just before their first use
        mbar_init(&s_bar, 1);
        fence_barrier_init();
        mbar_expect_tx(&s_bar, (uint32_t)c * 512u);
        bulk_g2s(s_rows, tiles, (uint32_t)c * 512u, &s_bar);
    }
    uint64_t own_ca = 0ull, own_cb = 0ull, own_word = 0ull;
    if (warp == 0) {
        own_ca = __ldg(tiles + (size_t)c * 64 + lane);
        own_cb = __ldg(tiles + (size_t)c * 64 + lane + 32);
        const int nrows = min(64, N - c * 64);
        own_word = nrows == 64 ? ~0ull : ((1ull << nrows) - 1ull);
    }
    if (tid < 2) s_any[tid] = 0;
    // CTA 0 has no rows to stage (nothing precedes chunk 0): it uses the room for the score order -> original index map and, later,
    // the survivor bits by original index, so that the emit at the end never leaves shared memory
    int* s_order = reinterpret_cast<int*>(s_rows + W);   // after the 2 W survivor-bit words
    if (c == 0) {
        for (int i = tid; i < N; i += nt) s_order[i] = __ldg(order + i);
        for (int i = tid; i < 2 * W; i += nt) reinterpret_cast<unsigned*>(s_rows)[i] = 0u;
    }
    __syncthreads();
    for (unsigned pass = 1;; ++pass) {
        if (tid < 2) s_sup[tid] = 0u;
        for (int i = tid; i < 2 * W; i += nt) {
            const uint64_t* src = pub + (size_t)((pass - 1) & 1u) * 2 * W + i;
            uint64_t v;
            do {
                v = ld_relaxed_u64(src);
            } while ((unsigned)(v >> 33) != pass - 1);
            reinterpret_cast<unsigned*>(s_keep)[i] = (unsigned)v;   // little endian: halves 2 w, 2 w + 1 make word w
            if ((v >> 32) & 1ull) s_any[pass & 1u] = 1;
        }
        __syncthreads();
        if (pass > 1 && s_any[pass & 1u] == 0) break;   // the previous pass changed nothing: s_keep is the answer
        if (tid == 0) s_any[(pass + 1u) & 1u] = 0;   // last read one pass ago, before that pass's second barrier; next set after this pass's
        if (pass == 1 && c > 0) mbar_wait(&s_bar, 0u);
        uint64_t acc = 0ull;
#pragma unroll 4
        for (int w = slice; w < c; w += S) {
            MRCNN_DBG(w >= 0 && w < W && w * 64 + box < c * 64);
            acc |= s_rows[w * 64 + box] & s_keep[w];
        }
        const unsigned hit = __ballot_sync(0xffffffffu, acc != 0ull);
        if (lane == 0 && hit) atomicOr(&s_sup[warp & 1], hit);
        __syncthreads();
        if (warp == 0) {
            const int nrows = min(64, N - c * 64);
            const uint64_t valid = nrows == 64 ? ~0ull : ((1ull << nrows) - 1ull);
            const uint64_t cand = valid & ~((uint64_t)s_sup[0] | ((uint64_t)s_sup[1] << 32));
            uint64_t alive = cand;
            for (;;) {
                const unsigned sa = __ballot_sync(0xffffffffu, (own_ca & alive) != 0ull);
                const unsigned sb = __ballot_sync(0xffffffffu, (own_cb & alive) != 0ull);
                const uint64_t next = cand & ~((uint64_t)sa | ((uint64_t)sb << 32));
                if (next == alive) break;
                alive = next;
            }
            if (lane < 2) {
                const uint64_t tag = (uint64_t)(pass * 2u + (alive != own_word ? 1u : 0u)) << 32;
                // an atomic exchange, not a store: measured on B200, a plain (even .relaxed.gpu) store took ~3 us to reach the
                // pollers of the other SMs, the atomic is performed at L2 at once (pass time 4.4 -> 1.7 us)
                atomicExch(reinterpret_cast<unsigned long long*>(pub + (size_t)(pass & 1u) * 2 * W + 2 * c + lane),
                           (unsigned long long)(tag | ((alive >> (32 * lane)) & 0xffffffffull)));
            }
            own_word = alive;
        }
    }
    if (c != 0) return;
    // CTA 0: survivors (score order) -> bits by original index in shared memory -> ascending list, one thread per index
    unsigned* s_obits = reinterpret_cast<unsigned*>(s_rows);   // zeroed at kernel start
    const int words = (N + 31) >> 5;
    for (int i = tid; i < N; i += nt)
        if ((s_keep[i >> 6] >> (i & 63)) & 1ull) {
            const int o = s_order[i];
            MRCNN_DBG(o >= 0 && o < N);
            atomicOr(s_obits + (o >> 5), 1u << (o & 31));
        }
    __syncthreads();
    // exclusive prefix of the per-word counts: thread t owns the contiguous words [t * per, (t + 1) * per)
    int* s_base = reinterpret_cast<int*>(s_keep);   // [words] (the survivor words are not needed any more; W * 8 bytes = 2 W ints)
    const int per = (words + nt - 1) / nt;
    const int beg = min(words, tid * per), end = min(words, beg + per);
    int cnt = 0;
    for (int i = beg; i < end; ++i) cnt += __popc(s_obits[i]);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int u = (lane < (nt >> 5)) ? s_warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, u, o);
            if (lane >= o) u += x;
        }
        s_warp_sums[lane] = u;
    }
    __syncthreads();
    int pos = incl - cnt + (warp > 0 ? s_warp_sums[warp - 1] : 0);
    for (int i = beg; i < end; ++i) {
        s_base[i] = pos;
        pos += __popc(s_obits[i]);
    }
    if (tid == nt - 1) *count_out = pos;   // the last thread's end position is the total
    __syncthreads();
    for (int i = tid; i < words * 32; i += nt) {   // a warp per word: coalesced stores
        const unsigned b = s_obits[i >> 5];
        if ((b >> lane) & 1u) {
            const int at = s_base[i >> 5] + __popc(b & ((1u << lane) - 1u));
            MRCNN_DBG(at < N && i < N);
            keep_out[at] = (int64_t)i;
        }
    }
}

__global__ void nms_empty_kernel(int32_t* count_out) { *count_out = 0; }

static size_t sweep_smem_bytes(int W, bool staged) {
    return (size_t)8 * ((staged ? (size_t)2 * 64 * W : 0) + 2 * (size_t)W);
}

}  // namespace mrcnn

using namespace mrcnn;

extern "C" {

size_t mrcnn_nms_workspace_bytes(int N) { return carve_nms(nullptr, N).bytes; }

int mrcnn_nms(const float* dets, int N, float threshold, int64_t* keep_out, int32_t* count_out, void* workspace,
              size_t workspace_bytes, mrcnn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRCNN_REQUIRE(N >= 0, "mrcnn_nms: N must be >= 0");
    MRCNN_REQUIRE(N <= (1 << 17), "mrcnn_nms: N = %d exceeds the supported maximum of 131072 boxes", N);
    MRCNN_REQUIRE_DEV(count_out);
    if (N == 0) {  // nms_cpu.cpp:16-18
        nms_empty_kernel<<<1, 1, 0, stream>>>(count_out);
        MRCNN_LAUNCH_CHECK();
        return MRCNN_OK;
    }
    MRCNN_REQUIRE_DEV(dets);
    MRCNN_REQUIRE_DEV(keep_out);
    MRCNN_REQUIRE_DEV(workspace);
    const NmsWorkspace ws = carve_nms(workspace, N);
    if (workspace_bytes < ws.bytes)
        return fail(MRCNN_E_WORKSPACE, "mrcnn_nms: workspace of %zu bytes < required %zu", workspace_bytes, ws.bytes);
    MRCNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "mrcnn_nms: workspace must be 256-byte aligned");
    const int P = next_pow2(N);
    const int W = (N + 63) / 64;

    if (P > 1024 && P <= kSortTile) {
        const int T = P / 1024;
        nms_tile_sort_kernel<<<T, 1024, 0, stream>>>(dets, N, ws.sortbuf);
        MRCNN_LAUNCH_CHECK();
        const size_t smem = (size_t)P * 8;
        MRCNN_CUDA(cudaFuncSetAttribute(nms_rank_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortTile * 8));
        nms_rank_gather_kernel<<<T * kRankSplit, 1024 / kRankSplit, smem, stream>>>(dets, ws.sortbuf, N, T, ws.sbox, ws.sarea, ws.order);
        MRCNN_LAUNCH_CHECK();
    } else if (P <= kSortTile) {
        const size_t smem = (size_t)P * 8;
        MRCNN_CUDA(cudaFuncSetAttribute(nms_prepare_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortTile * 8));
        const int threads = P >= 2048 ? 1024 : (P >= 64 ? P / 2 : 32);
        nms_prepare_small_kernel<<<1, threads, smem, stream>>>(dets, N, P, ws.sbox, ws.sarea, ws.order);
        MRCNN_LAUNCH_CHECK();
    } else {
        const size_t smem = (size_t)kSortTile * 8;
        MRCNN_CUDA(cudaFuncSetAttribute(bitonic_tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MRCNN_CUDA(cudaFuncSetAttribute(bitonic_tile_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nms_fill_keys_kernel<<<(P + 255) / 256, 256, 0, stream>>>(dets, N, P, ws.sortbuf);
        const int tiles = P / kSortTile;
        bitonic_tile_sort_kernel<<<tiles, 1024, smem, stream>>>(ws.sortbuf);
        for (unsigned k = 2u * kSortTile; k <= (unsigned)P; k <<= 1) {
            for (unsigned j = k >> 1; j >= (unsigned)kSortTile; j >>= 1)
                bitonic_global_step_kernel<<<(P / 2 + 255) / 256, 256, 0, stream>>>(ws.sortbuf, (unsigned)P, j, k);
            bitonic_tile_merge_kernel<<<tiles, 1024, smem, stream>>>(ws.sortbuf, k);
        }
        nms_gather_kernel<<<(N + 255) / 256, 256, 0, stream>>>(dets, ws.sortbuf, N, ws.sbox, ws.sarea, ws.order);
        MRCNN_LAUNCH_CHECK();
    }

    // Route: the grid-wide fixed-point iteration from kFixpointMinW chunks on (MRCNN_NMS_SWEEP=serial|fixpoint forces one);
    // below that the single-CTA sweep has fewer barriers to pay than the cooperative launch costs.
    static const char* route_env = getenv("MRCNN_NMS_SWEEP");
    const bool fixpoint = route_env && route_env[0] == 's' ? false : (route_env && route_env[0] == 'f' ? true : W >= kFixpointMinW);
    if (fixpoint) {
        const int tiles = W * (W + 1) / 2;
        nms_mask_lower_kernel<<<(tiles + kMaskTiles - 1) / kMaskTiles, 64 * kMaskTiles, 0, stream>>>(ws.sbox, ws.sarea, N, W, threshold, ws.mask, ws.keepw, ws.obits, ws.gbar, ws.pub);
        MRCNN_LAUNCH_CHECK();
        const int G = min(W, sm_count());
        int rows_in_smem = (W <= G) ? 1 : 0;
        size_t smem = (size_t)W * 8 + (rows_in_smem ? std::max((size_t)(W > 1 ? W - 1 : 1) * 512, (size_t)W * 8 + (size_t)W * 256) : 0);
        static const int fp_threads = getenv("MRCNN_NMS_THREADS") ? atoi(getenv("MRCNN_NMS_THREADS")) : 0;   // experiment knob
        int threads = W > 32 ? 512 : 256;
        if (fp_threads >= 64 && fp_threads <= 1024 && fp_threads % 64 == 0) threads = fp_threads;
        static const bool no_pub = getenv("MRCNN_NMS_PUB") != nullptr && getenv("MRCNN_NMS_PUB")[0] == '0';   // experiment knob
        if (rows_in_smem && !no_pub) {
            MRCNN_CUDA(cudaFuncSetAttribute(nms_fixpoint_pub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 512 + 160 * 8));
            const uint64_t* lower = ws.mask;
            const int32_t* order = ws.order;
            uint64_t* pub = ws.pub;
            int n = N, w = W;
            int keep_off = (int)((smem - (size_t)W * 8) / 8);
            void* args[] = {(void*)&lower, (void*)&order, (void*)&n, (void*)&w, (void*)&keep_off, (void*)&pub, (void*)&keep_out, (void*)&count_out};
            MRCNN_CUDA(cudaLaunchCooperativeKernel((const void*)nms_fixpoint_pub_kernel, dim3(G), dim3(threads), args, smem, stream));
            return MRCNN_OK;
        }
        MRCNN_CUDA(cudaFuncSetAttribute(nms_fixpoint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 512 + 160 * 8));
        const uint64_t* lower = ws.mask;
        const int32_t* order = ws.order;
        uint64_t* keepw = ws.keepw;
        uint32_t* obits = ws.obits;
        unsigned long long* gbar = ws.gbar;
        int n = N, w = W;
        void* args[] = {(void*)&lower, (void*)&order, (void*)&n, (void*)&w, (void*)&rows_in_smem, (void*)&keepw,
                        (void*)&obits, (void*)&gbar, (void*)&keep_out, (void*)&count_out};
        MRCNN_CUDA(cudaLaunchCooperativeKernel((const void*)nms_fixpoint_kernel, dim3(G), dim3(threads), args, smem, stream));
        return MRCNN_OK;
    }

    nms_mask_kernel<<<dim3(W + 1, (W + 1) / 2), 64, 0, stream>>>(ws.sbox, ws.sarea, N, W, threshold, ws.mask, ws.diag_t);
    MRCNN_LAUNCH_CHECK();

    // MRCNN_NMS_STAGED=0: the propagate step reads the survivors' rows straight from L2 instead of from the staged block
    // (measured slower at 6000 boxes: 388 us against 207 us before the prefetch of the transposed diagonal words).
    static const bool no_stage = getenv("MRCNN_NMS_STAGED") != nullptr && getenv("MRCNN_NMS_STAGED")[0] == '0';
    const bool staged = !no_stage && W <= kSweepStageMaxW;
    const size_t smem = sweep_smem_bytes(W, staged);
    MRCNN_REQUIRE(smem <= 220 * 1024, "mrcnn_nms: N too large for the single-CTA sweep");
    MRCNN_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    int threads = (W + 31) / 32 * 32;
    threads = threads < 128 ? 128 : (threads > 1024 ? 1024 : threads);
    // the propagate step spreads a chunk's survivors over several threads per word; measured on B200 (tools/time_nms.py, whole
    // nms call): 6000 boxes 342 / 246 / 207 / 230 us with 128 / 256 / 512 / 1024 threads, 2000 boxes 68 us with 256 (81 with 1024)
    if (W >= 64) threads = 512;
    else if (W >= 8) threads = 256;
    static const int force_threads = getenv("MRCNN_NMS_THREADS") ? atoi(getenv("MRCNN_NMS_THREADS")) : 0;   // experiment knob
    if (force_threads >= 128 && force_threads <= 1024 && force_threads % 32 == 0) threads = force_threads;
    nms_sweep_kernel<<<1, threads, smem, stream>>>(ws.mask, ws.order, N, W, staged ? 1 : 0, ws.flags, keep_out, count_out, ws.diag_t);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

}  // extern "C"
