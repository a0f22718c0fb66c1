// detection.cu — the detection layer for sm_100a, batched over images: ONE launch, one CTA per image.
//
// Replaces MaskRCNN.mrn_refine (model.py:1389-1487), which loops over classes in Python with a
// nonzero + sort + C++ nms + unique per class and a .tolist() host sync:
//   per-RoI argmax class -> class-specific delta decode -> scale to pixels -> clip to window -> round
//   -> drop background / low confidence -> per-class NMS -> top max_inst by score.
//
// NMS: lazy by default (stage D' below: chunks of 64 boxes against the same-class survivors so far, stop at the
// max_inst-th survivor); the N x N class-aware mask + sweep (stages D, E) stays selectable (mrcnn_set_detection_nms).
//
// Per-class NMS is done as ONE class-aware NMS over all kept RoIs sorted by score: a box may only be
// suppressed by a higher-scoring box of the SAME class, which is exactly the union of the per-class
// results (model.py:1454-1474); the final top-D (:1478-1480) is then a prefix of the sorted survivors.
#include <limits.h>

#include "api_util.h"
#include "nms_core.cuh"

namespace mrcnn {

constexpr int kDetThreads = 1024;
constexpr int kDetMaxN = 3968;  // 32768 + 48 N + 16 ceil(N / 64) bytes of shared memory must fit the 220 KB the kernel may opt into
constexpr int kDetLazyMaxInst = 1024;   // lazy NMS up to this many wanted detections
constexpr int kDetSmemMaskMaxN = 1024;  // suppression words kept in shared memory up to this many RoIs
constexpr int kDetMaxWorld = 8;         // ranks of the fused exchange (one NVSwitch domain)

struct DetParams {
    const float* rois;     // [B,N,4] normalised
    const float* probs;    // [B,N,NC]
    const float* deltas;   // [B,N,NC,4]
    const float* windows;  // [B,4]
    int B, N, NC, P, N64, W;
    float min_conf, thr;
    int max_inst;
    float std0, std1, std2, std3;
    float height, width;
    float* dets_out;      // [B,max_inst,6]
    int32_t* counts_out;  // [B]
    int32_t* index_out;   // [B,max_inst] or null
    uint64_t* gmask;      // [B][N64][W] (used when N > kDetSmemMaskMaxN)
    int mask_in_smem;
    int lazy;  // lazy NMS (no N x N mask)
    // mask-head RoIs of the detections (model.py:1188 mrn_boxes.float() * 1.0 / h), written by the same kernel
    float* mask_boxes;       // [B*max_inst,4] or null
    int32_t* mask_box_ind;   // [B*max_inst] or null: image index of every row = (ind_offset + img) % ind_mod
    int ind_offset, ind_mod;
    // fused all-gather over peer memory (NVLink / NVSwitch): every CTA stores its image's packed row straight into the receive
    // buffer of EVERY rank; the last CTA to finish raises this rank's flag on every rank.  world == 0: no exchange.
    float* peer[kDetMaxWorld];          // rank r's exchange buffer, mapped into this process (own buffer included)
    int world, rank, image_offset, total_images;
    int32_t* state;                     // local: [0] epoch of the last finished exchange, [1] CTAs done in this launch
};

// Exchange buffer of one rank, floats: [2 parities][total_images][max_inst * 6 + 1] packed rows (detections, zero padded, then
// the count), followed by [world] uint32 flags: flags[r] = epoch of the last exchange whose rows rank r has delivered here.
__device__ __host__ inline size_t det_row_width(int max_inst) { return (size_t)max_inst * 6 + 1; }
__device__ __host__ inline size_t det_flags_offset(int total_images, int max_inst) { return 2 * (size_t)total_images * det_row_width(max_inst); }

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kDetThreads, 1) detection_layer_kernel(const DetParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bars[2];
    __shared__ int s_total;
    __shared__ int s_count;
    __shared__ int s_prefix[kDetMaxN / 64 + 1];
    __shared__ int s_ksel[kDetLazyMaxInst + 64];  // lazy NMS: the survivors so far (positions in score order)
    __shared__ int s_hit[64];
    __shared__ uint64_t s_d[64];

    const int img = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, NC = p.NC;

    // shared-memory carve-up
    unsigned char* ptr = smem_raw;
    uint64_t* sortbuf = reinterpret_cast<uint64_t*>(ptr);  ptr += (size_t)p.P * (p.P == kDetThreads ? 16 : 8);  // + exchange buffer
    float4* rbox = reinterpret_cast<float4*>(ptr);          ptr += (size_t)N * 16;   // refined box per RoI (input order)
    float4* sbox = reinterpret_cast<float4*>(ptr);          ptr += (size_t)N * 16;   // boxes in score order
    float* sarea = reinterpret_cast<float*>(ptr);           ptr += (size_t)N * 4;
    int* scls = reinterpret_cast<int*>(ptr);                ptr += (size_t)N * 4;
    int* rcls = reinterpret_cast<int*>(ptr);                ptr += (size_t)N * 4;    // class per RoI (input order)
    float* rscore = reinterpret_cast<float*>(ptr);          ptr += (size_t)N * 4;
    uint64_t* remv = reinterpret_cast<uint64_t*>(ptr);      ptr += (size_t)p.W * 8;
    uint64_t* kept = reinterpret_cast<uint64_t*>(ptr);      ptr += (size_t)p.W * 8;
    uint64_t* smask = reinterpret_cast<uint64_t*>(ptr);     // [N64][W] when mask_in_smem

    const float* rois = p.rois + (size_t)img * N * 4;
    const float* probs = p.probs + (size_t)img * N * NC;
    const float* deltas = p.deltas + (size_t)img * N * NC * 4;
    const float* win = p.windows + (size_t)img * 4;

    if (tid == 0) s_count = 0;

    // ---- A. argmax over classes, one warp per RoI (model.py:1407, :1414) ----
    // A warp stalls at the first use of a load, so the probabilities of kArgRois RoIs (up to 3 loads each for <= 96
    // classes) are all requested before the first comparison: 12 loads in flight per warp instead of one.
    constexpr int kArgRois = 4;
    for (int n0 = warp * kArgRois; n0 < N; n0 += (kDetThreads / 32) * kArgRois) {
        float v[kArgRois][3];
        const bool fast = NC <= 96;
        if (fast) {
#pragma unroll
            for (int u = 0; u < kArgRois; ++u) {
                const float* row = probs + (size_t)min(n0 + u, N - 1) * NC;
#pragma unroll
                for (int k = 0; k < 3; ++k) v[u][k] = __ldg(row + min(lane + 32 * k, NC - 1));
            }
        }
#pragma unroll
        for (int u = 0; u < kArgRois; ++u) {
            const int n = n0 + u;
            if (n >= N) break;
            float best = -INFINITY;
            int bi = INT_MAX;
            if (fast) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int j = lane + 32 * k;
                    if (j < NC && (v[u][k] > best || bi == INT_MAX)) {  // the lane's first element initialises (handles -inf rows)
                        best = v[u][k];
                        bi = j;
                    }
                }
            } else {
                const float* row = probs + (size_t)n * NC;
                for (int j = lane; j < NC; j += 32) {
                    const float x = __ldg(row + j);
                    if (x > best || bi == INT_MAX) {
                        best = x;
                        bi = j;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) {  // first maximum wins ties
                    best = ov;
                    bi = oi;
                }
            }
            if (lane == 0) {
                rcls[n] = (bi == INT_MAX) ? 0 : bi;
                rscore[n] = best;
            }
        }
    }
    __syncthreads();

    // ---- B. decode / scale / clip / round / filter, one thread per RoI (model.py:1415-1443) ----
    const float wy1 = __ldg(win + 0), wx1 = __ldg(win + 1), wy2 = __ldg(win + 2), wx2 = __ldg(win + 3);
    for (int n = tid; n < p.P; n += kDetThreads) {
        uint64_t key = 0ull;
        if (n < N) {
            const int c = rcls[n];
            const float sc = rscore[n];
            const float4 r = __ldg(reinterpret_cast<const float4*>(rois) + n);
            const float4 dl = __ldg(reinterpret_cast<const float4*>(deltas) + (size_t)n * NC + c);
            const float b[4] = {r.x, r.y, r.z, r.w};
            const float d[4] = {__fmul_rn(dl.x, p.std0), __fmul_rn(dl.y, p.std1), __fmul_rn(dl.z, p.std2),
                                __fmul_rn(dl.w, p.std3)};
            float o[4];
            box_refine(b, d, o);
            float4 q;
            q.x = rintf(clampf(__fmul_rn(o[0], p.height), wy1, wy2));  // :1426, :1429, :1432 (half to even)
            q.y = rintf(clampf(__fmul_rn(o[1], p.width), wx1, wx2));
            q.z = rintf(clampf(__fmul_rn(o[2], p.height), wy1, wy2));
            q.w = rintf(clampf(__fmul_rn(o[3], p.width), wx1, wx2));
            rbox[n] = q;
            bool keep = c > 0;                                         // :1437
            if (p.min_conf > 0.0f) keep = keep && (sc >= p.min_conf);  // :1441-1442
            if (keep) {
                key = make_sort_key(sc, (uint32_t)n);
                atomicAdd(&s_count, 1);
            }
        }
        sortbuf[n] = key;
    }
    __syncthreads();
    const int M = s_count;  // RoIs entering NMS

    // ---- C. sort by score (descending; ties -> lower RoI index) and gather ----
    if (p.P == kDetThreads) {  // up to 1024 RoIs: one key per thread, register / shuffle network
        const uint64_t mine = sortbuf[tid];
        __syncthreads();  // every key is in a register before sortbuf is reused as the exchange buffer
        const uint64_t sorted = block_bitonic_desc_1024_reg(mine, sortbuf);
        __syncthreads();  // the last exchange stage has been read
        sortbuf[tid] = sorted;
        __syncthreads();
    } else {
        block_bitonic_desc(sortbuf, p.P, 0u, 2u, 1u, (unsigned)p.P);
    }
    for (int i = tid; i < M; i += kDetThreads) {
        const int n = (int)sort_key_index(sortbuf[i]);
        const float4 b = rbox[n];
        sbox[i] = b;
        sarea[i] = box_area_p1(b);
        scls[i] = rcls[n];
    }
    __syncthreads();

    const int W = (M + 63) >> 6;
    if (p.lazy) {
        // ---- D'. lazy class-aware NMS: only the first max_inst survivors are needed (:1478-1480) and a box only has to be
        // compared with the SURVIVORS of its own class that precede it.  Boxes are taken 64 at a time in score order:
        //   pull      the chunk's 64 columns x 16 survivor slices (a column stops at its first hit; the class test comes
        //             first, so with 81 classes almost no pair reaches the IoU arithmetic);
        //   diagonal  the chunk's own 64 x 64 tile;
        //   resolve   warp 0, Jacobi iteration on the triangular dependency (= the greedy answer), appends the survivors.
        // The loop ends at the max_inst-th survivor: ~64 * sum(S) tests instead of the M^2 / 2 of the mask (stage D below),
        // which was ~300 of this kernel's 400 us at 1000 RoIs.  Same decisions (iou_ge_m), same greedy order.
        const float margin = __fadd_rn(__fmul_rn(fabsf(p.thr), 1e-6f), 1e-37f);
        for (int w = tid; w < W; w += kDetThreads) kept[w] = 0ull;
        if (tid == 0) s_total = 0;
        __syncthreads();
        int S = 0;
        for (int c = 0; c < W && S < p.max_inst; ++c) {
            const int col0 = c << 6;
            const int ncols = min(64, M - col0);
            if (tid < 64) s_hit[tid] = 0;
            __syncthreads();
            {   // pull
                const int j = tid & 63, g = tid >> 6;
                if (j < ncols) {
                    const float4 cb = sbox[col0 + j];
                    const float ca = sarea[col0 + j];
                    const int cc = scls[col0 + j];
                    volatile int* hit = s_hit + j;
                    for (int s_ = g; s_ < S; s_ += kDetThreads / 64) {
                        if (*hit) break;
                        const int k = s_ksel[s_];
                        if (scls[k] == cc && iou_ge_m(sbox[k], sarea[k], cb, ca, p.thr, margin)) {
                            *hit = 1;
                            break;
                        }
                    }
                }
            }
            {   // diagonal tile by column: thread = (column i, rows 4q .. 4q + 3), nibbles ORed with shuffles
                const int i = tid >> 4, q = tid & 15;
                uint32_t nib = 0;
                if (i < ncols) {
                    const float4 cb = sbox[col0 + i];
                    const float ca = sarea[col0 + i];
                    const int cc = scls[col0 + i];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int row = 4 * q + k;
                        if (row < i && scls[col0 + row] == cc && iou_ge_m(sbox[col0 + row], sarea[col0 + row], cb, ca, p.thr, margin))
                            nib |= 1u << k;
                    }
                }
                uint32_t lo = q < 8 ? nib << (4 * q) : 0u, hi = q >= 8 ? nib << (4 * (q - 8)) : 0u;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    lo |= __shfl_xor_sync(0xffffffffu, lo, o);
                    hi |= __shfl_xor_sync(0xffffffffu, hi, o);
                }
                if (q == 0) s_d[i] = ((uint64_t)hi << 32) | lo;
            }
            __syncthreads();
            if (tid < 32) {
                const uint32_t h_lo = __ballot_sync(0xffffffffu, s_hit[lane] != 0);
                const uint32_t h_hi = __ballot_sync(0xffffffffu, s_hit[lane + 32] != 0);
                uint64_t cand = ~(((uint64_t)h_hi << 32) | h_lo);
                if (ncols < 64) cand &= (1ull << ncols) - 1ull;
                const uint64_t col_lo = s_d[lane], col_hi = s_d[lane + 32];
                const bool c_lo = (cand >> lane) & 1ull, c_hi = (cand >> (lane + 32)) & 1ull;
                uint64_t alive = cand;
                for (;;) {
                    const uint32_t a_lo = __ballot_sync(0xffffffffu, c_lo && (col_lo & alive) == 0ull);
                    const uint32_t a_hi = __ballot_sync(0xffffffffu, c_hi && (col_hi & alive) == 0ull);
                    const uint64_t next = ((uint64_t)a_hi << 32) | a_lo;
                    if (next == alive) break;
                    alive = next;
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int b = lane + 32 * half;
                    if ((alive >> b) & 1ull) s_ksel[S + __popcll(alive & ((1ull << b) - 1ull))] = col0 + b;
                }
                if (lane == 0) {
                    kept[c] = alive;
                    s_total = S + __popcll(alive);
                }
            }
            __syncthreads();
            S = s_total;
        }
    } else {
        // ---- D. class-aware suppression words, upper triangle ----
        uint64_t* mask = p.mask_in_smem ? smask : (p.gmask + (size_t)img * p.N64 * p.W);
        for (int item = tid; item < M * W; item += kDetThreads) {
            const int row = item / W;
            const int cb = item - row * W;
            if (cb < (row >> 6)) continue;
            const int col0 = cb << 6;
            const int ncols = min(64, M - col0);
            mask[(size_t)row * W + cb] =
                suppression_word<true>(sbox[row], sarea[row], scls[row], row, sbox + col0, sarea + col0, scls + col0, col0, ncols, p.thr);
        }
        __threadfence_block();
        __syncthreads();

        // ---- E. greedy sweep; only the first max_inst survivors are needed (:1478-1480) ----
        SweepSmem sm;
        sm.stage = nullptr;
        sm.remv = remv;
        sm.kept = kept;
        sm.bars = bars;
        sm.total = &s_total;
        block_nms_sweep(mask, M, W, sm, false, p.max_inst);
    }

    // ---- F. emit ----
    if (tid == 0) {
        int run = 0;
        for (int w = 0; w < W; ++w) {
            s_prefix[w] = run;
            run += __popcll(kept[w]);
        }
        s_prefix[W] = run;
    }
    __syncthreads();
    const int D = min(s_prefix[W], p.max_inst);
    float* out = p.dets_out + (size_t)img * p.max_inst * 6;
    int32_t* iout = p.index_out ? p.index_out + (size_t)img * p.max_inst : nullptr;
    for (int i = tid; i < M; i += kDetThreads) {
        const uint64_t kw = kept[i >> 6];
        if ((kw >> (i & 63)) & 1ull) {
            const int r = s_prefix[i >> 6] + __popcll(kw & ((1ull << (i & 63)) - 1ull));
            MRCNN_DBG(r >= 0 && i < M);
            if (r < p.max_inst) {
                const float4 b = sbox[i];
                float* o = out + (size_t)r * 6;
                o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w;
                o[4] = sort_key_score(sortbuf[i]);
                o[5] = (float)scls[i];
                if (iout) iout[r] = (int32_t)sort_key_index(sortbuf[i]);
            }
        }
    }
    for (int e = D * 6 + tid; e < p.max_inst * 6; e += kDetThreads) out[e] = 0.f;
    if (iout)
        for (int r = D + tid; r < p.max_inst; r += kDetThreads) iout[r] = -1;
    if (tid == 0) p.counts_out[img] = D;
    if (p.mask_boxes == nullptr && p.world == 0) return;

    // ---- G. what follows the detection layer, without a launch in between
    __syncthreads();   // the image's rows are complete in global memory (written by this CTA)
    if (p.mask_boxes) {
        // model.py:1188: mrn_rois = mrn_boxes.float() * 1.0 / h  (all four coordinates by the image height)
        for (int e = tid; e < p.max_inst * 4; e += kDetThreads) {
            const int r = e >> 2, k = e & 3;
            p.mask_boxes[((size_t)img * p.max_inst + r) * 4 + k] = __fdiv_rn(out[(size_t)r * 6 + k], p.height);
        }
        if (p.mask_box_ind)
            for (int r = tid; r < p.max_inst; r += kDetThreads)
                p.mask_box_ind[(size_t)img * p.max_inst + r] = (p.ind_offset + img) % p.ind_mod;
    }
    if (p.world > 0) {
        const int epoch0 = p.state[0];                   // every CTA reads it before any CTA of this launch can bump it
        const size_t width = det_row_width(p.max_inst);
        MRCNN_DBG(p.image_offset + img < p.total_images && p.rank < p.world && epoch0 >= 0);
        const size_t row = ((size_t)((epoch0 + 1) & 1) * p.total_images + (size_t)(p.image_offset + img)) * width;
        for (int e = tid; e < (int)width; e += kDetThreads) {
            const float v = (e < p.max_inst * 6) ? out[e] : (float)D;
#pragma unroll
            for (int r = 0; r < kDetMaxWorld; ++r)
                if (r < p.world) p.peer[r][row + e] = v;    // peer stores: NVLink writes into rank r's memory
        }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            const int done = atomicAdd(p.state + 1, 1);
            if (done == p.B - 1) {                       // last CTA of this rank: all of the rank's rows are on their way
                p.state[1] = 0;
                __threadfence_system();
                const size_t fo = det_flags_offset(p.total_images, p.max_inst);
                for (int r = 0; r < p.world; ++r)
                    st_release_sys(reinterpret_cast<uint32_t*>(p.peer[r] + fo) + p.rank, (uint32_t)(epoch0 + 1));
                p.state[0] = epoch0 + 1;
            }
        }
    }
}

// Second half of the fused all-gather: CTA i waits until the rank that owns image i has raised its flag for the current epoch
// (a peer's release store, observed with an acquire load from LOCAL memory), then copies the image's packed row out of the
// receive buffer into the caller's [total, max_inst, 6] / [total] tensors.  Image order = rank order (contiguous shards).
__global__ void __launch_bounds__(128) detection_collect_kernel(const float* buf, int world, int total, int max_inst, const int32_t* state,
                                                                float* dets_all, int32_t* counts_all) {
    const int img = blockIdx.x;
    const uint32_t epoch = (uint32_t)state[0];           // bumped by this rank's detection kernel earlier in the stream
    const int base = total / world, rem = total % world; // shard_range(): the first `rem` ranks own base + 1 images
    const int cut = rem * (base + 1);
    const int owner = (img < cut) ? img / (base + 1) : rem + (img - cut) / (base > 0 ? base : 1);
    MRCNN_DBG(owner >= 0 && owner < world);
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(buf + det_flags_offset(total, max_inst));
    if (threadIdx.x == 0) {
        while ((int32_t)(ld_acquire_sys(flags + owner) - epoch) < 0) __nanosleep(64);
    }
    __syncthreads();
    const size_t width = det_row_width(max_inst);
    const float* row = buf + ((size_t)(epoch & 1u) * total + img) * width;
    for (int e = threadIdx.x; e < max_inst * 6; e += blockDim.x) dets_all[(size_t)img * max_inst * 6 + e] = __ldcv(row + e);
    if (threadIdx.x == 0) counts_all[img] = (int32_t)__ldcv(row + max_inst * 6);
}

__global__ void detection_empty_kernel(float* dets, size_t n, int32_t* counts, int32_t* index, int B, int max_inst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dets[i] = 0.f;
    if (i < (size_t)B) counts[i] = 0;
    if (index && i < (size_t)B * max_inst) index[i] = -1;
}

static int g_detection_nms_algo = MRCNN_PROPOSAL_NMS_AUTO;

static size_t det_smem_bytes(int N, int P, int W, bool mask_in_smem) {
    size_t b = (size_t)P * (P == kDetThreads ? 16 : 8) + (size_t)N * (16 + 16 + 4 + 4 + 4 + 4) + (size_t)W * 16;
    if (mask_in_smem) b += (size_t)(W * 64) * W * 8;
    return align_up(b, 16);
}

}  // namespace mrcnn

using namespace mrcnn;

extern "C" {

int mrcnn_set_detection_nms(int algo) {
    MRCNN_REQUIRE(algo == MRCNN_PROPOSAL_NMS_AUTO || algo == MRCNN_PROPOSAL_NMS_MASK || algo == MRCNN_PROPOSAL_NMS_LAZY,
                  "mrcnn_set_detection_nms: unknown algorithm %d", algo);
    g_detection_nms_algo = algo;
    return MRCNN_OK;
}

size_t mrcnn_detection_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= kDetSmemMaskMaxN) return 256;
    const size_t N64 = align_up((size_t)N, 64);
    return align_up((size_t)B * N64 * (N64 / 64) * 8, 256);
}

}  // extern "C"

namespace mrcnn {

struct DetExtras {
    float* mask_boxes = nullptr;
    int32_t* mask_box_ind = nullptr;
    int ind_offset = 0, ind_mod = 1;
    void* const* peer_bufs_host = nullptr;
    int world = 0, rank = 0, image_offset = 0, total_images = 0;
    int32_t* state = nullptr;
};

static int detection_launch(const float* rois, const float* probs, const float* deltas, const float* windows, int B, int N,
                            int NC, float min_confidence, float nms_threshold, int max_inst, const float* std4_host,
                            float height, float width, float* dets_out, int32_t* counts_out, int32_t* index_out,
                            void* workspace, size_t workspace_bytes, const DetExtras& x, mrcnn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRCNN_REQUIRE(B > 0 && N >= 0 && NC > 0 && max_inst > 0, "mrcnn_detection_layer: bad sizes");
    MRCNN_REQUIRE(N <= kDetMaxN, "mrcnn_detection_layer: N = %d RoIs per image exceeds the supported %d", N, kDetMaxN);
    MRCNN_REQUIRE(std4_host != nullptr, "mrcnn_detection_layer: std4_host is null");
    MRCNN_REQUIRE_DEV(dets_out);
    MRCNN_REQUIRE_DEV(counts_out);
    if (index_out) MRCNN_REQUIRE_DEV(index_out);
    if (N == 0) {
        MRCNN_REQUIRE(x.world == 0 && x.mask_boxes == nullptr, "mrcnn_detection_layer_exchange: N must be positive");
        const size_t n = (size_t)B * max_inst * 6;
        detection_empty_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dets_out, n, counts_out, index_out, B, max_inst);
        MRCNN_LAUNCH_CHECK();
        return MRCNN_OK;
    }
    MRCNN_REQUIRE_DEV(rois);
    MRCNN_REQUIRE_DEV(probs);
    MRCNN_REQUIRE_DEV(deltas);
    MRCNN_REQUIRE_DEV(windows);
    MRCNN_REQUIRE((reinterpret_cast<uintptr_t>(rois) & 15u) == 0 && (reinterpret_cast<uintptr_t>(deltas) & 15u) == 0,
                  "mrcnn_detection_layer: rois and deltas must be 16-byte aligned");
    DetParams p;
    p.rois = rois; p.probs = probs; p.deltas = deltas; p.windows = windows;
    p.B = B; p.N = N; p.NC = NC;
    p.P = 32;
    while (p.P < N) p.P <<= 1;
    p.N64 = (int)align_up((size_t)N, 64);
    p.W = p.N64 / 64;
    p.min_conf = min_confidence; p.thr = nms_threshold; p.max_inst = max_inst;
    p.std0 = std4_host[0]; p.std1 = std4_host[1]; p.std2 = std4_host[2]; p.std3 = std4_host[3];
    p.height = height; p.width = width;
    p.dets_out = dets_out; p.counts_out = counts_out; p.index_out = index_out;
    p.mask_boxes = x.mask_boxes; p.mask_box_ind = x.mask_box_ind; p.ind_offset = x.ind_offset; p.ind_mod = x.ind_mod > 0 ? x.ind_mod : 1;
    p.world = x.world; p.rank = x.rank; p.image_offset = x.image_offset; p.total_images = x.total_images; p.state = x.state;
    for (int r = 0; r < kDetMaxWorld; ++r) p.peer[r] = (r < x.world) ? (float*)x.peer_bufs_host[r] : nullptr;
    p.lazy = (g_detection_nms_algo == MRCNN_PROPOSAL_NMS_LAZY ||
              (g_detection_nms_algo == MRCNN_PROPOSAL_NMS_AUTO && max_inst <= kDetLazyMaxInst)) ? 1 : 0;
    MRCNN_REQUIRE(!p.lazy || max_inst <= kDetLazyMaxInst, "mrcnn_detection_layer: max_inst too large for the lazy NMS (use MRCNN_PROPOSAL_NMS_MASK)");
    p.mask_in_smem = (N <= kDetSmemMaskMaxN && !p.lazy) ? 1 : 0;
    p.gmask = nullptr;
    if (!p.mask_in_smem && !p.lazy) {
        const size_t need = mrcnn_detection_workspace_bytes(B, N);
        MRCNN_REQUIRE_DEV(workspace);
        if (workspace_bytes < need)
            return fail(MRCNN_E_WORKSPACE, "mrcnn_detection_layer: workspace of %zu bytes < required %zu", workspace_bytes, need);
        p.gmask = (uint64_t*)workspace;
    }
    const size_t smem = det_smem_bytes(N, p.P, p.W, p.mask_in_smem != 0);
    MRCNN_REQUIRE(smem <= 220 * 1024, "mrcnn_detection_layer: shared memory need %zu exceeds the SM", smem);
    MRCNN_CUDA(cudaFuncSetAttribute(detection_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    detection_layer_kernel<<<B, kDetThreads, smem, stream>>>(p);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

}  // namespace mrcnn

extern "C" {

int mrcnn_detection_layer(const float* rois, const float* probs, const float* deltas, const float* windows, int B, int N,
                          int NC, float min_confidence, float nms_threshold, int max_inst, const float* std4_host,
                          float height, float width, float* dets_out, int32_t* counts_out, int32_t* index_out,
                          void* workspace, size_t workspace_bytes, mrcnn_stream_t stream) {
    return detection_launch(rois, probs, deltas, windows, B, N, NC, min_confidence, nms_threshold, max_inst, std4_host, height, width,
                            dets_out, counts_out, index_out, workspace, workspace_bytes, DetExtras(), stream);
}

size_t mrcnn_detection_exchange_bytes(int world, int total_images, int max_inst) {
    if (world <= 0 || total_images <= 0 || max_inst <= 0) return 0;
    return align_up((det_flags_offset(total_images, max_inst) + (size_t)world) * sizeof(float), 256);
}

int mrcnn_detection_layer_exchange(const float* rois, const float* probs, const float* deltas, const float* windows, int B, int N,
                                   int NC, float min_confidence, float nms_threshold, int max_inst, const float* std4_host,
                                   float height, float width, float* dets_out, int32_t* counts_out, float* mask_boxes,
                                   int32_t* mask_box_ind, int ind_offset, int ind_mod, void* const* peer_bufs_host, int world,
                                   int rank, int image_offset, int total_images, int32_t* state, void* workspace,
                                   size_t workspace_bytes, mrcnn_stream_t stream) {
    MRCNN_REQUIRE(world >= 0 && world <= kDetMaxWorld, "mrcnn_detection_layer_exchange: world must be 0..%d", kDetMaxWorld);
    DetExtras x;
    if (mask_boxes) {
        MRCNN_REQUIRE_DEV(mask_boxes);
        if (mask_box_ind) MRCNN_REQUIRE_DEV(mask_box_ind);
        MRCNN_REQUIRE(ind_mod > 0, "mrcnn_detection_layer_exchange: ind_mod must be positive");
        x.mask_boxes = mask_boxes; x.mask_box_ind = mask_box_ind; x.ind_offset = ind_offset; x.ind_mod = ind_mod;
    }
    if (world > 0) {
        MRCNN_REQUIRE(peer_bufs_host != nullptr && rank >= 0 && rank < world, "mrcnn_detection_layer_exchange: bad rank / peer table");
        MRCNN_REQUIRE(image_offset >= 0 && B > 0 && image_offset + B <= total_images, "mrcnn_detection_layer_exchange: images outside [0, total)");
        MRCNN_REQUIRE_DEV(state);
        for (int r = 0; r < world; ++r) MRCNN_REQUIRE(peer_bufs_host[r] != nullptr, "mrcnn_detection_layer_exchange: peer buffer %d is null", r);
        x.peer_bufs_host = peer_bufs_host; x.world = world; x.rank = rank; x.image_offset = image_offset; x.total_images = total_images;
        x.state = state;
    }
    return detection_launch(rois, probs, deltas, windows, B, N, NC, min_confidence, nms_threshold, max_inst, std4_host, height, width,
                            dets_out, counts_out, nullptr, workspace, workspace_bytes, x, stream);
}

int mrcnn_detection_collect(const void* local_buf, int world, int total_images, int max_inst, const int32_t* state, float* dets_all,
                            int32_t* counts_all, mrcnn_stream_t stream) {
    MRCNN_REQUIRE(world > 0 && world <= kDetMaxWorld && total_images > 0 && max_inst > 0, "mrcnn_detection_collect: bad sizes");
    MRCNN_REQUIRE_DEV(local_buf);
    MRCNN_REQUIRE_DEV(state);
    MRCNN_REQUIRE_DEV(dets_all);
    MRCNN_REQUIRE_DEV(counts_all);
    detection_collect_kernel<<<total_images, 128, 0, (cudaStream_t)stream>>>((const float*)local_buf, world, total_images, max_inst, state,
                                                                               dets_all, counts_all);
    MRCNN_LAUNCH_CHECK();
    return MRCNN_OK;
}

}  // extern "C"
