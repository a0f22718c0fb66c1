// common.cuh — shared device helpers for the sm_100a RoI kernels.
//
// Parity rules: every fp32 expression that decides an index, a weight or a
// comparison is written with explicit round-to-nearest intrinsics in the reference's association
// order (the reference's x86-64 build has no FMA), so results do not depend on -fmad.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

// -DMRCNN_DEBUG (make debug -> libmrcnn_b200_debug.so, loaded when MRCNN_B200_DEBUG=1): device-side asserts on the shared-memory,
// distributed-shared-memory, ring-slot and workspace indices of the cluster / mbarrier / bulk-copy protocols.  compute-sanitizer
// is closed on the GPU pool, so the whole GPU suite is run once under this build instead (profiles/r02_debug_build.txt).
#ifdef MRCNN_DEBUG
#include <assert.h>
#define MRCNN_DBG(cond) assert(cond)
#else
#define MRCNN_DBG(cond) ((void)0)
#endif

namespace mrcnn {

// Device-side error word (bit 0: box_index out of range): a 4-byte device allocation owned by api.cu,
// handed to kernels as a parameter and polled by mrcnn_poll_device_errors().
int* device_error_word();

constexpr int kMaxPool = 64;  // largest crop side handled by the smem-staged kernels

// One interpolation tap along an axis (rows or columns) of a crop.
struct AxisTap {
    int lo;      // floor index, -1 => sample lies outside the image (extrapolate)
    int hi;      // ceil index
    float lerp;  // in - lo
};

// Sample position of output index i along an input axis of `size` pixels for a box side [a1,a2]
// (normalised).  Restates cpu/crop_cpu.cpp:52-61, :63, :76-78 (rows) and :54-55, :82-85, :94-96 (cols).
__device__ __forceinline__ AxisTap axis_tap(float a1, float a2, int size, int crop, int i) {
    const float sm1 = (float)(size - 1);
    float in;
    if (crop > 1) {
        const float scale = __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), sm1), (float)(crop - 1));
        in = __fadd_rn(__fmul_rn(a1, sm1), __fmul_rn((float)i, scale));
    } else {
        // crop_cpu.cpp:61 — the literal 0.5 is a double, so this one expression is evaluated in fp64
        in = (float)__dmul_rn(__dmul_rn(0.5, (double)__fadd_rn(a1, a2)), (double)(size - 1));
    }
    AxisTap t;
    if (!(in >= 0.0f && in <= sm1)) {  // also catches NaN (the reference would read out of bounds)
        t.lo = -1;
        t.hi = -1;
        t.lerp = 0.0f;
        return t;
    }
    t.lo = (int)floorf(in);
    t.hi = (int)ceilf(in);
    t.lerp = __fsub_rn(in, (float)t.lo);
    return t;
}

// Bilinear blend, cpu/crop_cpu.cpp:107-110.
__device__ __forceinline__ float bilerp(float tl, float tr, float bl, float br, float xl, float yl) {
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
    return __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
}

// The same blend for four channels at once with Blackwell's packed fp32x2 pipe (FADD2 / FFMA2, sm_100+): half the
// FP instructions, and still ONE rounding per reference operation.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2
// into FFMA2 (unlike the scalar .rn forms), so the product is written as fma(a, b, nz) with nz = -0.0f read from
// a kernel parameter - exact (x*y + -0 == RN(x*y), sign of zero included) and opaque to the contraction.
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2_rn(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long sub2_rn(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2_rn(unsigned long long a, unsigned long long b, unsigned long long nz) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
    return r;
}
__device__ __forceinline__ unsigned long long lerp2(unsigned long long a, unsigned long long b, unsigned long long t,
                                                    unsigned long long nz) {
    return add2_rn(a, mul2_rn(sub2_rn(b, a), t, nz));  // a + (b - a) * t, three roundings
}
__device__ __forceinline__ float4 bilerp4(const float4 tl, const float4 tr, const float4 bl, const float4 br, float xl,
                                          float yl, float negzero) {
    const unsigned long long nz = pack2(negzero, negzero), x2 = pack2(xl, xl), y2 = pack2(yl, yl);
    const unsigned long long top_lo = lerp2(pack2(tl.x, tl.y), pack2(tr.x, tr.y), x2, nz);
    const unsigned long long top_hi = lerp2(pack2(tl.z, tl.w), pack2(tr.z, tr.w), x2, nz);
    const unsigned long long bot_lo = lerp2(pack2(bl.x, bl.y), pack2(br.x, br.y), x2, nz);
    const unsigned long long bot_hi = lerp2(pack2(bl.z, bl.w), pack2(br.z, br.w), x2, nz);
    float4 v;
    unpack2(lerp2(top_lo, bot_lo, y2, nz), v.x, v.y);
    unpack2(lerp2(top_hi, bot_hi, y2, nz), v.z, v.w);
    return v;
}

// FPN level thresholds (see host side, roialign.cu: level_thresholds()).  level(q) for
// q = sqrt(h*w) / (224/sqrt(image_area)) is 2 + [q>=t3] + [q>=t4] + [q>=t5]: the exact step
// function of clamp(round_half_even(4 + log2(q)), 2, 5) (model.py:331-338) with correctly rounded log2.
struct LevelRule {
    float denom;  // 224 / sqrt(image_area), fp32
    float t3, t4, t5;
};

__device__ __forceinline__ int roi_level(float y1, float x1, float y2, float x2, const LevelRule& r) {
    const float h = __fsub_rn(y2, y1);
    const float w = __fsub_rn(x2, x1);
    const float q = __fdiv_rn(__fsqrt_rn(__fmul_rn(h, w)), r.denom);
    if (!(q < INFINITY)) return 2;  // NaN / +inf: .int() gives INT_MIN on x86, clamp -> 2
    return 2 + (q >= r.t3) + (q >= r.t4) + (q >= r.t5);
}

// torch.clamp semantics (NaN propagates), data.py:86-92.
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Correctly rounded fp32 exp through fp64 (the oracle's definition of torch.exp, data.py:138-140).
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }

// data.py:124-148 boxes_refine for one box; d = deltas already multiplied by std.
__device__ __forceinline__ void box_refine(const float b[4], const float d[4], float o[4]) {
    float h = __fsub_rn(b[2], b[0]);
    float w = __fsub_rn(b[3], b[1]);
    float cy = __fadd_rn(b[0], __fmul_rn(0.5f, h));
    float cx = __fadd_rn(b[1], __fmul_rn(0.5f, w));
    cy = __fadd_rn(cy, __fmul_rn(d[0], h));
    cx = __fadd_rn(cx, __fmul_rn(d[1], w));
    h = __fmul_rn(h, exp_cr(d[2]));
    w = __fmul_rn(w, exp_cr(d[3]));
    o[0] = __fsub_rn(cy, __fmul_rn(0.5f, h));
    o[1] = __fsub_rn(cx, __fmul_rn(0.5f, w));
    o[2] = __fadd_rn(o[0], h);
    o[3] = __fadd_rn(o[1], w);
}

// IoU >= threshold test with the +1 pixel convention, cpu/nms_cpu.cpp:26, :56-65.
// Boxes are (y1,x1,y2,x2); areas precomputed as ((x2-x1)+1)*((y2-y1)+1).
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// `margin` = |thr| * 1e-6 + 1e-37 (hoist it out of loops).  The decision is that of the correctly rounded quotient
// (NaN -> false, kept, as on the CPU): a ~2-ulp approximate quotient settles it unless it lands within ~8 ulp of the
// threshold or the denominator is outside the range where the approximation holds; only then divide exactly.
// Branch-free up to that rare fallback (disjoint boxes give q = 0 < thr - margin).
__device__ __forceinline__ bool iou_ge_m(const float4 a, float area_a, const float4 b, float area_b, float thr, float margin) {
    const float yy1 = fmaxf(a.x, b.x);
    const float xx1 = fmaxf(a.y, b.y);
    const float yy2 = fminf(a.z, b.z);
    const float xx2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
    const float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
    const float inter = __fmul_rn(w, h);
    const float denom = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    const float q = __fmul_rn(inter, rcp_approx(denom));
    const float ad = fabsf(denom);
    const bool yes = q > __fadd_rn(thr, margin);
    const bool sure = (yes || q < __fsub_rn(thr, margin)) && ad < 1e37f && ad > 1e-30f;
    if (sure) return yes;
    return __fdiv_rn(inter, denom) >= thr;
}

// The same decision with an early exit for disjoint boxes: cheaper where whole warps are disjoint most of the time
// (the one-CTA-per-image detection layer), dearer in the dense 64x64 tiles of the proposal mask kernel.
__device__ __forceinline__ bool iou_ge(const float4 a, float area_a, const float4 b, float area_b, float thr) {
    const float yy1 = fmaxf(a.x, b.x);
    const float xx1 = fmaxf(a.y, b.y);
    const float yy2 = fminf(a.z, b.z);
    const float xx2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
    const float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
    const float inter = __fmul_rn(w, h);
    if (inter == 0.0f && thr > 0.0f) return false;  // 0 / x is 0 (or NaN), never >= a positive threshold
    const float denom = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    if (fabsf(denom) < 1e37f && fabsf(denom) > 1e-30f) {
        const float q = __fdividef(inter, denom);
        const float margin = fabsf(thr) * 1e-6f + 1e-37f;
        if (q > thr + margin) return true;
        if (q < thr - margin) return false;
    }
    return __fdiv_rn(inter, denom) >= thr;
}

__device__ __forceinline__ float box_area_p1(const float4 b) {
    return __fmul_rn(__fadd_rn(__fsub_rn(b.w, b.y), 1.0f), __fadd_rn(__fsub_rn(b.z, b.x), 1.0f));
}

// Order-preserving map float -> uint32 (larger float => larger key).
__device__ __forceinline__ uint32_t float_to_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// 128-bit streaming load (read-only path, do not pollute L1) and vector reduction to global memory.
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 ldg_f4_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ void red_add_f4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ void stg_f4_stream(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace mrcnn
