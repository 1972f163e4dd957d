// common.cuh -- shared host/device helpers for libxrs_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/xrs_b200.h"

namespace xrs {

// ----------------------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

struct LaunchInfo {  // for tests / profiling: what the last launch on this thread chose
    int used_tma;    // 0 cp.async strip kernel, 1 TMA strip kernel, 2 direct-ingest, 3 running-box kernel,
                     // 4 generic tiled convolve, 5 bounds-checked fallback
    int grid, block, smem_bytes;
};
LaunchInfo &last_launch_info();

#define XRS_CUDA(call)                                              \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return ::xrs::cuda_fail(_e, #call);  \
    } while (0)

#define XRS_REQUIRE(cond, msg)                     \
    do {                                           \
        if (!(cond)) {                             \
            ::xrs::set_error("%s", msg);           \
            return XRS_EINVAL;                     \
        }                                          \
    } while (0)

// cached per-device properties
int sm_count(int device = -1);

// Encodes a 2-D tiled tensor map over a row-major (H, W) raster of `elem_bytes` elements
// with out-of-bounds fill = NaN (the raster-edge semantics of the reference's
// `boundary=np.nan` overlap, slope.py:94-97).  Returns false if TMA cannot describe it
// (base/pitch not 16-byte aligned, ...); callers then use the direct-load kernels.
bool make_tensor_map_2d(CUtensorMap *map, const void *base, int64_t pitch_bytes, int64_t H,
                        int64_t W, int elem_bytes, int box_w, int box_h);

// ----------------------------------------------------------------------------- device PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "XRS_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra XRS_DONE_%=;\n"
        "bra XRS_WAIT_%=;\n"
        "XRS_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA: 2-D tiled bulk tensor load global -> shared, completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// same, with an L2 eviction-priority hint (0: evict_first for streamed inputs, 1: evict_last)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <typename T> __device__ __forceinline__ T nan_of();
template <> __device__ __forceinline__ float nan_of<float>() { return __int_as_float(0x7fc00000); }
template <> __device__ __forceinline__ double nan_of<double>() {
    return __longlong_as_double(0x7ff8000000000000LL);
}

__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// atan(a)/a on a in [0,1] as a degree-7 polynomial in z = a*a (minimax, rel err 1e-7;
// 2.5e-7 evaluated in f32).  The reference evaluates np.arctan in f64 and rounds to f32;
// the parity bar is 1e-5 relative.  SCALE folds a unit conversion into the coefficients.
template <int DEG> __device__ __forceinline__ float atan_poly01(float z) {
    constexpr float k = DEG ? 57.29578f : 1.0f;  // slope.py:75 uses the literal 57.29578
    float p = -4.693183854e-03f * k;
    p = fmaf(p, z, 2.425208207e-02f * k);
    p = fmaf(p, z, -5.948595430e-02f * k);
    p = fmaf(p, z, 9.914263125e-02f * k);
    p = fmaf(p, z, -1.401947061e-01f * k);
    p = fmaf(p, z, 1.996972220e-01f * k);
    p = fmaf(p, z, -3.333199064e-01f * k);
    p = fmaf(p, z, 9.999999010e-01f * k);
    return p;
}

// Packed pairs (Blackwell fma.rn.f32x2 / mul / add: one instruction, two cells; each half rounds
// exactly like the scalar instruction, so results do not depend on how cells are paired).
__device__ __forceinline__ float2 splat2(float c) { return make_float2(c, c); }
template <int DEG> __device__ __forceinline__ float2 atan_poly01x2(float2 z) {
    constexpr float k = DEG ? 57.29578f : 1.0f;
    float2 p = splat2(-4.693183854e-03f * k);
    p = __ffma2_rn(p, z, splat2(2.425208207e-02f * k));
    p = __ffma2_rn(p, z, splat2(-5.948595430e-02f * k));
    p = __ffma2_rn(p, z, splat2(9.914263125e-02f * k));
    p = __ffma2_rn(p, z, splat2(-1.401947061e-01f * k));
    p = __ffma2_rn(p, z, splat2(1.996972220e-01f * k));
    p = __ffma2_rn(p, z, splat2(-3.333199064e-01f * k));
    p = __ffma2_rn(p, z, splat2(9.999999010e-01f * k));
    return p;
}

// degrees(atan(sqrt(p))) for p >= 0 (NaN propagates), two cells at a time.  One MUFU.RSQ per cell,
// no division: for p > 1 uses atan(s) = pi/2 - atan(1/s) with 1/s = rsqrt(p), folded into the last
// FMA as  a * poly(a^2) + off  with (a, off) = (s, 0) or (-1/s, 90 deg).  p below 1e-30 (slope below
// 6e-14 degrees) evaluates as p * 1e15, far under any tolerance.
__device__ __forceinline__ void atan_sqrt_sel(float p, float &a, float &off) {
    const float r = rsqrt_approx(fmaxf(p, 1e-30f));
    const bool big = p > 1.0f;          // false for NaN -> a = p * r = NaN propagates
    a = big ? -r : p * r;               // +-atan argument in [0, 1]
    off = big ? (1.57079632679489662f * 57.29578f) : 0.0f;
}
__device__ __forceinline__ float2 atan_sqrt_deg2(float2 p) {
    const float rx = rsqrt_approx(fmaxf(p.x, 1e-30f)), ry = rsqrt_approx(fmaxf(p.y, 1e-30f));
    const float2 s = __fmul2_rn(p, make_float2(rx, ry));
    const bool bx = p.x > 1.0f, by = p.y > 1.0f;
    const float2 a = make_float2(bx ? -rx : s.x, by ? -ry : s.y);
    constexpr float kQ = 1.57079632679489662f * 57.29578f;
    const float2 off = make_float2(bx ? kQ : 0.0f, by ? kQ : 0.0f);
    return __ffma2_rn(a, atan_poly01x2<1>(__fmul2_rn(a, a)), off);
}
__device__ __forceinline__ float atan_sqrt_deg(float p) {
    float a, off;
    atan_sqrt_sel(p, a, off);
    return fmaf(a, atan_poly01<1>(a * a), off);
}

// Compass aspect from the exact Horn sums X = 8 dz_dx, Y = 8 dz_dy (aspect.py:74-88):
// the reference's (90 - atan2(Y, -X) deg) folded to [0, 360) is atan2(u, v) with u = -X, v = Y,
// folded to [0, 360).  One octant reduction (ratio t of the smaller to the larger magnitude,
// one MUFU.RCP, degree-7 polynomial already scaled to degrees), then compass = K + sigma*atan(t)
// with (K, sigma) picked per octant; sigma is applied to t's sign bit (the polynomial is even in t),
// so the tail is one FMA:  ts * poly(ts^2) + K.  Evaluating the compass angle directly keeps full
// relative accuracy near 0 degrees, where `90 - theta` would cancel.  Flat cells (both sums zero)
// give -1; NaN propagates (selects, not fmin/fmax, pick the operands; the flat test is NaN-safe:
// a NaN sum next to a zero sum is NaN, like atan2(0, NaN) in the reference).
struct CompassPre {
    float ts, K;
    bool flat;
};
// NaN-propagating max / min (FMNMX.NAN): a NaN operand gives NaN, unlike fmaxf / fminf
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ CompassPre compass_pre(float u, float v) {
    const float au = fabsf(u), av = fabsf(v);
    const bool swap = au > av;                // closer to the +-u axis (east / west)
    const float mx = max_nan(au, av), mn = min_nan(au, av);   // a NaN sum makes both NaN
    const float t = mn * rcp_approx(mx);
    // sigma = sign(u) * sign(v), negated in the swapped octants: flip t's sign bit
    const unsigned sgn = ((__float_as_uint(u) ^ __float_as_uint(v)) & 0x80000000u) ^ (swap ? 0x80000000u : 0u);
    CompassPre c;
    c.ts = __uint_as_float(__float_as_uint(t) ^ sgn);
    const float k_ns = (v > 0.0f) ? ((u < 0.0f) ? 360.0f : 0.0f) : 180.0f;
    const float k_ew = (u > 0.0f) ? 90.0f : 270.0f;
    c.K = swap ? k_ew : k_ns;
    c.flat = mx == 0.0f;                       // false for NaN: atan2(0, NaN) is NaN in the reference, not "flat"
    return c;
}
__device__ __forceinline__ float compass_deg(float u, float v) {
    const CompassPre c = compass_pre(u, v);
    const float r = fmaf(c.ts, atan_poly01<1>(c.ts * c.ts), c.K);
    return c.flat ? -1.0f : r;
}
__device__ __forceinline__ float2 compass_deg2(float2 u, float2 v) {
    const CompassPre c0 = compass_pre(u.x, v.x), c1 = compass_pre(u.y, v.y);
    const float2 ts = make_float2(c0.ts, c1.ts);
    const float2 r = __ffma2_rn(ts, atan_poly01x2<1>(__fmul2_rn(ts, ts)), make_float2(c0.K, c1.K));
    return make_float2(c0.flat ? -1.0f : r.x, c1.flat ? -1.0f : r.y);
}

}  // namespace xrs
