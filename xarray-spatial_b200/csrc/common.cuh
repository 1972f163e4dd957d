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
    int used_tma;    // 0 cp.async strip kernel, 1 TMA strip kernel, 2 direct-ingest, 3 summed-area box convolve,
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

// degrees(atan(sqrt(p))) for p >= 0 (NaN propagates).  One MUFU.RSQ, no division:
// for p > 1 uses atan(s) = pi/2 - atan(1/s) with 1/s = rsqrt(p).  p below 1e-30 (slope
// below 6e-14 degrees) evaluates as p * 1e15, far under any tolerance.
__device__ __forceinline__ float atan_sqrt_deg(float p) {
    const float r = rsqrt_approx(fmaxf(p, 1e-30f));
    const float s = p * r;            // sqrt(p); NaN stays NaN
    const bool big = p > 1.0f;        // false for NaN -> the NaN in `s` propagates
    const float a = big ? r : s;
    const float t = a * atan_poly01<1>(a * a);
    return big ? (1.57079632679489662f * 57.29578f - t) : t;
}

// Compass aspect from the exact Horn sums X = 8 dz_dx, Y = 8 dz_dy (aspect.py:74-88):
// the reference's (90 - atan2(Y, -X) deg) folded to [0, 360) is atan2(u, v) with u = -X, v = Y,
// folded to [0, 360).  One octant reduction (ratio of the smaller to the larger magnitude,
// one MUFU.RCP, degree-7 polynomial already scaled to degrees), then compass = K + sigma*base
// with (K, sigma) picked per octant.  Evaluating the compass angle directly keeps full
// relative accuracy near 0 degrees, where `90 - theta` would cancel.  Flat cells (both
// sums zero) give -1; NaN propagates (selects, not fmin/fmax, pick the operands).
__device__ __forceinline__ float compass_deg(float u, float v) {
    const float au = fabsf(u), av = fabsf(v);
    const bool swap = au > av;                // closer to the +-u axis (east / west)
    const float mx = swap ? au : av, mn = swap ? av : au;
    const float t = mn * rcp_approx(mx);
    const float base = t * atan_poly01<1>(t * t);                 // [0, 45] degrees
    // sigma = sign(u) * sign(v), negated in the swapped octants: flip base's sign bit
    const unsigned sgn = ((__float_as_uint(u) ^ __float_as_uint(v)) & 0x80000000u) ^ (swap ? 0x80000000u : 0u);
    const float sb = __uint_as_float(__float_as_uint(base) ^ sgn);
    const float k_ns = (v > 0.0f) ? ((u < 0.0f) ? 360.0f : 0.0f) : 180.0f;
    const float k_ew = (u > 0.0f) ? 90.0f : 270.0f;
    const float r = (swap ? k_ew : k_ns) + sb;
    return (mx == 0.0f) ? -1.0f : r;
}

}  // namespace xrs
