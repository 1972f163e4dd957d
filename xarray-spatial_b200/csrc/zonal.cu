// zonal.cu -- per-zone partial statistics for zonal.stats (zonal.py:280-332, 422-667).
//
// The reference sorts the whole raster by zone id (np.argsort, O(N log N)) and then runs
// NumPy reductions per zone.  Here ONE streaming pass (8 B/cell for f32 values + i32 zones,
// HBM-bound) accumulates, per zone, count / sum / sum of squares (about a per-zone pivot) /
// min / max:
//   * each thread owns runs of 4 consecutive cells (128-bit loads) and keeps a private
//     accumulator for the zone it is currently in; the accumulator is flushed to a
//     shared-memory table (privatised per CTA) only when the zone changes;
//   * zone value -> dense index by a cached lookup: last zone first, then a small direct
//     table for integer ids in a compact range, else binary search of the sorted id list
//     (both in shared memory);
//   * at the end every CTA merges the zones it touched into the global partial arrays with
//     atomics.  Partials from several GPUs combine with sum / min / max (NCCL AllReduce).
// Counts are exact integers; sums are float64.
#include <math.h>

#include "common.cuh"

namespace xrs {

constexpr int kZonalThreads = 256;
constexpr int kMaxSharedZones = 2048;   // table rows kept in shared memory
constexpr int kLutSize = 8192;          // direct table for compact integer ids

struct ZAcc {
    double s1, s2, mn, mx;
    unsigned int cnt;
};

__device__ __forceinline__ void atomic_min_f64(double *addr, double v) {
    unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *a;
    while (v < __longlong_as_double((long long)old)) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_f64(double *addr, double v) {
    unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *a;
    while (v > __longlong_as_double((long long)old)) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

struct ZonalArgs {
    const void *values;
    const void *zones;
    int64_t n;
    int64_t W;               // row length (n = H * W): the scan walks down 128-column strips
    const double *zone_ids;  // sorted unique, device
    const double *pivot;     // device, nz
    int nz;
    int has_nodata;
    double nodata;
    int use_lut;
    long long lut_base;      // LUT index = zone - lut_base
    long long *count;
    double *sum, *sumsq, *vmin, *vmax;
};

template <typename T> struct Quad { T v[4]; };

template <typename T> __device__ __forceinline__ Quad<T> load_quad(const T *p, int64_t i, int64_t n, bool aligned) {
    Quad<T> q;
    if (aligned && i + 4 <= n) {
        if constexpr (sizeof(T) == 4) {
            const int4 r = __ldcs(reinterpret_cast<const int4 *>(p + i));
            memcpy(&q.v[0], &r, 16);
        } else {
            const int4 r0 = __ldcs(reinterpret_cast<const int4 *>(p + i));
            const int4 r1 = __ldcs(reinterpret_cast<const int4 *>(p + i + 2));
            memcpy(&q.v[0], &r0, 16);
            memcpy(&q.v[2], &r1, 16);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) q.v[k] = (i + k < n) ? p[i + k] : T(0);
    }
    return q;
}

template <typename VT, typename ZT>
__global__ void __launch_bounds__(kZonalThreads) zonal_kernel(const __grid_constant__ ZonalArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nzs = a.nz <= kMaxSharedZones ? a.nz : 0;  // 0: table lives in global memory only
    double *s_ids = reinterpret_cast<double *>(smem_raw);
    double *s_piv = s_ids + nzs;
    double *s_s1 = s_piv + nzs;
    double *s_s2 = s_s1 + nzs;
    double *s_mn = s_s2 + nzs;
    double *s_mx = s_mn + nzs;
    unsigned int *s_cnt = reinterpret_cast<unsigned int *>(s_mx + nzs);
    short *s_lut = reinterpret_cast<short *>(s_cnt + nzs);

    for (int i = threadIdx.x; i < nzs; i += blockDim.x) {
        s_ids[i] = a.zone_ids[i];
        s_piv[i] = a.pivot[i];
        s_s1[i] = 0.0;
        s_s2[i] = 0.0;
        s_mn[i] = INFINITY;
        s_mx[i] = -INFINITY;
        s_cnt[i] = 0u;
    }
    if (a.use_lut && nzs) {
        for (int i = threadIdx.x; i < kLutSize; i += blockDim.x) s_lut[i] = (short)-1;
        __syncthreads();
        for (int i = threadIdx.x; i < nzs; i += blockDim.x) {
            const long long k = (long long)s_ids[i] - a.lut_base;
            if (k >= 0 && k < kLutSize) s_lut[k] = (short)i;
        }
    }
    __syncthreads();

    const double *ids = nzs ? s_ids : a.zone_ids;
    const double *piv = nzs ? s_piv : a.pivot;
    const VT *values = reinterpret_cast<const VT *>(a.values);
    const ZT *zones = reinterpret_cast<const ZT *>(a.zones);
    const bool v_al = (reinterpret_cast<uintptr_t>(values) & 15) == 0;
    const bool z_al = (reinterpret_cast<uintptr_t>(zones) & 15) == 0;

    int cur = -1;       // dense index of the cached zone (-1: none / not a requested zone)
    ZT cur_z = ZT(0);
    bool have = false;  // cur_z valid
    ZAcc acc = {0.0, 0.0, INFINITY, -INFINITY, 0u};
    double cur_p = 0.0;

    auto flush = [&]() {
        if (cur >= 0 && acc.cnt) {
            if (nzs) {
                atomicAdd(&s_cnt[cur], acc.cnt);
                atomicAdd(&s_s1[cur], acc.s1);
                atomicAdd(&s_s2[cur], acc.s2);
                atomic_min_f64(&s_mn[cur], acc.mn);
                atomic_max_f64(&s_mx[cur], acc.mx);
            } else {
                atomicAdd(reinterpret_cast<unsigned long long *>(&a.count[cur]), (unsigned long long)acc.cnt);
                atomicAdd(&a.sum[cur], acc.s1);
                atomicAdd(&a.sumsq[cur], acc.s2);
                atomic_min_f64(&a.vmin[cur], acc.mn);
                atomic_max_f64(&a.vmax[cur], acc.mx);
            }
        }
        acc.s1 = acc.s2 = 0.0;
        acc.mn = INFINITY;
        acc.mx = -INFINITY;
        acc.cnt = 0u;
    };
    auto lookup = [&](ZT z) -> int {
        const double zd = (double)z;
        if (!(zd == zd) || isinf(zd)) return -1;  // non-finite zones are ignored (zonal.py:290)
        if (a.use_lut && nzs) {
            const long long k = (long long)z - a.lut_base;
            return (k >= 0 && k < kLutSize) ? (int)s_lut[k] : -1;
        }
        int lo = 0, hi = a.nz - 1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const double m = ids[mid];
            if (m == zd) return mid;
            if (m < zd) lo = mid + 1; else hi = mid - 1;
        }
        return -1;
    };

    // column-strip traversal (see zonal_hash.cu): a warp walks down a 128-column strip, so a
    // lane's run of equal zone ids is long and flushes are rare
    const int lane = threadIdx.x & 31;
    const int64_t H = a.n / a.W;
    const int64_t n_strips = (a.W + 127) / 128;
    constexpr int kSegRows = 256, kUnroll = 4;
    const int64_t n_segs = (H + kSegRows - 1) / kSegRows;
    const int64_t n_tasks = n_strips * n_segs;
    const int64_t warps_total = (int64_t)gridDim.x * (kZonalThreads / 32);
    const bool row_vec = v_al && z_al && (a.W % 4 == 0);
    for (int64_t task = (int64_t)blockIdx.x * (kZonalThreads / 32) + (threadIdx.x >> 5); task < n_tasks;
         task += warps_total) {
        const int64_t seg = task / n_strips, strip = task % n_strips;
        const int64_t x = strip * 128 + 4 * lane;
        const int64_t y0 = seg * kSegRows, y1 = min(y0 + (int64_t)kSegRows, H);
        const int nv = (int)max((int64_t)0, min((int64_t)4, a.W - x));  // 0: lane past the right edge
        for (int64_t y = y0; y < y1; y += kUnroll) {
            Quad<VT> vq[kUnroll];
            Quad<ZT> zq[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t i0 = (y + u) * a.W + x;
                const int nvu = (y + u < y1) ? nv : 0;
                vq[u] = load_quad<VT>(values, i0, i0 + nvu, row_vec);
                zq[u] = load_quad<ZT>(zones, i0, i0 + nvu, row_vec);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int nvu = (y + u < y1) ? nv : 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool live = k < nvu;
                    const ZT zk = zq[u].v[k];
                    if (live && (!have || !(zk == cur_z))) {
                        flush();
                        cur = lookup(zk);
                        cur_z = zk;
                        have = (zk == zk);
                        cur_p = cur >= 0 ? piv[cur] : 0.0;
                    }
                    const double xv = (double)vq[u].v[k];
                    // finite and != nodata (zonal.py:159); branch-free accumulate
                    const bool ok = live && cur >= 0 && (fabs(xv) <= 1.7976931348623157e308) &&
                                    !(a.has_nodata && xv == a.nodata);
                    const double d = ok ? xv - cur_p : 0.0;
                    acc.s1 += d;
                    acc.s2 = fma(d, d, acc.s2);
                    acc.mn = ok ? fmin(acc.mn, xv) : acc.mn;
                    acc.mx = ok ? fmax(acc.mx, xv) : acc.mx;
                    acc.cnt += ok ? 1u : 0u;
                }
                __syncwarp();  // lanes that took the (rare) flush path rejoin the warp here
            }
        }
    }
    flush();
    __syncthreads();
    for (int i = threadIdx.x; i < nzs; i += blockDim.x) {
        const unsigned int c = s_cnt[i];
        if (c) {
            atomicAdd(reinterpret_cast<unsigned long long *>(&a.count[i]), (unsigned long long)c);
            atomicAdd(&a.sum[i], s_s1[i]);
            atomicAdd(&a.sumsq[i], s_s2[i]);
            atomic_min_f64(&a.vmin[i], s_mn[i]);
            atomic_max_f64(&a.vmax[i], s_mx[i]);
        }
    }
}

__global__ void zonal_init_kernel(long long *count, double *sum, double *sumsq, double *vmin, double *vmax, int nz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nz) {
        count[i] = 0;
        sum[i] = 0.0;
        sumsq[i] = 0.0;
        vmin[i] = INFINITY;
        vmax[i] = -INFINITY;
    }
}

template <typename VT, typename ZT> static int launch_zonal(const ZonalArgs &a, cudaStream_t s) {
    const int nzs = a.nz <= kMaxSharedZones ? a.nz : 0;
    const size_t smem = (size_t)nzs * (6 * sizeof(double) + sizeof(unsigned int)) +
                        (a.use_lut && nzs ? kLutSize * sizeof(short) : 0) + 16;
    auto kern = zonal_kernel<VT, ZT>;
    XRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sm_count() * per_sm;
    const int64_t n_tasks = ((a.W + 127) / 128) * ((a.n / a.W + 255) / 256);
    const int64_t need = (n_tasks + kZonalThreads / 32 - 1) / (kZonalThreads / 32);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kZonalThreads, smem, s>>>(a);
    XRS_CUDA(cudaGetLastError());
    return XRS_OK;
}

}  // namespace xrs

using namespace xrs;

extern "C" {

int xrs_zonal_init(int64_t *count, double *sum, double *sumsq, double *vmin, double *vmax, int nz,
                   xrs_stream_t s) {
    if (nz <= 0) return XRS_OK;
    XRS_REQUIRE(count && sum && sumsq && vmin && vmax, "NULL output pointer");
    zonal_init_kernel<<<(nz + 255) / 256, 256, 0, (cudaStream_t)s>>>((long long *)count, sum, sumsq, vmin, vmax, nz);
    XRS_CUDA(cudaGetLastError());
    return XRS_OK;
}

// lut_base / use_lut are derived here from a host copy of the id range passed through
// `nodata`-independent arguments: the caller gives the ids on the device only, so the range
// test is done by the caller and communicated with xrs_zonal_partials_ex (below);
// xrs_zonal_partials is the portable entry point (binary search).
int xrs_zonal_partials_ex(const void *values, int values_dtype, const void *zones, int zones_dtype, int64_t n,
                          const double *zone_ids, int nz, const double *pivot, int has_nodata, double nodata,
                          int use_lut, int64_t lut_base, int64_t row_len, int64_t *count, double *sum,
                          double *sumsq, double *vmin, double *vmax, xrs_stream_t s) {
    if (n <= 0 || nz <= 0) return XRS_OK;
    XRS_REQUIRE(row_len >= 1 && n % row_len == 0, "n must be a multiple of row_len");
    XRS_REQUIRE(values && zones && zone_ids && pivot && count && sum && sumsq && vmin && vmax, "NULL pointer");
    XRS_REQUIRE(values_dtype == XRS_F32 || values_dtype == XRS_F64, "values must be float32 or float64");
    XRS_REQUIRE(zones_dtype >= XRS_F32 && zones_dtype <= XRS_I64, "unknown zones dtype");
    if (zones_dtype == XRS_F32 || zones_dtype == XRS_F64) use_lut = 0;
    ZonalArgs a;
    a.values = values; a.zones = zones; a.n = n; a.W = row_len; a.zone_ids = zone_ids; a.pivot = pivot; a.nz = nz;
    a.has_nodata = has_nodata; a.nodata = nodata; a.use_lut = use_lut; a.lut_base = lut_base;
    a.count = (long long *)count; a.sum = sum; a.sumsq = sumsq; a.vmin = vmin; a.vmax = vmax;
    cudaStream_t st = (cudaStream_t)s;
#define XRS_Z(VT)                                                          \
    switch (zones_dtype) {                                                 \
        case XRS_I32: return launch_zonal<VT, int>(a, st);                 \
        case XRS_I64: return launch_zonal<VT, long long>(a, st);           \
        case XRS_F32: return launch_zonal<VT, float>(a, st);               \
        default: return launch_zonal<VT, double>(a, st);                   \
    }
    if (values_dtype == XRS_F32) { XRS_Z(float) } else { XRS_Z(double) }
#undef XRS_Z
}

int xrs_zonal_partials(const void *values, int values_dtype, const void *zones, int zones_dtype, int64_t n,
                       const double *zone_ids, int nz, const double *pivot, int has_nodata, double nodata,
                       int64_t *count, double *sum, double *sumsq, double *vmin, double *vmax, xrs_stream_t s) {
    return xrs_zonal_partials_ex(values, values_dtype, zones, zones_dtype, n, zone_ids, nz, pivot, has_nodata,
                                 nodata, 0, 0, n, count, sum, sumsq, vmin, vmax, s);
}

}  // extern "C"
