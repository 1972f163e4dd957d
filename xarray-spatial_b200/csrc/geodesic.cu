// geodesic.cu -- slope / aspect with method='geodesic' (geodesic.py:40-231, slope.py:167-264):
// every 3x3 neighbourhood goes (lat, lon, h) -> ECEF -> local East/North/Up frame of the centre
// cell, gets the curvature correction u += (e^2 + n^2) / (2 R) and a centred least-squares plane
// u = A e + B n; slope = atan(|(A, B)|), aspect = atan2(-A, -B).  All float64, float32 output.
//
// Regular grids (1-D lat per row, 1-D lon per column -- the common case, utils.py:644-649) first
// build two small trig tables (sin/cos lat and the prime-vertical radius N per row, sin/cos lon
// per column), so a cell needs no transcendental except the final atan / atan2: ~300 FP64
// operations per cell, FP64-bound.  Curvilinear grids (2-D lat/lon) evaluate the trig per
// neighbour.  One thread per output cell; the 27 loads hit L1/L2.
#include <math.h>

#include "common.cuh"

namespace xrs {

constexpr double kA2 = 6378137.0 * 6378137.0;
constexpr double kB2 = 6356752.314245 * 6356752.314245;
constexpr double kInv2R = 1.0 / (2.0 * 6370994.884953014);  // geodesic.py:183
constexpr double kDeg2Rad = 3.141592653589793 / 180.0;
constexpr double kRad2Deg = 180.0 / 3.141592653589793;

struct RowTrig { double s, c, n; };   // sin(lat), cos(lat), N(lat)
struct ColTrig { double s, c; };      // sin(lon), cos(lon)

__global__ void geo_tables_kernel(const double *lat, const double *lon, int64_t H, int64_t W, RowTrig *rt,
                                  ColTrig *ct) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < H) {
        const double r = lat[i] * kDeg2Rad, s = sin(r), c = cos(r);
        rt[i].s = s; rt[i].c = c;
        rt[i].n = kA2 / sqrt(kA2 * c * c + kB2 * s * s);
    }
    if (i < W) {
        const double r = lon[i] * kDeg2Rad;
        ct[i].s = sin(r); ct[i].c = cos(r);
    }
}

struct GeoArgs {
    const void *elev;
    int64_t pitch_elems;
    const double *lat, *lon;     // 2-D mode: (H, W) contiguous; 1-D mode: unused
    const RowTrig *rt;
    const ColTrig *ct;
    float *out;
    int64_t out_pitch_elems;
    int64_t H, W;
    double z_factor;
    int want_aspect;
};

template <typename TE, bool GRID2D>
__global__ void __launch_bounds__(256) geodesic_kernel(const __grid_constant__ GeoArgs a) {
    const TE *elev = reinterpret_cast<const TE *>(a.elev);
    const int64_t n = a.H * a.W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t y = i / a.W, x = i % a.W;
        float res = nan_of<float>();
        if (y >= 1 && y < a.H - 1 && x >= 1 && x < a.W - 1) {
            double h9[9];
            bool ok = true;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const double v = (double)elev[(y + dy - 1) * a.pitch_elems + x + dx - 1];
                    h9[dy * 3 + dx] = v;
                    ok = ok && (v == v);
                }
            if (ok) {
                RowTrig r3[3];
                ColTrig c3[3];
                double X[9], Y[9], Z[9];
                if constexpr (!GRID2D) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { r3[d] = a.rt[y + d - 1]; c3[d] = a.ct[x + d - 1]; }
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const RowTrig &r = r3[k / 3];
                        const ColTrig &c = c3[k % 3];
                        const double h = h9[k] * a.z_factor;
                        X[k] = (r.n + h) * r.c * c.c;
                        Y[k] = (r.n + h) * r.c * c.s;
                        Z[k] = (kB2 / kA2 * r.n + h) * r.s;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const int64_t j = (y + k / 3 - 1) * a.W + x + k % 3 - 1;
                        const double la = a.lat[j] * kDeg2Rad, lo = a.lon[j] * kDeg2Rad;
                        const double sl = sin(la), cl = cos(la), so = sin(lo), co = cos(lo);
                        const double N = kA2 / sqrt(kA2 * cl * cl + kB2 * sl * sl);
                        const double h = h9[k] * a.z_factor;
                        X[k] = (N + h) * cl * co;
                        Y[k] = (N + h) * cl * so;
                        Z[k] = (kB2 / kA2 * N + h) * sl;
                        if (k == 4) { r3[1].s = sl; r3[1].c = cl; c3[1].s = so; c3[1].c = co; }
                    }
                }
                const double sin_lat = r3[1].s, cos_lat = r3[1].c, sin_lon = c3[1].s, cos_lon = c3[1].c;
                const double ex = -sin_lon, ey = cos_lon;
                const double nx = -sin_lat * cos_lon, ny = -sin_lat * sin_lon, nz = cos_lat;
                const double ux = cos_lat * cos_lon, uy = cos_lat * sin_lon, uz = sin_lat;
                double e9[9], n9[9], u9[9];
                double me = 0.0, mn = 0.0, mu = 0.0;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const double dx = X[k] - X[4], dy = Y[k] - Y[4], dz = Z[k] - Z[4];
                    const double ek = dx * ex + dy * ey;  // ez = 0
                    const double nk = dx * nx + dy * ny + dz * nz;
                    double uk = dx * ux + dy * uy + dz * uz;
                    uk += (ek * ek + nk * nk) * kInv2R;
                    e9[k] = ek; n9[k] = nk; u9[k] = uk;
                    me += ek; mn += nk; mu += uk;
                }
                me *= (1.0 / 9.0); mn *= (1.0 / 9.0); mu *= (1.0 / 9.0);
                double See = 0.0, Snn = 0.0, Sen = 0.0, Seu = 0.0, Snu = 0.0;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const double de = e9[k] - me, dn = n9[k] - mn, du = u9[k] - mu;
                    See += de * de; Snn += dn * dn; Sen += de * dn; Seu += de * du; Snu += dn * du;
                }
                const double det = See * Snn - Sen * Sen;
                double A = 0.0, B = 0.0;
                if (!(fabs(det) < 1e-30)) {
                    A = (Seu * Snn - Snu * Sen) / det;
                    B = (Snu * See - Seu * Sen) / det;
                }
                const double mag = sqrt(A * A + B * B);
                if (!a.want_aspect) {
                    res = (float)(atan(mag) * kRad2Deg);
                } else if (mag < 1e-7) {
                    res = -1.0f;
                } else {
                    double deg = atan2(-A, -B) * kRad2Deg;
                    if (deg < 0) deg += 360.0;
                    if (deg >= 360.0) deg -= 360.0;
                    res = (float)deg;
                }
            }
        }
        a.out[y * a.out_pitch_elems + x] = res;
    }
}

}  // namespace xrs

using namespace xrs;

extern "C" int xrs_geodesic(const void *elev, int elev_dtype, int64_t elev_pitch, const double *lat,
                            const double *lon, int latlon_2d, float *out, int64_t out_pitch, int64_t H, int64_t W,
                            double z_factor, int want_aspect, xrs_stream_t s) {
    if (H <= 0 || W <= 0) return XRS_OK;
    XRS_REQUIRE(elev && lat && lon && out, "NULL pointer");
    XRS_REQUIRE(elev_dtype == XRS_F32 || elev_dtype == XRS_F64, "elevation must be float32 or float64");
    const int esz = elev_dtype == XRS_F32 ? 4 : 8;
    XRS_REQUIRE(elev_pitch % esz == 0 && elev_pitch >= W * esz, "bad elevation pitch");
    XRS_REQUIRE(out_pitch % 4 == 0 && out_pitch >= W * 4, "bad output pitch");
    cudaStream_t st = (cudaStream_t)s;
    GeoArgs a;
    a.elev = elev; a.pitch_elems = elev_pitch / esz; a.lat = lat; a.lon = lon; a.rt = nullptr; a.ct = nullptr;
    a.out = out; a.out_pitch_elems = out_pitch / 4; a.H = H; a.W = W; a.z_factor = z_factor;
    a.want_aspect = want_aspect;
    void *scratch = nullptr;
    if (!latlon_2d) {
        XRS_CUDA(cudaMallocAsync(&scratch, (size_t)H * sizeof(RowTrig) + (size_t)W * sizeof(ColTrig), st));
        RowTrig *rt = (RowTrig *)scratch;
        ColTrig *ct = (ColTrig *)(rt + H);
        const int64_t m = H > W ? H : W;
        geo_tables_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(lat, lon, H, W, rt, ct);
        a.rt = rt; a.ct = ct;
    }
    int64_t grid = (H * W + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    if (elev_dtype == XRS_F32) {
        if (latlon_2d) geodesic_kernel<float, true><<<(unsigned)grid, 256, 0, st>>>(a);
        else geodesic_kernel<float, false><<<(unsigned)grid, 256, 0, st>>>(a);
    } else {
        if (latlon_2d) geodesic_kernel<double, true><<<(unsigned)grid, 256, 0, st>>>(a);
        else geodesic_kernel<double, false><<<(unsigned)grid, 256, 0, st>>>(a);
    }
    cudaError_t e = cudaGetLastError();
    if (scratch) cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) return cuda_fail(e, "geodesic_kernel launch");
    return XRS_OK;
}
