// surface.cu -- C-ABI entry points of the 3x3 family (slope, aspect, curvature, hillshade,
// fused suite, focal.mean).  Argument checking and launch geometry live in stencil3.cuh.
#include <math.h>

#include "surface_ops.cuh"

using namespace xrs;

// TMA ring geometry: ROWS rows per box, STAGES boxes in flight per warp.
// f32: 4 x 136 x 4 B = 2176 B per stage, 4 stages, 8 warps -> 68 KiB per CTA (3 CTAs / SM).
constexpr int kRowsF32 = 4, kStagesF32 = 4;
constexpr int kRowsF64 = 2, kStagesF64 = 4;

static SlopeOp::Params slope_params(double cellsize_x, double cellsize_y) {
    const double kx = 1.0 / (8.0 * cellsize_x), ky = 1.0 / (8.0 * cellsize_y);
    SlopeOp::Params p;
    p.rxy = kx / ky;
    p.ky2 = (float)(ky * ky);
    return p;
}

static HillshadeOp::Params hillshade_params(double azimuth, double angle_altitude) {
    // hillshade.py:23-27: azimuth = 360 - azimuth; rad conversions in Python float (f64)
    const double az = 360.0 - azimuth;
    const double azimuthrad = az * M_PI / 180.;
    const double altituderad = angle_altitude * M_PI / 180.;
    const double A = azimuthrad - M_PI / 2.;
    HillshadeOp::Params p;
    p.s0 = (float)sin(altituderad);
    p.cy = (float)(0.5 * cos(altituderad) * cos(A));
    p.cx = (float)(0.5 * cos(altituderad) * sin(A));
    return p;
}

// 3x3 kernels of convolve_2d take the warp-strip path (called from conv.cu)
int xrs_conv3_strip(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H, int64_t W,
                    const double *kernel, cudaStream_t s) {
    Conv3Op::Params p;
    for (int i = 0; i < 9; ++i) p.w[i] = kernel[i];
    float *outs[1] = {out};
    return launch_stencil3<Conv3Op, kRowsF32, kStagesF32>(in, in_pitch, p, outs, out_pitch, H, W, s);
}

template <typename T, typename TOUT, int ROWS, int STAGES>
static int focal_mean_impl(const T *in, int64_t in_pitch, TOUT *out, int64_t out_pitch, int64_t H, int64_t W,
                           const double *excludes, int n_ex, xrs_stream_t s) {
    using Op = FocalMeanOp<T, TOUT, false>;
    using OpEx = FocalMeanOp<T, TOUT, true>;
    static_assert(sizeof(typename Op::Params) == sizeof(typename OpEx::Params), "same parameter block");
    XRS_REQUIRE(n_ex >= 0 && n_ex <= Op::kMaxEx, "at most 8 exclude values are supported");
    XRS_REQUIRE(n_ex == 0 || excludes != nullptr, "excludes is NULL");
    typename OpEx::Params p;
    p.n_ex = 0;
    p.ex_nan = 0;
    for (int i = 0; i < Op::kMaxEx; ++i) p.ex[i] = 0.0;
    for (int i = 0; i < n_ex; ++i) {
        if (excludes[i] != excludes[i]) p.ex_nan = 1;
        else p.ex[p.n_ex++] = excludes[i];
    }
    TOUT *outs[1] = {out};
    if (p.n_ex > 0 || !p.ex_nan) return launch_stencil3<OpEx, ROWS, STAGES>(in, in_pitch, p, outs, out_pitch, H, W, (cudaStream_t)s);
    typename Op::Params q;
    memcpy(&q, &p, sizeof(q));
    return launch_stencil3<Op, ROWS, STAGES>(in, in_pitch, q, outs, out_pitch, H, W, (cudaStream_t)s);
}

extern "C" {

int xrs_slope_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H, int64_t W,
                  double cellsize_x, double cellsize_y, xrs_stream_t s) {
    const SlopeOp::Params p = slope_params(cellsize_x, cellsize_y);
    float *outs[1] = {out};
    return launch_stencil3<SlopeOp, kRowsF32, kStagesF32>(in, in_pitch, p, outs, out_pitch, H, W,
                                                          (cudaStream_t)s);
}

int xrs_aspect_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H, int64_t W,
                   xrs_stream_t s) {
    AspectOp::Params p = {0};
    float *outs[1] = {out};
    return launch_stencil3<AspectOp, kRowsF32, kStagesF32>(in, in_pitch, p, outs, out_pitch, H, W,
                                                           (cudaStream_t)s);
}

int xrs_curvature_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H,
                      int64_t W, double cellsize, xrs_stream_t s) {
    CurvatureOp::Params p;
    p.k = 100.0 / (cellsize * cellsize);
    float *outs[1] = {out};
    return launch_stencil3<CurvatureOp, kRowsF32, kStagesF32>(in, in_pitch, p, outs, out_pitch, H, W,
                                                              (cudaStream_t)s);
}

int xrs_hillshade_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H,
                      int64_t W, double azimuth, double angle_altitude, xrs_stream_t s) {
    const HillshadeOp::Params p = hillshade_params(azimuth, angle_altitude);
    float *outs[1] = {out};
    return launch_stencil3<HillshadeOp, kRowsF32, kStagesF32>(in, in_pitch, p, outs, out_pitch, H, W,
                                                              (cudaStream_t)s);
}

int xrs_surface_suite_f32(const float *in, int64_t in_pitch, float *slope_out, float *aspect_out,
                          float *curvature_out, float *hillshade_out, int64_t out_pitch, int64_t H,
                          int64_t W, double cellsize_x, double cellsize_y, double azimuth,
                          double angle_altitude, xrs_stream_t s) {
    SuiteOp::Params p;
    p.slope = slope_params(cellsize_x, cellsize_y);
    const double cs = (cellsize_x + cellsize_y) / 2;  // curvature.py:234
    p.curv.k = 100.0 / (cs * cs);
    p.hill = hillshade_params(azimuth, angle_altitude);
    float *outs[4] = {slope_out, aspect_out, curvature_out, hillshade_out};
    return launch_stencil3<SuiteOp, kRowsF32, kStagesF32>(in, in_pitch, p, outs, out_pitch, H, W,
                                                          (cudaStream_t)s);
}

int xrs_focal_mean_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H,
                       int64_t W, const double *excludes, int n_ex, xrs_stream_t s) {
    return focal_mean_impl<float, float, kRowsF32, kStagesF32>(in, in_pitch, out, out_pitch, H, W, excludes, n_ex, s);
}
int xrs_focal_mean_f64(const double *in, int64_t in_pitch, double *out, int64_t out_pitch, int64_t H,
                       int64_t W, const double *excludes, int n_ex, xrs_stream_t s) {
    return focal_mean_impl<double, double, kRowsF64, kStagesF64>(in, in_pitch, out, out_pitch, H, W, excludes,
                                                                 n_ex, s);
}
int xrs_focal_mean_f32_f64(const float *in, int64_t in_pitch, double *out, int64_t out_pitch, int64_t H,
                           int64_t W, const double *excludes, int n_ex, xrs_stream_t s) {
    return focal_mean_impl<float, double, kRowsF32, kStagesF32>(in, in_pitch, out, out_pitch, H, W, excludes,
                                                                n_ex, s);
}

}  // extern "C"
