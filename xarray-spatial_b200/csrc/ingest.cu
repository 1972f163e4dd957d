// ingest.cu -- slope / aspect / curvature / hillshade reading int16 / uint16 / int32 / float64
// rasters DIRECTLY.  The reference casts every input to float32 first
// (`data.astype(np.float32)`, slope.py:58,150) -- an extra full pass over the raster; here the
// CTA-wide TMA pipeline of stencil3.cuh moves the raw elements and the consumer lanes widen / narrow
// them to float32 in registers (same rounding as astype: int -> f32 and f64 -> f32, round to nearest
// even), then run the unchanged float32 operators.  For integer rasters the TMA unit can only
// zero-fill out-of-raster cells, so the loader substitutes NaN from the cell coordinates.
#include <math.h>
#include <type_traits>

#include "surface_ops.cuh"

namespace xrs {

bool make_tensor_map_2d_raw(CUtensorMap *map, const void *base, int64_t pitch_bytes, int64_t H, int64_t W,
                            int dtype, int box_w, int box_h);  // lib_core.cu

// The CTA-wide TMA pipeline of stencil3.cuh with a source element type TS != float: the ring holds the raw
// cells (2-byte cells: 8-cell halos so that box starts stay 16-byte aligned), consumer lanes convert.
template <typename Op, typename TS, int ROWS, int STAGES, int WARPS, int CTAS>
static int launch_ingest(const void *in, int dtype, int64_t in_pitch, const typename Op::Params &prm, float *out,
                         int64_t out_pitch, int64_t H, int64_t W, cudaStream_t stream) {
    static_assert(std::is_same<typename Op::in_t, float>::value, "ingest feeds float32 operators");
    CUtensorMap tmap;
    const bool ok = (W % 4 == 0) && (out_pitch % 16 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                    make_tensor_map_2d_raw(&tmap, in, in_pitch, H, W, dtype, kSubW, ROWS);
    if (!ok) {
        set_error("raster layout not supported by the direct-ingest path (needs 16-byte aligned rows, W %% 4 == 0)");
        return XRS_EUNSUPPORTED;
    }
    OutPtrs<Op> outs;
    outs.p[0] = out;
    outs.pitch_elems = out_pitch / 4;
    constexpr int kPadS = SrcPad<TS>::value;
    constexpr int kTileW = TileShape<WARPS, kPadS>::kTileW;
    constexpr size_t smem = (size_t)STAGES * TileShape<WARPS, kPadS>::kNSub * ROWS * kSubW * sizeof(TS) +
                            (size_t)2 * STAGES * sizeof(uint64_t);
    auto kern = stencil3_tma_kernel<Op, ROWS, STAGES, WARPS, TS>;
    XRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    XRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (WARPS + 1) * 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > CTAS) per_sm = CTAS;
    const int64_t resident = (int64_t)sm_count() * per_sm;
    const TileGeom g = make_tile_geom(H, W, kTileW, ROWS, resident);
    const int64_t n_tasks = (int64_t)g.n_tiles * g.n_segs;
    const int64_t grid = resident < n_tasks ? resident : n_tasks;
    LaunchInfo &li = last_launch_info();
    li.used_tma = 2;  // 2 = direct-ingest TMA kernel
    li.grid = (int)grid; li.block = (WARPS + 1) * 32; li.smem_bytes = (int)smem;
    kern<<<(unsigned)grid, (WARPS + 1) * 32, smem, stream>>>(tmap, prm, outs, g);
    XRS_CUDA(cudaGetLastError());
    return XRS_OK;
}

// ring geometry: 2-byte cells need ROWS % 4 == 0 (128-byte aligned boxes); bytes in flight follow the
// float32 kernels' sweet spot (~65 KB per SM), arithmetic-heavy operators get 16 consumer warps
template <typename Op> struct IngestCfg { static constexpr int kWarps = 16, kCtas = 1; };
template <> struct IngestCfg<SlopeOp> { static constexpr int kWarps = 8, kCtas = 2; };

template <typename Op>
static int dispatch_dtype(const void *in, int dtype, int64_t in_pitch, const typename Op::Params &prm, float *out,
                          int64_t out_pitch, int64_t H, int64_t W, cudaStream_t st) {
    constexpr int W_ = IngestCfg<Op>::kWarps, C_ = IngestCfg<Op>::kCtas;
    switch (dtype) {
        case XRS_I16: return launch_ingest<Op, short, 4, 4, W_, C_>(in, dtype, in_pitch, prm, out, out_pitch, H, W, st);
        case XRS_U16: return launch_ingest<Op, unsigned short, 4, 4, W_, C_>(in, dtype, in_pitch, prm, out, out_pitch, H, W, st);
        case XRS_I32: return launch_ingest<Op, int, 4, 3, W_, C_>(in, dtype, in_pitch, prm, out, out_pitch, H, W, st);
        case XRS_F64: return launch_ingest<Op, double, 2, 3, W_, C_>(in, dtype, in_pitch, prm, out, out_pitch, H, W, st);
    }
    set_error("direct ingest supports int16, uint16, int32 and float64 rasters");
    return XRS_EUNSUPPORTED;
}

}  // namespace xrs

using namespace xrs;

extern "C" int xrs_surface_typed(int op, const void *in, int in_dtype, int64_t in_pitch, float *out,
                                 int64_t out_pitch, int64_t H, int64_t W, const double *p, xrs_stream_t s) {
    if (H <= 0 || W <= 0) return XRS_OK;
    XRS_REQUIRE(in && out, "NULL pointer");
    XRS_REQUIRE(H < (1LL << 31) - 8 && W < (1LL << 31) - 256, "raster dimension too large");
    cudaStream_t st = (cudaStream_t)s;
    switch (op) {
        case XRS_OP_SLOPE: {
            XRS_REQUIRE(p != nullptr, "cell sizes missing");
            const double kx = 1.0 / (8.0 * p[0]), ky = 1.0 / (8.0 * p[1]);
            SlopeOp::Params q;
            q.rxy = kx / ky;
            q.ky2 = (float)(ky * ky);
            return dispatch_dtype<SlopeOp>(in, in_dtype, in_pitch, q, out, out_pitch, H, W, st);
        }
        case XRS_OP_ASPECT: {
            AspectOp::Params q = {0};
            return dispatch_dtype<AspectOp>(in, in_dtype, in_pitch, q, out, out_pitch, H, W, st);
        }
        case XRS_OP_CURVATURE: {
            XRS_REQUIRE(p != nullptr, "cell size missing");
            CurvatureOp::Params q;
            q.k = 100.0 / (p[0] * p[0]);
            return dispatch_dtype<CurvatureOp>(in, in_dtype, in_pitch, q, out, out_pitch, H, W, st);
        }
        case XRS_OP_HILLSHADE: {
            XRS_REQUIRE(p != nullptr, "azimuth / altitude missing");
            const double az = 360.0 - p[0], azr = az * M_PI / 180., altr = p[1] * M_PI / 180., A = azr - M_PI / 2.;
            HillshadeOp::Params q;
            q.s0 = (float)sin(altr);
            q.cy = (float)(0.5 * cos(altr) * cos(A));
            q.cx = (float)(0.5 * cos(altr) * sin(A));
            return dispatch_dtype<HillshadeOp>(in, in_dtype, in_pitch, q, out, out_pitch, H, W, st);
        }
    }
    set_error("xrs_surface_typed serves slope, aspect, curvature and hillshade");
    return XRS_EINVAL;
}
