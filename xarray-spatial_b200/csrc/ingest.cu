// ingest.cu -- slope / aspect / curvature / hillshade reading int16 / uint16 / int32 / float64
// rasters DIRECTLY.  The reference casts every input to float32 first
// (`data.astype(np.float32)`, slope.py:58,150) -- an extra full pass over the raster; here the
// warp-strip pipeline moves the raw elements with TMA and widens / narrows them to float32 in
// registers (same rounding as astype: int -> f32 and f64 -> f32, round to nearest even), then
// runs the unchanged float32 operators.  For integer rasters the TMA unit can only zero-fill
// out-of-raster cells, so the loader substitutes NaN from the cell coordinates.
#include <math.h>
#include <type_traits>

#include "surface_ops.cuh"

namespace xrs {

template <typename TS> struct IngestBox {
    static constexpr int pad = sizeof(TS) >= 4 ? 4 : 16 / (int)sizeof(TS);  // 16-byte aligned box start
    static constexpr int w = kStripW + 2 * pad;
};

template <typename TS> __device__ __forceinline__ void ingest_load4(const TS *p, float (&c)[4]) {
    if constexpr (sizeof(TS) == 2) {
        const uint2 q = *reinterpret_cast<const uint2 *>(p);
        TS e[4];
        memcpy(e, &q, 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = (float)e[i];
    } else if constexpr (sizeof(TS) == 4) {
        const int4 q = *reinterpret_cast<const int4 *>(p);
        TS e[4];
        memcpy(e, &q, 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = (float)e[i];
    } else {
        const double2 q0 = *reinterpret_cast<const double2 *>(p);
        const double2 q1 = *reinterpret_cast<const double2 *>(p + 2);
        c[0] = (float)q0.x; c[1] = (float)q0.y; c[2] = (float)q1.x; c[3] = (float)q1.y;
    }
}

template <typename Op, typename TS, int ROWS, int STAGES>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
stencil3_tma_ingest_kernel(const __grid_constant__ CUtensorMap tmap,
                           const __grid_constant__ typename Op::Params prm, const OutPtrs<Op> outs,
                           const StripGeom g) {
    static_assert(std::is_same<typename Op::in_t, float>::value, "ingest feeds float32 operators");
    using TO = typename Op::out_t;
    constexpr bool kIntegral = std::is_integral<TS>::value;
    constexpr int kW = IngestBox<TS>::w, kP = IngestBox<TS>::pad;
    constexpr int kStageElems = ROWS * kW;
    constexpr uint32_t kStageBytes = kStageElems * sizeof(TS);
    static_assert(kStageBytes % 128 == 0, "TMA destination must stay 128-byte aligned");

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TS *ring = reinterpret_cast<TS *>(smem_raw) + (size_t)warp * STAGES * kStageElems;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kWarpsPerCta * STAGES * kStageBytes) +
                     warp * STAGES;
    if (lane == 0) {
        tma_prefetch_desc(&tmap);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    const int64_t n_tasks = (int64_t)g.n_strips * g.n_segs;
    const int64_t total_warps = (int64_t)gridDim.x * kWarpsPerCta;
    uint32_t phase = 0;
    const float qnan = nan_of<float>();
    for (int64_t task = (int64_t)blockIdx.x * kWarpsPerCta + warp; task < n_tasks; task += total_warps) {
        const int seg = (int)(task / g.n_strips), strip = (int)(task % g.n_strips);
        const int64_t x0 = (int64_t)strip * kStripW;
        const int64_t y0 = (int64_t)seg * g.seg_rows;
        const int64_t y1 = min(y0 + (int64_t)g.seg_rows, g.H);
        const int rows_in = (int)(y1 - y0) + 2;
        const int n_chunks = (rows_in + ROWS - 1) / ROWS;
        const int bx = (int)x0 - kP, by = (int)y0 - 1;
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s)
                if (s < n_chunks) {
                    mbar_arrive_expect_tx(&bars[s], kStageBytes);
                    tma_load_2d(ring + s * kStageElems, &tmap, &bars[s], bx, by + s * ROWS);
                }
        }
        Op op(prm);
        const int64_t xl = x0 + kLaneCells * lane;
        const bool lane_ok = xl < g.W;       // W % 4 == 0: a lane is all-in or all-out
        const bool left_oob = xl == 0, right_oob = xl + 4 >= g.W;
        const int seg_h = (int)(y1 - y0);
        const TS *lane_smem = ring + kP + kLaneCells * lane;
        TO *optr[Op::kOutputs];
#pragma unroll
        for (int k = 0; k < Op::kOutputs; ++k) optr[k] = outs.p[k] + (y0 - 3) * outs.pitch_elems + xl;

        int stage = 0;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&bars[stage], (phase >> stage) & 1u);
            phase ^= (1u << stage);
            const TS *buf = lane_smem + stage * kStageElems;
            const int rel = c * ROWS - 2;
            const int64_t ybase = y0 - 1 + (int64_t)c * ROWS;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const TS *p = buf + r * kW;
                Row6<float> row;
                ingest_load4<TS>(p, row.c);
                row.l = (float)p[-1];
                row.r = (float)p[4];
                if constexpr (kIntegral) {  // the TMA unit zero-fills integers: NaN by coordinates
                    const int64_t y = ybase + r;
                    const bool row_oob = (y < 0) || (y >= g.H);
                    if (row_oob || left_oob) row.l = qnan;
                    if (row_oob || right_oob) row.r = qnan;
                    if (row_oob || !lane_ok) row.c[0] = row.c[1] = row.c[2] = row.c[3] = qnan;
                }
                Vec4<TO> o[Op::kOutputs];
                op.step(row, o);
                const bool st = lane_ok && (unsigned)(rel + r) < (unsigned)seg_h;
#pragma unroll
                for (int k = 0; k < Op::kOutputs; ++k) {
                    optr[k] += outs.pitch_elems;
                    if (st) store4v<TO>(optr[k], o[k]);
                }
            }
            __syncwarp();
            if (lane == 0 && c + STAGES < n_chunks) {
                mbar_arrive_expect_tx(&bars[stage], kStageBytes);
                tma_load_2d(ring + stage * kStageElems, &tmap, &bars[stage], bx, by + (c + STAGES) * ROWS);
            }
            stage = (stage + 1 == STAGES) ? 0 : stage + 1;
        }
    }
}

bool make_tensor_map_2d_raw(CUtensorMap *map, const void *base, int64_t pitch_bytes, int64_t H, int64_t W,
                            int dtype, int box_w, int box_h);  // lib_core.cu

template <typename Op, typename TS, int ROWS, int STAGES>
static int launch_ingest(const void *in, int dtype, int64_t in_pitch, const typename Op::Params &prm, float *out,
                         int64_t out_pitch, int64_t H, int64_t W, cudaStream_t stream) {
    CUtensorMap tmap;
    const bool ok = (W % 4 == 0) && (out_pitch % 16 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                    make_tensor_map_2d_raw(&tmap, in, in_pitch, H, W, dtype, IngestBox<TS>::w, ROWS);
    if (!ok) {
        set_error("raster layout not supported by the direct-ingest path (needs 16-byte aligned rows, W %% 4 == 0)");
        return XRS_EUNSUPPORTED;
    }
    OutPtrs<Op> outs;
    outs.p[0] = out;
    outs.pitch_elems = out_pitch / 4;
    const int sms = sm_count();
    StripGeom g;
    g.H = H; g.W = W;
    g.n_strips = (int)((W + kStripW - 1) / kStripW);
    const int64_t resident_warps = (int64_t)sms * 2 * kWarpsPerCta;
    int64_t want_segs = (resident_warps * 8 + g.n_strips - 1) / g.n_strips;
    int64_t seg_rows = (H + want_segs - 1) / (want_segs > 0 ? want_segs : 1);
    if (seg_rows < 64) seg_rows = 64;
    if (seg_rows > H) seg_rows = H;
    seg_rows = ((seg_rows + 2 + ROWS - 1) / ROWS) * ROWS - 2;
    if (seg_rows < 1) seg_rows = 1;
    g.seg_rows = (int)seg_rows;
    g.n_segs = (int)((H + seg_rows - 1) / seg_rows);
    const int64_t n_tasks = (int64_t)g.n_strips * g.n_segs;
    constexpr size_t smem = (size_t)kWarpsPerCta * STAGES * ROWS * IngestBox<TS>::w * sizeof(TS) +
                            (size_t)kWarpsPerCta * STAGES * sizeof(uint64_t);
    auto kern = stencil3_tma_ingest_kernel<Op, TS, ROWS, STAGES>;
    XRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    XRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarpsPerCta * 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    int64_t grid = (int64_t)sms * per_sm;
    const int64_t need = (n_tasks + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid > need) grid = need;
    LaunchInfo &li = last_launch_info();
    li.used_tma = 2;  // 2 = direct-ingest TMA kernel
    li.grid = (int)grid; li.block = kWarpsPerCta * 32; li.smem_bytes = (int)smem;
    kern<<<(unsigned)grid, kWarpsPerCta * 32, smem, stream>>>(tmap, prm, outs, g);
    XRS_CUDA(cudaGetLastError());
    return XRS_OK;
}

template <typename Op>
static int dispatch_dtype(const void *in, int dtype, int64_t in_pitch, const typename Op::Params &prm, float *out,
                          int64_t out_pitch, int64_t H, int64_t W, cudaStream_t st) {
    switch (dtype) {
        case XRS_I16: return launch_ingest<Op, short, 4, 4>(in, dtype, in_pitch, prm, out, out_pitch, H, W, st);
        case XRS_U16: return launch_ingest<Op, unsigned short, 4, 4>(in, dtype, in_pitch, prm, out, out_pitch, H, W, st);
        case XRS_I32: return launch_ingest<Op, int, 4, 4>(in, dtype, in_pitch, prm, out, out_pitch, H, W, st);
        case XRS_F64: return launch_ingest<Op, double, 2, 4>(in, dtype, in_pitch, prm, out, out_pitch, H, W, st);
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
