// Round-2 tuning harness for the warp-strip skeleton (same headers as the product library).
// Sweeps operators x (rows per box, stages, CTAs per SM, tasks per warp) and prints one line each;
// also a plain float4 copy kernel and cudaMemcpy D2D as the box's ceiling on the day.
// Build (scripts/tune/build.sh): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false
//   -I xarray-spatial_b200/csrc tune2.cu xarray-spatial_b200/csrc/lib_core.cu -o tune2
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>
#include "surface_ops.cuh"
using namespace xrs;
#ifdef TUNE_OLD  // round-1 headers (git show HEAD~:...): no square / packed variants
using SlopeSqOp = SlopeOp;
using SuiteSqOp = SuiteOp;
using SlopeParams = SlopeOp::Params;
using SuiteParams = SuiteOp::Params;
#endif

__global__ void fill(float *p, size_t n, int W) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = (float)(i % W), y = (float)(i / W);
        p[i] = 2000.f + 900.f * __sinf(x * 0.0013f) * __cosf(y * 0.0011f) + 35.f * __sinf(x * 0.071f + y * 0.053f);
    }
}
__global__ void copyk(const float4 *a, float4 *b, size_t n4) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

// skeleton-only operator: copies row y-1 (what a 3x3 operator with no arithmetic would cost)
struct CopyOp {
    using in_t = float;
    using out_t = float;
    static constexpr int kOutputs = 1;
    struct Params { int unused; };
    float r1[4];
    __device__ explicit CopyOp(const Params &) { r1[0] = r1[1] = r1[2] = r1[3] = 0.f; }
    __device__ __forceinline__ void step(const Row6<float> &row, Vec4<float> (&out)[1]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { out[0].v[i] = r1[i] + 0.0f * (row.l + row.r); r1[i] = row.c[i]; }
    }
};

static int g_reps = 7;
static cudaEvent_t e0, e1;

template <typename Op, int ROWS, int STAGES>
float run(const typename Op::in_t *in, typename Op::out_t *const *outp, int64_t H, int64_t W,
          const typename Op::Params &prm, int per_sm, int tpw) {
    using T = typename Op::in_t;
    CUtensorMap tmap;
    if (!make_tensor_map_2d(&tmap, in, W * sizeof(T), H, W, sizeof(T), kBoxW, ROWS)) return -3.f;
    OutPtrs<Op> outs;
    for (int k = 0; k < Op::kOutputs; ++k) outs.p[k] = outp[k];
    outs.pitch_elems = W;
    const int sms = sm_count();
    StripGeom g; g.H = H; g.W = W; g.n_strips = (int)((W + kStripW - 1) / kStripW);
    const int64_t resident = (int64_t)sms * per_sm * kWarpsPerCta;
    int64_t want = (resident * tpw + g.n_strips - 1) / g.n_strips;
    int64_t seg_rows = (H + want - 1) / want;
    seg_rows = ((seg_rows + 2 + ROWS - 1) / ROWS) * ROWS - 2;
    g.seg_rows = (int)seg_rows; g.n_segs = (int)((H + seg_rows - 1) / seg_rows);
    constexpr size_t smem = (size_t)kWarpsPerCta * STAGES * ROWS * kBoxW * sizeof(T) + (size_t)kWarpsPerCta * STAGES * 8;
    auto kern = stencil3_tma_kernel<Op, ROWS, STAGES>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
    if (occ < per_sm) return -1.f;
    for (int i = 0; i < 2; ++i) kern<<<sms * per_sm, 256, smem>>>(tmap, prm, outs, g);
    std::vector<float> t;
    for (int i = 0; i < g_reps; ++i) {
        cudaEventRecord(e0);
        kern<<<sms * per_sm, 256, smem>>>(tmap, prm, outs, g);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); t.push_back(ms);
    }
    if (cudaGetLastError() != cudaSuccess) return -2.f;
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

static double g_peak = 6569.6;
static void report(const char *name, int rows, int stages, int per_sm, int tpw, float ms, double bytes) {
    if (ms < 0) { printf("%-22s rows=%d st=%d cta/sm=%d tpw=%2d : n/a (%d)\n", name, rows, stages, per_sm, tpw, (int)ms); return; }
    const double gbs = bytes / (ms * 1e-3) / 1e9;
    printf("%-22s rows=%d st=%d cta/sm=%d tpw=%2d : %7.3f ms %6.0f GB/s  %.3f\n", name, rows, stages, per_sm, tpw, ms, gbs, gbs / g_peak);
    fflush(stdout);
}

int main(int argc, char **argv) {
    const int64_t H = 32768, W = 32768; const size_t n = (size_t)H * W;
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");
    float *in, *o[4];
    cudaMalloc(&in, n * 4);
    for (int k = 0; k < 4; ++k) cudaMalloc(&o[k], n * 4);
    double *od; cudaMalloc(&od, n * 8);
    fill<<<148 * 8, 256>>>(in, n, (int)W); cudaDeviceSynchronize();
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double B8 = 8.0 * n;
    {
        std::vector<float> t;
        for (int i = 0; i < 10; ++i) {
            cudaEventRecord(e0); copyk<<<148 * 16, 256>>>((const float4 *)in, (float4 *)o[0], n / 4); cudaEventRecord(e1);
            cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); t.push_back(ms);
        }
        std::sort(t.begin(), t.end());
        printf("float4 copy kernel : %.3f ms %.0f GB/s %.3f\n", t[5], B8 / (t[5] * 1e-3) / 1e9, B8 / (t[5] * 1e-3) / 1e9 / g_peak);
        t.clear();
        for (int i = 0; i < 10; ++i) {
            cudaEventRecord(e0); cudaMemcpyAsync(o[0], in, n * 4, cudaMemcpyDeviceToDevice); cudaEventRecord(e1);
            cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); t.push_back(ms);
        }
        std::sort(t.begin(), t.end());
        printf("cudaMemcpy D2D     : %.3f ms %.0f GB/s %.3f\n", t[5], B8 / (t[5] * 1e-3) / 1e9, B8 / (t[5] * 1e-3) / 1e9 / g_peak);
    }
    HillshadeOp::Params hp = {0.42f, 0.2f, -0.3f};
    SlopeParams sp = {1.0, 1.7e-5f};
    SlopeParams sp2 = {1.25, 1.7e-5f};
    AspectOp::Params ap = {0};
    CurvatureOp::Params cp = {100.0 / 900.0};
    CopyOp::Params kp = {0};
    using FM = FocalMeanOp<float, float, false>;
    using FMD = FocalMeanOp<float, double, false>;
    FM::Params fp; memset(&fp, 0, sizeof(fp)); fp.ex_nan = 1;
    FMD::Params fdp; memset(&fdp, 0, sizeof(fdp)); fdp.ex_nan = 1;
    SuiteParams up; up.slope = sp; up.curv = cp; up.hill = hp;
    Conv3Op::Params c3; for (int i = 0; i < 9; ++i) c3.w[i] = 0.1 * (i + 1);
    float *o1[1] = {o[0]};
    double *od1[1] = {od};
    float *o4[4] = {o[0], o[1], o[2], o[3]};
    float *o3[4] = {o[0], o[1], o[2], nullptr};

#define RUN1(NAME, OP, PRM, OUT, R, S, P, TPW, BYTES) report(NAME, R, S, P, TPW, run<OP, R, S>(in, OUT, H, W, PRM, P, TPW), BYTES);
#define SWEEP(NAME, OP, PRM, OUT, BYTES)                                  \
    RUN1(NAME, OP, PRM, OUT, 4, 4, 2, 8, BYTES)                           \
    RUN1(NAME, OP, PRM, OUT, 4, 4, 3, 8, BYTES)                           \
    RUN1(NAME, OP, PRM, OUT, 4, 4, 1, 8, BYTES)                           \
    if (!quick) {                                                         \
        RUN1(NAME, OP, PRM, OUT, 8, 3, 2, 8, BYTES)                       \
        RUN1(NAME, OP, PRM, OUT, 8, 2, 3, 8, BYTES)                       \
        RUN1(NAME, OP, PRM, OUT, 8, 4, 1, 8, BYTES)                       \
        RUN1(NAME, OP, PRM, OUT, 4, 4, 2, 4, BYTES)                       \
        RUN1(NAME, OP, PRM, OUT, 4, 4, 2, 16, BYTES)                      \
        RUN1(NAME, OP, PRM, OUT, 4, 6, 2, 8, BYTES)                       \
    }
    SWEEP("copyop", CopyOp, kp, o1, B8)
    SWEEP("hillshade", HillshadeOp, hp, o1, B8)
    SWEEP("slope(square)", SlopeSqOp, sp, o1, B8)
    SWEEP("slope(rxy)", SlopeOp, sp2, o1, B8)
    SWEEP("aspect", AspectOp, ap, o1, B8)
    SWEEP("curvature", CurvatureOp, cp, o1, B8)
    SWEEP("focal.mean f32", FM, fp, o1, B8)
    SWEEP("conv3", Conv3Op, c3, o1, B8)
    SWEEP("suite4", SuiteSqOp, up, o4, 20.0 * n)
    SWEEP("suite3", SuiteSqOp, up, o3, 16.0 * n)
    {
        // f32 -> f64 focal.mean (host-path flavour): 12 B/cell, half the rows (8 GiB output)
        const int64_t H2 = H / 2;
        report("focal.mean f32->f64", 4, 4, 2, 8, run<FMD, 4, 4>(in, od1, H2, W, fdp, 2, 8), 12.0 * H2 * W);
        report("focal.mean f32->f64", 4, 4, 3, 8, run<FMD, 4, 4>(in, od1, H2, W, fdp, 3, 8), 12.0 * H2 * W);
    }
    return 0;
}
