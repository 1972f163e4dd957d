// Round-2 harness, part 3: geometry sweep of the PRODUCT kernel (stencil3_tma_kernel, CTA-wide pipeline)
// for the operators that tune3 did not cover (suite, 3x3 convolution, float64 focal.mean) and a finer
// sweep for slope / aspect.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>
#include "surface_ops.cuh"
using namespace xrs;

__global__ void fill(float *p, size_t n, int W) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = (float)(i % W), y = (float)(i / W);
        p[i] = 2000.f + 900.f * __sinf(x * 0.0013f) * __cosf(y * 0.0011f) + 35.f * __sinf(x * 0.071f + y * 0.053f);
    }
}
__global__ void widen(const float *a, double *b, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = (double)a[i];
}
static int g_reps = 7;
static cudaEvent_t e0, e1;
static double g_peak = 6569.6;
template <typename F> float time_it(F f) {
    for (int i = 0; i < 2; ++i) f();
    std::vector<float> t;
    for (int i = 0; i < g_reps; ++i) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); t.push_back(ms);
    }
    if (cudaGetLastError() != cudaSuccess) return -2.f;
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}
template <typename Op, int ROWS, int STAGES, int WARPS>
float run(const typename Op::in_t *in, typename Op::out_t *const *outp, int64_t H, int64_t W, const typename Op::Params &prm, int per_sm) {
    using T = typename Op::in_t;
    CUtensorMap tmap;
    if (!make_tensor_map_2d(&tmap, in, W * sizeof(T), H, W, sizeof(T), kSubW, ROWS)) return -3.f;
    OutPtrs<Op> outs;
    for (int k = 0; k < Op::kOutputs; ++k) outs.p[k] = outp[k];
    outs.pitch_elems = W;
    const int sms = sm_count();
    constexpr int kTileW = TileShape<WARPS>::kTileW;
    TileGeom g; g.H = H; g.W = W; g.n_tiles = (int)((W + kTileW - 1) / kTileW);
    const int64_t grid = (int64_t)sms * per_sm;
    int64_t want = (grid * 8 + g.n_tiles - 1) / g.n_tiles;
    int64_t seg_rows = (H + want - 1) / want;
    seg_rows = ((seg_rows + 2 + ROWS - 1) / ROWS) * ROWS - 2;
    g.seg_rows = (int)seg_rows; g.n_segs = (int)((H + seg_rows - 1) / seg_rows);
    constexpr size_t smem = (size_t)STAGES * TileShape<WARPS>::kNSub * ROWS * kSubW * sizeof(T) + 2 * STAGES * 8;
    auto kern = stencil3_tma_kernel<Op, ROWS, STAGES, WARPS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (WARPS + 1) * 32, smem);
    if (occ < per_sm) return -1.f;
    return time_it([&] { kern<<<(unsigned)grid, (WARPS + 1) * 32, smem>>>(tmap, prm, outs, g); });
}
static void report(const char *name, const char *cfg, float ms, double bytes) {
    if (ms < 0) { printf("%-18s %-30s : n/a (%d)\n", name, cfg, (int)ms); return; }
    const double gbs = bytes / (ms * 1e-3) / 1e9;
    printf("%-18s %-30s : %7.3f ms %6.0f GB/s  %.3f\n", name, cfg, ms, gbs, gbs / g_peak);
    fflush(stdout);
}
int main() {
    const int64_t H = 32768, W = 32768; const size_t n = (size_t)H * W;
    float *in, *o[4];
    cudaMalloc(&in, n * 4);
    for (int k = 0; k < 4; ++k) cudaMalloc(&o[k], n * 4);
    double *ind, *od;
    cudaMalloc(&ind, n * 4); cudaMalloc(&od, n * 8);   // f64 input: half the rows
    fill<<<148 * 8, 256>>>(in, n, (int)W);
    widen<<<148 * 8, 256>>>(in, ind, n / 2);
    cudaDeviceSynchronize();
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    char cfg[96];
    HillshadeOp::Params hp = {0.42f, 0.2f, -0.3f};
    SlopeParams sp = {1.0, 1.7e-5f};
    SlopeParams sp2 = {1.25, 1.7e-5f};
    AspectOp::Params ap = {0};
    CurvatureOp::Params cp = {100.0 / 900.0};
    using FM = FocalMeanOp<float, float, false>;
    using FMD = FocalMeanOp<float, double, false>;
    using FDD = FocalMeanOp<double, double, false>;
    FM::Params fp; memset(&fp, 0, sizeof(fp)); fp.ex_nan = 1;
    FMD::Params fdp; memset(&fdp, 0, sizeof(fdp)); fdp.ex_nan = 1;
    FDD::Params fddp; memset(&fddp, 0, sizeof(fddp)); fddp.ex_nan = 1;
    SuiteParams up; up.slope = sp; up.curv = cp; up.hill = hp;
    Conv3Op::Params c3; for (int i = 0; i < 9; ++i) c3.w[i] = 0.1 * (i + 1);
    float *o1[1] = {o[0]};
    double *od1[1] = {od};
    float *o4[4] = {o[0], o[1], o[2], o[3]};
    float *o3[4] = {o[0], o[1], o[2], nullptr};
    const double B8 = 8.0 * n;
#define RUN(NAME, OP, IN, OUT, HH, PRM, R, S, WP, P, BYTES) { snprintf(cfg, sizeof cfg, "r%d s%d warps=%d cta/sm=%d", R, S, WP, P); \
        report(NAME, cfg, run<OP, R, S, WP>(IN, OUT, HH, W, PRM, P), BYTES); }
#define SW_LIGHT(NAME, OP, IN, OUT, HH, PRM, BYTES) RUN(NAME, OP, IN, OUT, HH, PRM, 2, 4, 16, 1, BYTES) RUN(NAME, OP, IN, OUT, HH, PRM, 4, 2, 8, 2, BYTES) \
        RUN(NAME, OP, IN, OUT, HH, PRM, 4, 4, 8, 1, BYTES) RUN(NAME, OP, IN, OUT, HH, PRM, 4, 3, 16, 1, BYTES) RUN(NAME, OP, IN, OUT, HH, PRM, 4, 3, 8, 2, BYTES) \
        RUN(NAME, OP, IN, OUT, HH, PRM, 2, 6, 16, 1, BYTES) RUN(NAME, OP, IN, OUT, HH, PRM, 2, 4, 8, 2, BYTES) RUN(NAME, OP, IN, OUT, HH, PRM, 2, 5, 16, 1, BYTES)
    SW_LIGHT("hillshade", HillshadeOp, in, o1, H, hp, B8)
    SW_LIGHT("curvature", CurvatureOp, in, o1, H, cp, B8)
    SW_LIGHT("focal.mean f32", FM, in, o1, H, fp, B8)
    SW_LIGHT("conv3", Conv3Op, in, o1, H, c3, B8)
    SW_LIGHT("focal f32->f64", FMD, in, od1, H / 2, fdp, 12.0 * (n / 2))
#define SW_HEAVY(NAME, OP, PRM) RUN(NAME, OP, in, o1, H, PRM, 4, 3, 8, 2, B8) RUN(NAME, OP, in, o1, H, PRM, 4, 4, 8, 2, B8) RUN(NAME, OP, in, o1, H, PRM, 4, 3, 16, 1, B8) \
        RUN(NAME, OP, in, o1, H, PRM, 4, 4, 16, 1, B8) RUN(NAME, OP, in, o1, H, PRM, 8, 2, 8, 2, B8) RUN(NAME, OP, in, o1, H, PRM, 8, 2, 16, 1, B8) \
        RUN(NAME, OP, in, o1, H, PRM, 4, 5, 16, 1, B8) RUN(NAME, OP, in, o1, H, PRM, 4, 3, 12, 1, B8) RUN(NAME, OP, in, o1, H, PRM, 2, 6, 8, 2, B8) \
        RUN(NAME, OP, in, o1, H, PRM, 2, 8, 16, 1, B8) RUN(NAME, OP, in, o1, H, PRM, 4, 2, 8, 3, B8) RUN(NAME, OP, in, o1, H, PRM, 4, 3, 10, 2, B8)
    SW_HEAVY("slope(square)", SlopeSqOp, sp)
    SW_HEAVY("slope(rxy)", SlopeOp, sp2)
    SW_HEAVY("aspect", AspectOp, ap)
#define SW_SUITE(NAME, OUT, BYTES) RUN(NAME, SuiteSqOp, in, OUT, H, up, 4, 4, 8, 1, BYTES) RUN(NAME, SuiteSqOp, in, OUT, H, up, 4, 3, 8, 1, BYTES) \
        RUN(NAME, SuiteSqOp, in, OUT, H, up, 4, 3, 12, 1, BYTES) RUN(NAME, SuiteSqOp, in, OUT, H, up, 4, 4, 12, 1, BYTES) RUN(NAME, SuiteSqOp, in, OUT, H, up, 2, 4, 12, 1, BYTES) \
        RUN(NAME, SuiteSqOp, in, OUT, H, up, 8, 2, 12, 1, BYTES) RUN(NAME, SuiteSqOp, in, OUT, H, up, 2, 6, 12, 1, BYTES) RUN(NAME, SuiteSqOp, in, OUT, H, up, 4, 2, 12, 1, BYTES)
    SW_SUITE("suite4", o4, 20.0 * n)
    SW_SUITE("suite3", o3, 16.0 * n)
#define SW_F64(R, S, WP, P) RUN("focal.mean f64", FDD, ind, od1, H / 2, fddp, R, S, WP, P, 16.0 * (n / 2))
    SW_F64(2, 2, 8, 2) SW_F64(2, 4, 8, 1) SW_F64(2, 2, 16, 1) SW_F64(2, 3, 8, 2) SW_F64(1, 4, 16, 1) SW_F64(1, 4, 8, 2) SW_F64(1, 8, 8, 1) SW_F64(2, 3, 16, 1) SW_F64(4, 2, 8, 1)
    return 0;
}
