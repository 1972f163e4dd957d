// Standalone tuning harness for the warp-strip skeleton: times HillshadeOp / SlopeOp with different
// (ROWS, STAGES, CTAs per SM).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
//   -fmad=false -I../../xarray-spatial_b200/csrc tune_stencil.cu ../../xarray-spatial_b200/csrc/lib_core.cu -o tune_stencil
#include <cstdio>
#include <vector>
#include "surface_ops.cuh"
using namespace xrs;

__global__ void fill(float *p, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 1000.f + 500.f * __sinf((float)(i % 32768) * 0.001f) + (float)((i / 32768) % 977) * 0.37f;
}
__global__ void copyk(const float4 *a, float4 *b, size_t n4) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <typename Op, int ROWS, int STAGES>
float run(const float *in, float *out, int64_t H, int64_t W, const typename Op::Params &prm, int per_sm, int reps) {
    CUtensorMap tmap;
    make_tensor_map_2d(&tmap, in, W * 4, H, W, 4, kBoxW, ROWS);
    OutPtrs<Op> outs; outs.p[0] = out; outs.pitch_elems = W;
    const int sms = sm_count();
    StripGeom g; g.H = H; g.W = W; g.n_strips = (int)((W + kStripW - 1) / kStripW);
    const int64_t resident = (int64_t)sms * per_sm * kWarpsPerCta;
    int64_t want = (resident * 8 + g.n_strips - 1) / g.n_strips;
    int64_t seg_rows = (H + want - 1) / want;
    seg_rows = ((seg_rows + 2 + ROWS - 1) / ROWS) * ROWS - 2;
    g.seg_rows = (int)seg_rows; g.n_segs = (int)((H + seg_rows - 1) / seg_rows);
    constexpr size_t smem = (size_t)kWarpsPerCta * STAGES * ROWS * kBoxW * 4 + (size_t)kWarpsPerCta * STAGES * 8;
    auto kern = stencil3_tma_kernel<Op, ROWS, STAGES>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
    if (occ < per_sm) return -1.f;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<<<sms * per_sm, 256, smem>>>(tmap, prm, outs, g);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) kern<<<sms * per_sm, 256, smem>>>(tmap, prm, outs, g);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (cudaGetLastError() != cudaSuccess) return -2.f;
    return ms / reps;
}

int main() {
    const int64_t H = 32768, W = 32768; const size_t n = (size_t)H * W;
    float *in, *out; cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 4);
    fill<<<148 * 8, 256>>>(in, n); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) copyk<<<148 * 16, 256>>>((const float4 *)in, (float4 *)out, n / 4);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) copyk<<<148 * 16, 256>>>((const float4 *)in, (float4 *)out, n / 4);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("float4 copy kernel        : %.3f ms  %.0f GB/s\n", ms / 10, 2.0 * n * 4 / (ms / 10 * 1e-3) / 1e9);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) cudaMemcpyAsync(out, in, n * 4, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D            : %.3f ms  %.0f GB/s\n", ms / 10, 2.0 * n * 4 / (ms / 10 * 1e-3) / 1e9);
    HillshadeOp::Params hp = {0.42f, 0.2f, -0.3f};
    SlopeOp::Params sp = {1.0, 1.7e-5f};
#define RUNH(R, S, P) { float t = run<HillshadeOp, R, S>(in, out, H, W, hp, P, 10); \
    printf("hillshade rows=%d stages=%d cta/sm=%d : %.3f ms  %.0f GB/s\n", R, S, P, t, 2.0 * n * 4 / (t * 1e-3) / 1e9); }
#define RUNS(R, S, P) { float t = run<SlopeOp, R, S>(in, out, H, W, sp, P, 10); \
    printf("slope     rows=%d stages=%d cta/sm=%d : %.3f ms  %.0f GB/s\n", R, S, P, t, 2.0 * n * 4 / (t * 1e-3) / 1e9); }
    RUNH(4, 4, 1) RUNH(4, 4, 2) RUNH(8, 3, 1) RUNS(4, 4, 2) RUNS(4, 6, 2)
    return 0;
}
