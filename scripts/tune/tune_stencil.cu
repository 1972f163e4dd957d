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


// ---- experiment: outputs leave through TMA bulk tensor stores instead of STG.128 -------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(x), "r"(y) : "memory");
}
template <typename Op, int STAGES, int WARPS, int OBUF>
__global__ void __launch_bounds__(WARPS * 32)
stencil3_tmastore_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap omap,
                         const __grid_constant__ typename Op::Params prm, const StripGeom g) {
    constexpr int ROWS = 4;
    using T = float;
    constexpr int kStageElems = ROWS * kBoxW;
    constexpr uint32_t kStageBytes = kStageElems * sizeof(T);
    constexpr int kOutElems = ROWS * kStripW;  // 2 KB per staging buffer
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T *ring = reinterpret_cast<T *>(smem_raw) + (size_t)warp * STAGES * kStageElems;
    float *ostage = reinterpret_cast<float *>(smem_raw + (size_t)WARPS * STAGES * kStageBytes) + (size_t)warp * OBUF * kOutElems;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)WARPS * (STAGES * kStageBytes + OBUF * kOutElems * 4)) + warp * STAGES;
    if (lane == 0) {
        tma_prefetch_desc(&tmap); tma_prefetch_desc(&omap);
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const int64_t n_tasks = (int64_t)g.n_strips * g.n_segs;
    const int64_t total_warps = (int64_t)gridDim.x * WARPS;
    uint32_t phase = 0;
    int ob = 0;
    for (int64_t task = (int64_t)blockIdx.x * WARPS + warp; task < n_tasks; task += total_warps) {
        const int seg = (int)(task / g.n_strips), strip = (int)(task % g.n_strips);
        const int64_t x0 = (int64_t)strip * kStripW;
        const int64_t y0 = (int64_t)seg * g.seg_rows;   // seg_rows % 4 == 0
        const int64_t y1 = min(y0 + (int64_t)g.seg_rows, g.H);
        const int rows_in = (int)(y1 - y0) + 4;          // input rows y0-3 .. y1: output groups stay 4-aligned
        const int n_chunks = (rows_in + ROWS - 1) / ROWS;
        const int bx = (int)x0 - kPad, by = (int)y0 - 3;
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s)
                if (s < n_chunks) {
                    mbar_arrive_expect_tx(&bars[s], kStageBytes);
                    tma_load_2d(ring + s * kStageElems, &tmap, &bars[s], bx, by + s * ROWS);
                }
        }
        Op op(prm);
        const T *lane_smem = ring + kPad + kLaneCells * lane;
        int stage = 0;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&bars[stage], (phase >> stage) & 1u);
            phase ^= (1u << stage);
            const T *buf = lane_smem + stage * kStageElems;
            float *ost = ostage + ob * kOutElems + kLaneCells * lane;
            if (c > 0) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(OBUF - 1) : "memory");
                __syncwarp();
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const Row6<T> row = load_row_smem<T>(buf + r * kBoxW);
                Vec4<float> o[1];
                op.step(row, o);
                if (c > 0) *reinterpret_cast<float4 *>(ost + r * kStripW) = make_float4(o[0].v[0], o[0].v[1], o[0].v[2], o[0].v[3]);
            }
            if (c > 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&omap, ostage + ob * kOutElems, (int)x0, (int)(y0 + 4 * (c - 1)));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                ob = (ob + 1 == OBUF) ? 0 : ob + 1;
            } else {
                __syncwarp();
            }
            if (lane == 0 && c + STAGES < n_chunks) {
                mbar_arrive_expect_tx(&bars[stage], kStageBytes);
                tma_load_2d(ring + stage * kStageElems, &tmap, &bars[stage], bx, by + (c + STAGES) * ROWS);
            }
            stage = (stage + 1 == STAGES) ? 0 : stage + 1;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename Op, int STAGES, int WARPS = 8, int OBUF = 2>
float run_ts(const float *in, float *out, int64_t H, int64_t W, const typename Op::Params &prm, int per_sm, int reps, int tpw = 8) {
    CUtensorMap tmap, omap;
    make_tensor_map_2d(&tmap, in, W * 4, H, W, 4, kBoxW, 4);
    make_tensor_map_2d(&omap, out, W * 4, H, W, 4, kStripW, 4);
    const int sms = sm_count();
    StripGeom g; g.H = H; g.W = W; g.n_strips = (int)((W + kStripW - 1) / kStripW);
    const int64_t resident = (int64_t)sms * per_sm * WARPS;
    int64_t want = (resident * tpw + g.n_strips - 1) / g.n_strips;
    int64_t seg_rows = (H + want - 1) / want;
    seg_rows = (seg_rows + 3) / 4 * 4;
    g.seg_rows = (int)seg_rows; g.n_segs = (int)((H + seg_rows - 1) / seg_rows);
    constexpr size_t smem = (size_t)WARPS * (STAGES * 4 * kBoxW * 4 + OBUF * 4 * kStripW * 4) + (size_t)WARPS * STAGES * 8;
    auto kern = stencil3_tmastore_kernel<Op, STAGES, WARPS, OBUF>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem);
    if (occ < per_sm) return -1.f;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<<<sms * per_sm, WARPS * 32, smem>>>(tmap, omap, prm, g);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) kern<<<sms * per_sm, WARPS * 32, smem>>>(tmap, omap, prm, g);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return -2.f; }
    return ms / reps;
}

// checks the TMA-store variant against the plain one (bitwise)
__global__ void diffk(const float *a, const float *b, size_t n, unsigned long long *cnt) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    unsigned long long c = 0;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned x = __float_as_uint(a[i]), y = __float_as_uint(b[i]);
        if (x != y && !(a[i] != a[i] && b[i] != b[i])) c++;
    }
    if (c) atomicAdd(cnt, c);
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
    RUNH(4, 4, 1) RUNH(4, 4, 2) RUNS(4, 4, 2)
    {
        float *out2; cudaMalloc(&out2, n * 4);
        run<HillshadeOp, 4, 4>(in, out, H, W, hp, 2, 1);
#define TS(OPN, OP, PRM, S, WP, OB, PS, TPW) { float t = run_ts<OP, S, WP, OB>(in, out2, H, W, PRM, PS, 10, TPW); \
    printf("%-9s TMA-store stages=%d warps=%2d obuf=%d cta/sm=%d tpw=%2d : %.3f ms  %.0f GB/s\n", OPN, S, WP, OB, PS, TPW, t, 2.0 * n * 4 / (t * 1e-3) / 1e9); }
        TS("hillshade", HillshadeOp, hp, 4, 8, 2, 1, 8)
        TS("hillshade", HillshadeOp, hp, 4, 8, 2, 1, 4)
        TS("hillshade", HillshadeOp, hp, 4, 8, 2, 1, 16)
        TS("hillshade", HillshadeOp, hp, 4, 8, 3, 1, 8)
        TS("hillshade", HillshadeOp, hp, 3, 8, 2, 1, 8)
        TS("hillshade", HillshadeOp, hp, 6, 8, 2, 1, 8)
        TS("hillshade", HillshadeOp, hp, 8, 8, 2, 1, 8)
        TS("hillshade", HillshadeOp, hp, 4, 4, 2, 1, 8)
        TS("hillshade", HillshadeOp, hp, 4, 4, 2, 2, 8)
        TS("hillshade", HillshadeOp, hp, 4, 6, 2, 1, 8)
        TS("hillshade", HillshadeOp, hp, 4, 10, 2, 1, 8)
        TS("hillshade", HillshadeOp, hp, 4, 12, 2, 1, 8)
        TS("hillshade", HillshadeOp, hp, 4, 16, 2, 1, 8)
        TS("hillshade", HillshadeOp, hp, 3, 16, 2, 1, 8)
        TS("hillshade", HillshadeOp, hp, 2, 16, 2, 1, 8)
        TS("slope", SlopeOp, sp, 4, 8, 2, 1, 8)
        TS("slope", SlopeOp, sp, 4, 12, 2, 1, 8)
        TS("slope", SlopeOp, sp, 6, 8, 2, 1, 8)
        TS("hillshade", HillshadeOp, hp, 4, 8, 2, 1, 8)
        unsigned long long *cnt; cudaMalloc(&cnt, 8); cudaMemset(cnt, 0, 8);
        diffk<<<148 * 8, 256>>>(out, out2, n, cnt);
        unsigned long long hc = 0; cudaMemcpy(&hc, cnt, 8, cudaMemcpyDeviceToHost);
        printf("cells differing between STG and TMA-store outputs: %llu\n", hc);
    }
    return 0;
}
