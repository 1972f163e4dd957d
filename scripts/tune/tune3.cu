// Round-2 harness, part 2: what limits the warp-strip skeleton at ~0.88 of the copy peak?
//  (1) PaceOp<K>: the copy operator plus K dependent FMAs per row -- does pacing the warps help?
//  (2) ring geometry at 1 / 2 CTAs per SM (bytes in flight per SM)
//  (3) relative placement of the input and output buffers
//  (4) a CTA-wide producer / consumer pipeline: one producer warp issues TMA boxes for the whole
//      CTA tile (WARPS x 128 cells + halo, in boxes of 208 cells), WARPS consumer warps, full / empty
//      mbarriers per stage.
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
__global__ void copyk(const float4 *a, float4 *b, size_t n4) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void diffk(const float *a, const float *b, size_t n, unsigned long long *cnt) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    unsigned long long c = 0;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned x = __float_as_uint(a[i]), y = __float_as_uint(b[i]);
        if (x != y && !(a[i] != a[i] && b[i] != b[i])) c++;
    }
    if (c) atomicAdd(cnt, c);
}

template <int K> struct PaceOp {
    using in_t = float;
    using out_t = float;
    static constexpr int kOutputs = 1;
    struct Params { float a; };
    const Params &p;
    float r1[4];
    __device__ explicit PaceOp(const Params &pp) : p(pp) { r1[0] = r1[1] = r1[2] = r1[3] = 0.f; }
    __device__ __forceinline__ void step(const Row6<float> &row, Vec4<float> (&out)[1]) {
        float t = row.l + row.r;
#pragma unroll
        for (int k = 0; k < K; ++k) t = fmaf(t, p.a, 1.0f);
#pragma unroll
        for (int i = 0; i < 4; ++i) { out[0].v[i] = r1[i] + 0.0f * t; r1[i] = row.c[i]; }
    }
};

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

template <typename Op, int ROWS, int STAGES>
float run(const float *in, float *const *outp, int64_t H, int64_t W, const typename Op::Params &prm, int per_sm) {
    CUtensorMap tmap;
    if (!make_tensor_map_2d(&tmap, in, W * 4, H, W, 4, kBoxW, ROWS)) return -3.f;
    OutPtrs<Op> outs;
    for (int k = 0; k < Op::kOutputs; ++k) outs.p[k] = outp[k];
    outs.pitch_elems = W;
    const int sms = sm_count();
    StripGeom g; g.H = H; g.W = W; g.n_strips = (int)((W + kStripW - 1) / kStripW);
    const int64_t resident = (int64_t)sms * per_sm * kWarpsPerCta;
    int64_t want = (resident * 8 + g.n_strips - 1) / g.n_strips;
    int64_t seg_rows = (H + want - 1) / want;
    seg_rows = ((seg_rows + 2 + ROWS - 1) / ROWS) * ROWS - 2;
    g.seg_rows = (int)seg_rows; g.n_segs = (int)((H + seg_rows - 1) / seg_rows);
    constexpr size_t smem = (size_t)kWarpsPerCta * STAGES * ROWS * kBoxW * 4 + (size_t)kWarpsPerCta * STAGES * 8;
    auto kern = stencil3_tma_kernel<Op, ROWS, STAGES>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
    if (occ < per_sm) return -1.f;
    return time_it([&] { kern<<<sms * per_sm, 256, smem>>>(tmap, prm, outs, g); });
}

// ------------------------------------------------------------------ CTA-wide pipeline
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
constexpr int kSubW = 208;  // cells per TMA box (832 B rows)
struct CtaGeom { int64_t H, W; int n_tiles, n_segs, seg_rows; };

template <typename Op, int ROWS, int STAGES, int WARPS>
__global__ void __launch_bounds__((WARPS + 1) * 32)
stencil3_cta_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ typename Op::Params prm,
                    const OutPtrs<Op> outs, const CtaGeom g) {
    using T = float;
    using TO = typename Op::out_t;
    constexpr int kTileW = 128 * WARPS;
    constexpr int kNSub = (kTileW + 8 + kSubW - 1) / kSubW;
    constexpr int kStageElems = kNSub * ROWS * kSubW;
    constexpr uint32_t kStageBytes = kStageElems * sizeof(T);
    static_assert((ROWS * kSubW * sizeof(T)) % 128 == 0, "box destination alignment");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T *ring = reinterpret_cast<T *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * kStageBytes);
    uint64_t *empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t n_tasks = (int64_t)g.n_tiles * g.n_segs;
    if (warp == WARPS) {
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            for (int64_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
                const int seg = (int)(task / g.n_tiles), tile = (int)(task % g.n_tiles);
                const int64_t y0 = (int64_t)seg * g.seg_rows, y1 = min(y0 + (int64_t)g.seg_rows, g.H);
                const int n_chunks = ((int)(y1 - y0) + 2 + ROWS - 1) / ROWS;
                const int bx = tile * kTileW - 4, by = (int)y0 - 1;
                for (int c = 0; c < n_chunks; ++c) {
                    mbar_wait(&empty[stage], ((ph >> stage) & 1u) ^ 1u);
                    ph ^= (1u << stage);
                    mbar_arrive_expect_tx(&full[stage], kStageBytes);
#pragma unroll
                    for (int b = 0; b < kNSub; ++b)
                        tma_load_2d(ring + stage * kStageElems + b * ROWS * kSubW, &tmap, &full[stage], bx + b * kSubW, by + c * ROWS);
                    stage = (stage + 1 == STAGES) ? 0 : stage + 1;
                }
            }
        }
        return;
    }
    const int cc = 4 + 128 * warp + 4 * lane;
    const int off_c = (cc / kSubW) * ROWS * kSubW + cc % kSubW;
    const int off_l = ((cc - 1) / kSubW) * ROWS * kSubW + (cc - 1) % kSubW;
    const int off_r = ((cc + 4) / kSubW) * ROWS * kSubW + (cc + 4) % kSubW;
    int stage = 0; uint32_t ph = 0;
    for (int64_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const int seg = (int)(task / g.n_tiles), tile = (int)(task % g.n_tiles);
        const int64_t y0 = (int64_t)seg * g.seg_rows, y1 = min(y0 + (int64_t)g.seg_rows, g.H);
        const int seg_h = (int)(y1 - y0);
        const int n_chunks = (seg_h + 2 + ROWS - 1) / ROWS;
        Op op(prm);
        const int64_t xl = (int64_t)tile * kTileW + 128 * warp + 4 * lane;
        const bool lane_ok = xl < g.W;
        TO *optr[Op::kOutputs];
#pragma unroll
        for (int k = 0; k < Op::kOutputs; ++k) optr[k] = outs.p[k] + (y0 - 3) * outs.pitch_elems + xl;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&full[stage], (ph >> stage) & 1u);
            ph ^= (1u << stage);
            const T *buf = ring + stage * kStageElems;
            const int rel = c * ROWS - 2;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                Row6<T> row;
                const float4 q = *reinterpret_cast<const float4 *>(buf + off_c + r * kSubW);
                row.c[0] = q.x; row.c[1] = q.y; row.c[2] = q.z; row.c[3] = q.w;
                row.l = buf[off_l + r * kSubW];
                row.r = buf[off_r + r * kSubW];
                Vec4<TO> o[Op::kOutputs];
                op.step(row, o);
                const bool st = lane_ok && (unsigned)(rel + r) < (unsigned)seg_h;
#pragma unroll
                for (int k = 0; k < Op::kOutputs; ++k) {
                    optr[k] += outs.pitch_elems;
                    if (st && (Op::kOutputs == 1 || outs.p[k] != nullptr)) store4v<TO>(optr[k], o[k]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            stage = (stage + 1 == STAGES) ? 0 : stage + 1;
        }
    }
}

template <typename Op, int ROWS, int STAGES, int WARPS>
float run_cta(const float *in, float *const *outp, int64_t H, int64_t W, const typename Op::Params &prm, int per_sm, int tasks_per_cta = 8) {
    CUtensorMap tmap;
    if (!make_tensor_map_2d(&tmap, in, W * 4, H, W, 4, kSubW, ROWS)) return -3.f;
    OutPtrs<Op> outs;
    for (int k = 0; k < Op::kOutputs; ++k) outs.p[k] = outp[k];
    outs.pitch_elems = W;
    const int sms = sm_count();
    constexpr int kTileW = 128 * WARPS;
    constexpr int kNSub = (kTileW + 8 + kSubW - 1) / kSubW;
    CtaGeom g; g.H = H; g.W = W; g.n_tiles = (int)((W + kTileW - 1) / kTileW);
    const int64_t grid = (int64_t)sms * per_sm;
    int64_t want = (grid * tasks_per_cta + g.n_tiles - 1) / g.n_tiles;
    int64_t seg_rows = (H + want - 1) / want;
    seg_rows = ((seg_rows + 2 + ROWS - 1) / ROWS) * ROWS - 2;
    g.seg_rows = (int)seg_rows; g.n_segs = (int)((H + seg_rows - 1) / seg_rows);
    constexpr size_t smem = (size_t)STAGES * kNSub * ROWS * kSubW * 4 + 2 * STAGES * 8;
    auto kern = stencil3_cta_kernel<Op, ROWS, STAGES, WARPS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (WARPS + 1) * 32, smem);
    if (occ < per_sm) return -1.f;
    return time_it([&] { kern<<<(unsigned)grid, (WARPS + 1) * 32, smem>>>(tmap, prm, outs, g); });
}

static void report(const char *name, const char *cfg, float ms, double bytes) {
    if (ms < 0) { printf("%-16s %-34s : n/a (%d)\n", name, cfg, (int)ms); return; }
    const double gbs = bytes / (ms * 1e-3) / 1e9;
    printf("%-16s %-34s : %7.3f ms %6.0f GB/s  %.3f\n", name, cfg, ms, gbs, gbs / g_peak);
    fflush(stdout);
}

int main(int argc, char **argv) {
    const int64_t H = 32768, W = 32768; const size_t n = (size_t)H * W;
    float *in, *obig, *oref;
    cudaMalloc(&in, n * 4);
    cudaMalloc(&obig, n * 4 + (64 << 20));
    cudaMalloc(&oref, n * 4);
    float *o = obig;
    fill<<<148 * 8, 256>>>(in, n, (int)W); cudaDeviceSynchronize();
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double B8 = 8.0 * n;
    char cfg[128];
    report("copy kernel", "float4 grid-stride 148x16x256", time_it([&] { copyk<<<148 * 16, 256>>>((const float4 *)in, (float4 *)o, n / 4); }), B8);
    report("copy kernel", "float4 grid-stride 148x8x256", time_it([&] { copyk<<<148 * 8, 256>>>((const float4 *)in, (float4 *)o, n / 4); }), B8);
    report("cudaMemcpy", "D2D", time_it([&] { cudaMemcpyAsync(o, in, n * 4, cudaMemcpyDeviceToDevice); }), B8);

    HillshadeOp::Params hp = {0.42f, 0.2f, -0.3f};
    SlopeParams sp = {1.0, 1.7e-5f};
    AspectOp::Params ap = {0};
    CurvatureOp::Params cp = {100.0 / 900.0};
    using FM = FocalMeanOp<float, float, false>;
    FM::Params fp; memset(&fp, 0, sizeof(fp)); fp.ex_nan = 1;
    float *o1[1] = {o};

    // (1) pacing
#define PACE(K, P) { PaceOp<K>::Params pp = {0.5f}; snprintf(cfg, sizeof cfg, "K=%d r4 s4 cta/sm=%d", K, P); \
        report("pace", cfg, run<PaceOp<K>, 4, 4>(in, o1, H, W, pp, P), B8); }
    PACE(0, 1) PACE(16, 1) PACE(32, 1) PACE(64, 1) PACE(96, 1) PACE(128, 1) PACE(192, 1) PACE(256, 1)
    PACE(0, 2) PACE(32, 2) PACE(64, 2) PACE(128, 2) PACE(256, 2)

    // (2) ring geometry
#define GEO(NAME, OP, PRM, R, S, P) { snprintf(cfg, sizeof cfg, "warp-ring r%d s%d cta/sm=%d (%d KB/SM)", R, S, P, P * 8 * R * S * 544 / 1024); \
        report(NAME, cfg, run<OP, R, S>(in, o1, H, W, PRM, P), B8); }
#define GEOS(NAME, OP, PRM) GEO(NAME, OP, PRM, 4, 2, 1) GEO(NAME, OP, PRM, 4, 3, 1) GEO(NAME, OP, PRM, 4, 4, 1) GEO(NAME, OP, PRM, 8, 2, 1) \
        GEO(NAME, OP, PRM, 4, 2, 2) GEO(NAME, OP, PRM, 4, 3, 2) GEO(NAME, OP, PRM, 4, 4, 2) GEO(NAME, OP, PRM, 8, 2, 2)
    { PaceOp<0>::Params pp = {0.5f}; GEOS("copyop", PaceOp<0>, pp) }
    GEOS("hillshade", HillshadeOp, hp)
    GEOS("slope", SlopeSqOp, sp)
    GEOS("focal.mean", FM, fp)

    // (3) buffer placement
    for (size_t off : {(size_t)0, (size_t)2048, (size_t)(16 << 10), (size_t)(256 << 10) + 4096, (size_t)(1 << 20) + 8192, (size_t)(32 << 20) + 65536 + 2048}) {
        float *oo[1] = {obig + off / 4};
        snprintf(cfg, sizeof cfg, "out shifted by %zu B, r4 s4 cta/sm=1", off);
        report("hillshade", cfg, run<HillshadeOp, 4, 4>(in, oo, H, W, hp, 1), B8);
    }

    // (4) CTA-wide pipeline; first a bit-for-bit check against the warp-ring kernel
    {
        float *orr[1] = {oref};
        run<SlopeSqOp, 4, 4>(in, orr, H, W, sp, 2);
        run_cta<SlopeSqOp, 4, 3, 8>(in, o1, H, W, sp, 1);
        unsigned long long *cnt; cudaMalloc(&cnt, 8); cudaMemset(cnt, 0, 8);
        diffk<<<148 * 8, 256>>>(o, oref, n, cnt);
        unsigned long long hc = 0; cudaMemcpy(&hc, cnt, 8, cudaMemcpyDeviceToHost);
        printf("cta-wide vs warp-ring slope outputs, cells differing: %llu (%s)\n", hc, cudaGetErrorString(cudaGetLastError()));
    }
#define CTA(NAME, OP, PRM, R, S, WP, P) { snprintf(cfg, sizeof cfg, "cta-wide r%d s%d warps=%d cta/sm=%d (%d KB/SM)", R, S, WP, P, \
        P * S * R * ((128 * WP + 8 + 207) / 208) * 832 / 1024); report(NAME, cfg, run_cta<OP, R, S, WP>(in, o1, H, W, PRM, P), B8); }
#define CTAS(NAME, OP, PRM) CTA(NAME, OP, PRM, 4, 2, 8, 1) CTA(NAME, OP, PRM, 4, 3, 8, 1) CTA(NAME, OP, PRM, 4, 4, 8, 1) CTA(NAME, OP, PRM, 2, 4, 8, 1) \
        CTA(NAME, OP, PRM, 2, 6, 8, 1) CTA(NAME, OP, PRM, 8, 2, 8, 1) CTA(NAME, OP, PRM, 4, 3, 8, 2) CTA(NAME, OP, PRM, 4, 2, 8, 2) CTA(NAME, OP, PRM, 2, 3, 8, 2) \
        CTA(NAME, OP, PRM, 4, 2, 16, 1) CTA(NAME, OP, PRM, 4, 3, 16, 1) CTA(NAME, OP, PRM, 2, 3, 16, 1) CTA(NAME, OP, PRM, 2, 4, 16, 1) CTA(NAME, OP, PRM, 4, 3, 12, 1)
    { PaceOp<0>::Params pp = {0.5f}; CTAS("copyop", PaceOp<0>, pp) }
    CTAS("hillshade", HillshadeOp, hp)
    CTAS("curvature", CurvatureOp, cp)
    CTAS("slope", SlopeSqOp, sp)
    CTAS("aspect", AspectOp, ap)
    CTAS("focal.mean", FM, fp)
    return 0;
}
