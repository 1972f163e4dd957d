// stencil3.cuh -- the 3x3 "warp strip" skeleton shared by slope / aspect / curvature /
// hillshade / focal.mean / the fused surface suite.
//
// Decomposition (B200-first, not a port of the reference's one-thread-per-cell kernels,
// slope.py:133-142): the raster is cut into column strips of 128 cells and row segments;
// one WARP owns one (segment, strip) task at a time and marches down its rows.  Each lane
// owns 4 adjacent cells (one float4), so a warp reads / writes 512 contiguous bytes per row.
//   * input rows arrive through a per-warp ring of TMA 2-D boxes (136 x ROWS cells:
//     the strip plus a 4-cell pad on both sides, which keeps every lane's float4 16-byte
//     aligned in shared memory); out-of-raster cells are filled with NaN by the TMA unit,
//     which is exactly the reference's raster-edge rule (NaN ring for the Horn family,
//     NaN-skipping clamped windows for focal.mean);
//   * lane 0 of the warp is the producer (arms the stage's mbarrier with expect_tx and
//     issues cp.async.bulk.tensor.2d), all 32 lanes are consumers: no __syncthreads at all;
//   * the three input rows a cell needs are never re-read: operators keep per-row partial
//     results (column differences / weighted row sums) of the two previous rows in
//     registers and combine them with the new row ("push, then emit the row above");
//   * left / right neighbours come from warp shuffles (strip-edge lanes read the pad);
//   * outputs are written as float4 with streaming stores.
// Rasters TMA cannot describe (width not a multiple of 4 cells, unaligned base / pitch) run the
// same pipeline with cp.async fills (stencil3_cpasync_kernel); a plain bounds-checked
// direct-load kernel is kept as the reference implementation of the loader.  All three drive
// the very same operator code.
#pragma once
#include "common.cuh"

namespace xrs {

constexpr int kLaneCells = 4;                        // cells per lane
constexpr int kStripW = 32 * kLaneCells;             // 128 cells per warp-row
constexpr int kPad = 4;                              // pad cells each side (16 B for f32)
constexpr int kBoxW = kStripW + 2 * kPad;            // 136
constexpr int kWarpsPerCta = 8;

template <typename T> struct Vec4;  // 4 consecutive cells
template <> struct Vec4<float> { float v[4]; };
template <> struct Vec4<double> { double v[4]; };

template <typename T> __device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <typename T> __device__ __forceinline__ T shfl_dn1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// One input row as seen by a lane: left neighbour, own 4 cells, right neighbour.
template <typename T> struct Row6 {
    T l, c[4], r;
};

// `p` points at the lane's first cell inside a shared-memory row (16-byte aligned).  The left /
// right neighbours are plain scalar loads (4-way bank conflicted, 8 extra wavefronts per
// warp-row against a budget of ~44 cycles): cheaper in issue slots than shuffles plus
// strip-edge fix-ups, and issue slots are what these kernels run out of.
template <typename T> __device__ __forceinline__ Row6<T> load_row_smem(const T *p) {
    Row6<T> o;
    if constexpr (sizeof(T) == 4) {
        const float4 q = *reinterpret_cast<const float4 *>(p);
        o.c[0] = q.x; o.c[1] = q.y; o.c[2] = q.z; o.c[3] = q.w;
    } else {
        const double2 q0 = *reinterpret_cast<const double2 *>(p);
        const double2 q1 = *reinterpret_cast<const double2 *>(p + 2);
        o.c[0] = q0.x; o.c[1] = q0.y; o.c[2] = q1.x; o.c[3] = q1.y;
    }
    o.l = p[-1];
    o.r = p[4];
    return o;
}

// Direct global loads with bounds checks (out-of-raster cells read as NaN): six independent scalar
// loads per lane and row, so that several rows can be in flight at once.
template <typename T>
__device__ __forceinline__ Row6<T> load_row_direct(const T *in, int64_t pitch_elems, int64_t H,
                                                   int64_t W, int64_t y, int64_t x) {
    Row6<T> o;
    const bool yin = (y >= 0) && (y < H);
    const T *rp = in + (yin ? y : 0) * pitch_elems;
    o.l = (yin && x >= 1 && x - 1 < W) ? __ldg(rp + x - 1) : nan_of<T>();
#pragma unroll
    for (int i = 0; i < 4; ++i) o.c[i] = (yin && (x + i) < W) ? __ldg(rp + x + i) : nan_of<T>();
    o.r = (yin && (x + 4) < W) ? __ldg(rp + x + 4) : nan_of<T>();
    return o;
}

template <typename TO> __device__ __forceinline__ void store4(TO *p, const Vec4<TO> &v, bool vec_ok,
                                                            int nvalid) {
    if (vec_ok && nvalid == 4) {
        if constexpr (sizeof(TO) == 4) {
            __stcs(reinterpret_cast<float4 *>(p), make_float4(v.v[0], v.v[1], v.v[2], v.v[3]));
        } else {
            __stcs(reinterpret_cast<double2 *>(p), make_double2(v.v[0], v.v[1]));
            __stcs(reinterpret_cast<double2 *>(p + 2), make_double2(v.v[2], v.v[3]));
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < nvalid) p[i] = v.v[i];
    }
}

template <typename TO> __device__ __forceinline__ void store4v(TO *p, const Vec4<TO> &v) {
    if constexpr (sizeof(TO) == 4) {
#ifdef XRS_STORE_PLAIN
        *reinterpret_cast<float4 *>(p) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
#elif defined(XRS_STORE_CG)
        __stcg(reinterpret_cast<float4 *>(p), make_float4(v.v[0], v.v[1], v.v[2], v.v[3]));
#else
        __stcs(reinterpret_cast<float4 *>(p), make_float4(v.v[0], v.v[1], v.v[2], v.v[3]));
#endif
    } else {
        __stcs(reinterpret_cast<double2 *>(p), make_double2(v.v[0], v.v[1]));
        __stcs(reinterpret_cast<double2 *>(p + 2), make_double2(v.v[2], v.v[3]));
    }
}

// Operator concept (see surface_ops.cuh):
//   using in_t = float|double;  static constexpr int kOutputs;  using out_t;
//   struct Params;  __device__ Op(const Params&);
//   __device__ void step(const Row6<in_t>& row, Vec4<out_t> (&out)[kOutputs]);
//       -- consumes input row y and produces the outputs of row y-1 (valid once three
//          rows have been pushed).

struct StripGeom {
    int64_t H, W;
    int n_strips, n_segs, seg_rows;
};

template <typename Op> struct OutPtrs {
    typename Op::out_t *p[Op::kOutputs];
    int64_t pitch_elems;  // same for every output
};

// ----------------------------------------------------------------------------- TMA kernel
template <typename Op, int ROWS, int STAGES>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
stencil3_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                    const __grid_constant__ typename Op::Params prm,
                    const OutPtrs<Op> outs, const StripGeom g) {
    using T = typename Op::in_t;
    using TO = typename Op::out_t;
    constexpr int kStageElems = ROWS * kBoxW;
    constexpr uint32_t kStageBytes = kStageElems * sizeof(T);
    static_assert(kStageBytes % 128 == 0, "TMA destination must stay 128-byte aligned");

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T *ring = reinterpret_cast<T *>(smem_raw) + (size_t)warp * STAGES * kStageElems;
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
    uint32_t phase = 0;  // bit s = parity to wait for on stage s
#ifdef XRS_TMA_EVICT_FIRST
    const uint64_t l2pol = l2_policy_evict_first();
#define XRS_TMA_LOAD(dst, bar, x, y) tma_load_2d_hint(dst, &tmap, bar, x, y, l2pol)
#else
#define XRS_TMA_LOAD(dst, bar, x, y) tma_load_2d(dst, &tmap, bar, x, y)
#endif

    for (int64_t task = (int64_t)blockIdx.x * kWarpsPerCta + warp; task < n_tasks; task += total_warps) {
        const int seg = (int)(task / g.n_strips), strip = (int)(task % g.n_strips);
        const int64_t x0 = (int64_t)strip * kStripW;
        const int64_t y0 = (int64_t)seg * g.seg_rows;
        const int64_t y1 = min(y0 + (int64_t)g.seg_rows, g.H);
        const int rows_in = (int)(y1 - y0) + 2;  // input rows y0-1 .. y1
        const int n_chunks = (rows_in + ROWS - 1) / ROWS;
        const int bx = (int)x0 - kPad, by = (int)y0 - 1;

        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s)
                if (s < n_chunks) {
                    mbar_arrive_expect_tx(&bars[s], kStageBytes);
                    XRS_TMA_LOAD(ring + s * kStageElems, &bars[s], bx, by + s * ROWS);
                }
        }

        Op op(prm);
        const int64_t xl = x0 + kLaneCells * lane;
        const bool lane_ok = xl < g.W;  // W % 4 == 0 on this path: a lane is all-in or all-out
        const int seg_h = (int)(y1 - y0);
        const T *lane_smem = ring + kPad + kLaneCells * lane;
        // output pointers one row above the first emitted row (y0 - 2): advanced before every store
        TO *optr[Op::kOutputs];
#pragma unroll
        for (int k = 0; k < Op::kOutputs; ++k) optr[k] = outs.p[k] + (y0 - 3) * outs.pitch_elems + xl;

        int stage = 0;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&bars[stage], (phase >> stage) & 1u);
            phase ^= (1u << stage);
            const T *buf = lane_smem + stage * kStageElems;
            const int rel = c * ROWS - 2;  // output row (relative to y0) of the box's first row
            // Straight-line over the ROWS rows of the box (no branch around the operator state
            // update, so the rolling registers are renamed, not moved).  Rows whose output row
            // falls outside [y0, y1) -- the two lead-in rows and the tail of the last chunk --
            // still update the state; only their store is predicated off.
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const Row6<T> row = load_row_smem<T>(buf + r * kBoxW);
                Vec4<TO> o[Op::kOutputs];
                op.step(row, o);
                const bool st = lane_ok && (unsigned)(rel + r) < (unsigned)seg_h;
#pragma unroll
                for (int k = 0; k < Op::kOutputs; ++k) {
                    optr[k] += outs.pitch_elems;
                    if (st && (Op::kOutputs == 1 || outs.p[k] != nullptr)) store4v<TO>(optr[k], o[k]);
                }
            }
            __syncwarp();  // every lane is done reading this stage
            if (lane == 0 && c + STAGES < n_chunks) {
                mbar_arrive_expect_tx(&bars[stage], kStageBytes);
                XRS_TMA_LOAD(ring + stage * kStageElems, &bars[stage], bx, by + (c + STAGES) * ROWS);
            }
            stage = (stage + 1 == STAGES) ? 0 : stage + 1;
        }
    }
}

// ----------------------------------------------------------------------------- direct kernel
template <typename Op>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
stencil3_direct_kernel(const typename Op::in_t *__restrict__ in, int64_t in_pitch_elems,
                       const __grid_constant__ typename Op::Params prm, const OutPtrs<Op> outs,
                       const StripGeom g,
                       int vec_ok) {
    using T = typename Op::in_t;
    using TO = typename Op::out_t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tasks = (int64_t)g.n_strips * g.n_segs;
    const int64_t total_warps = (int64_t)gridDim.x * kWarpsPerCta;
    for (int64_t task = (int64_t)blockIdx.x * kWarpsPerCta + warp; task < n_tasks; task += total_warps) {
        const int seg = (int)(task / g.n_strips), strip = (int)(task % g.n_strips);
        const int64_t x0 = (int64_t)strip * kStripW;
        const int64_t y0 = (int64_t)seg * g.seg_rows;
        const int64_t y1 = min(y0 + (int64_t)g.seg_rows, g.H);
        Op op(prm);
        const int64_t xl = x0 + kLaneCells * lane;
        const int nvalid = (int)max((int64_t)0, min((int64_t)4, g.W - xl));
        // batches of kBatch rows: all loads of a batch are issued before the first row is consumed
        // (software pipelining; the operator state update itself stays branch-free)
        constexpr int kBatch = 4;
        for (int64_t yb = y0 - 1; yb <= y1; yb += kBatch) {
            Row6<T> rows[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) rows[u] = load_row_direct<T>(in, in_pitch_elems, g.H, g.W, yb + u, xl);
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                Vec4<TO> o[Op::kOutputs];
                op.step(rows[u], o);
                const int64_t yout = yb + u - 1;
                if (yout >= y0 && yout < y1 && nvalid > 0) {
#pragma unroll
                    for (int k = 0; k < Op::kOutputs; ++k)
                        if (outs.p[k] != nullptr)
                            store4<TO>(outs.p[k] + yout * outs.pitch_elems + xl, o[k], vec_ok != 0, nvalid);
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------- cp.async kernel
// Same warp-strip pipeline for rasters TMA cannot describe (width not a multiple of 4 cells, base
// or pitch not 16-byte aligned): the per-warp ring is filled with 4-/8-byte cp.async copies
// (coalesced: lane i copies cells i, i+32, ... of the box row; out-of-raster cells are written
// as NaN with plain stores), completion is tracked with cp.async groups instead of mbarriers.
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <typename T> __device__ __forceinline__ void cp_async_elem(T *dst, const T *src) {
    if constexpr (sizeof(T) == 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

template <typename Op, int ROWS, int STAGES>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
stencil3_cpasync_kernel(const typename Op::in_t *__restrict__ in, int64_t in_pitch_elems,
                        const __grid_constant__ typename Op::Params prm, const OutPtrs<Op> outs,
                        const StripGeom g, int vec_ok) {
    using T = typename Op::in_t;
    using TO = typename Op::out_t;
    constexpr int kStageElems = ROWS * kBoxW;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T *ring = reinterpret_cast<T *>(smem_raw) + (size_t)warp * STAGES * kStageElems;

    const int64_t n_tasks = (int64_t)g.n_strips * g.n_segs;
    const int64_t total_warps = (int64_t)gridDim.x * kWarpsPerCta;
    for (int64_t task = (int64_t)blockIdx.x * kWarpsPerCta + warp; task < n_tasks; task += total_warps) {
        const int seg = (int)(task / g.n_strips), strip = (int)(task % g.n_strips);
        const int64_t x0 = (int64_t)strip * kStripW;
        const int64_t y0 = (int64_t)seg * g.seg_rows;
        const int64_t y1 = min(y0 + (int64_t)g.seg_rows, g.H);
        const int rows_in = (int)(y1 - y0) + 2;
        const int n_chunks = (rows_in + ROWS - 1) / ROWS;
        const int64_t bx = x0 - kPad, by = y0 - 1;

        auto fill = [&](int c) {  // chunk c -> stage c % STAGES
            if (c < n_chunks) {
                T *dst = ring + (c % STAGES) * kStageElems;
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const int64_t y = by + (int64_t)c * ROWS + r;
                    const bool yin = (y >= 0) && (y < g.H);
                    const T *rp = in + (yin ? y : 0) * in_pitch_elems;
#pragma unroll
                    for (int j0 = 0; j0 < kBoxW; j0 += 32) {
                        const int j = j0 + lane;
                        const int64_t x = bx + j;
                        if (j < kBoxW) {
                            if (yin && x >= 0 && x < g.W) cp_async_elem<T>(dst + r * kBoxW + j, rp + x);
                            else dst[r * kBoxW + j] = nan_of<T>();
                        }
                    }
                }
            }
            cp_async_commit();  // one group per chunk slot, possibly empty
        };

#pragma unroll
        for (int c = 0; c < STAGES - 1; ++c) fill(c);

        Op op(prm);
        const int64_t xl = x0 + kLaneCells * lane;
        const int nvalid = (int)max((int64_t)0, min((int64_t)4, g.W - xl));
        const int seg_h = (int)(y1 - y0);
        const T *lane_smem = ring + kPad + kLaneCells * lane;
        TO *optr[Op::kOutputs];
#pragma unroll
        for (int k = 0; k < Op::kOutputs; ++k) optr[k] = outs.p[k] + (y0 - 3) * outs.pitch_elems + xl;

        for (int c = 0; c < n_chunks; ++c) {
            fill(c + STAGES - 1);
            cp_async_wait<STAGES - 1>();  // this lane's copies of chunk c have landed ...
            __syncwarp();                 // ... and so have the other lanes'
            const T *buf = lane_smem + (c % STAGES) * kStageElems;
            const int rel = c * ROWS - 2;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const Row6<T> row = load_row_smem<T>(buf + r * kBoxW);
                Vec4<TO> o[Op::kOutputs];
                op.step(row, o);
                const bool st = nvalid > 0 && (unsigned)(rel + r) < (unsigned)seg_h;
#pragma unroll
                for (int k = 0; k < Op::kOutputs; ++k) {
                    optr[k] += outs.pitch_elems;
                    if (st && (Op::kOutputs == 1 || outs.p[k] != nullptr))
                        store4<TO>(optr[k], o[k], vec_ok != 0, nvalid);
                }
            }
            __syncwarp();  // every lane is done with this stage before it is refilled
        }
        cp_async_wait<0>();
    }
}

// ----------------------------------------------------------------------------- host launcher

template <typename Op, int ROWS, int STAGES>
int launch_stencil3(const typename Op::in_t *in, int64_t in_pitch_bytes, const typename Op::Params &prm,
                    typename Op::out_t *const *out_ptrs, int64_t out_pitch_bytes, int64_t H, int64_t W,
                    cudaStream_t stream) {
    using T = typename Op::in_t;
    using TO = typename Op::out_t;
    if (H <= 0 || W <= 0) return XRS_OK;  // empty raster: nothing to do
    XRS_REQUIRE(in != nullptr, "input pointer is NULL");
    XRS_REQUIRE(in_pitch_bytes % (int64_t)sizeof(T) == 0 && in_pitch_bytes >= W * (int64_t)sizeof(T),
                "input pitch must be a multiple of the element size and >= row bytes");
    XRS_REQUIRE(out_pitch_bytes % (int64_t)sizeof(TO) == 0 && out_pitch_bytes >= W * (int64_t)sizeof(TO),
                "output pitch must be a multiple of the element size and >= row bytes");
    XRS_REQUIRE(H < (1LL << 31) - 8 && W < (1LL << 31) - 256, "raster dimension too large");

    OutPtrs<Op> outs;
    bool any = false, out_vec_ok = (out_pitch_bytes % 16 == 0);
    for (int k = 0; k < Op::kOutputs; ++k) {
        outs.p[k] = out_ptrs[k];
        if (out_ptrs[k]) {
            any = true;
            XRS_REQUIRE((const void *)out_ptrs[k] != (const void *)in, "in and out must not alias");
            if (reinterpret_cast<uintptr_t>(out_ptrs[k]) % 16 != 0) out_vec_ok = false;
        }
    }
    XRS_REQUIRE(any, "no output pointer given");
    outs.pitch_elems = out_pitch_bytes / (int64_t)sizeof(TO);

    const int sms = sm_count();
    StripGeom g;
    g.H = H;
    g.W = W;
    g.n_strips = (int)((W + kStripW - 1) / kStripW);
    // Row segments: enough tasks for ~8 per resident warp, but segments of >= 64 rows so
    // the 2 halo rows re-read per segment stay a ~3% overhead.
    const int64_t resident_warps = (int64_t)sms * 2 * kWarpsPerCta;
    int64_t want_segs = (resident_warps * 8 + g.n_strips - 1) / g.n_strips;
    int64_t seg_rows = (H + want_segs - 1) / (want_segs > 0 ? want_segs : 1);
    if (seg_rows < 64) seg_rows = 64;
    if (seg_rows > H) seg_rows = H;
    // multiple of ROWS so the last TMA chunk of a segment wastes < ROWS rows
    seg_rows = ((seg_rows + 2 + ROWS - 1) / ROWS) * ROWS - 2;
    if (seg_rows < 1) seg_rows = 1;
    g.seg_rows = (int)seg_rows;
    g.n_segs = (int)((H + seg_rows - 1) / seg_rows);
    const int64_t n_tasks = (int64_t)g.n_strips * g.n_segs;
    int64_t ctas_needed = (n_tasks + kWarpsPerCta - 1) / kWarpsPerCta;

    LaunchInfo &li = last_launch_info();
    li.block = kWarpsPerCta * 32;

    CUtensorMap tmap;
    const bool tma_ok = out_vec_ok && (W % 4 == 0) &&
                        make_tensor_map_2d(&tmap, in, in_pitch_bytes, H, W, (int)sizeof(T), kBoxW, ROWS);
    if (tma_ok) {
        constexpr size_t smem = (size_t)kWarpsPerCta * STAGES * ROWS * kBoxW * sizeof(T) +
                                (size_t)kWarpsPerCta * STAGES * sizeof(uint64_t);
        auto kern = stencil3_tma_kernel<Op, ROWS, STAGES>;
        XRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // persistent grid: as many CTAs as fit on an SM (registers / shared memory), times the SMs
        int per_sm = 0;
        XRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarpsPerCta * 32, smem));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2) per_sm = 2;  // measured: 3 CTAs / SM is ~4% slower for hillshade than 2
        int64_t grid = (int64_t)sms * per_sm;
        if (grid > ctas_needed) grid = ctas_needed;
        li.used_tma = 1;
        li.grid = (int)grid;
        li.smem_bytes = (int)smem;
        kern<<<(unsigned)grid, kWarpsPerCta * 32, smem, stream>>>(tmap, prm, outs, g);
    } else {
        // smaller ring than the TMA path's is not needed: same geometry, cp.async fill
        constexpr size_t smem = (size_t)kWarpsPerCta * STAGES * ROWS * kBoxW * sizeof(T);
        auto kern = stencil3_cpasync_kernel<Op, ROWS, STAGES>;
        XRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        XRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarpsPerCta * 32, smem));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2) per_sm = 2;
        int64_t grid = (int64_t)sms * per_sm;
        if (grid > ctas_needed) grid = ctas_needed;
        li.used_tma = 0;
        li.grid = (int)grid;
        li.smem_bytes = (int)smem;
        kern<<<(unsigned)grid, kWarpsPerCta * 32, smem, stream>>>(in, in_pitch_bytes / (int64_t)sizeof(T), prm, outs, g,
                                                                out_vec_ok ? 1 : 0);
    }
    XRS_CUDA(cudaGetLastError());
    return XRS_OK;
}

}  // namespace xrs
