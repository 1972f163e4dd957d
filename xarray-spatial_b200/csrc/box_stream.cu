// box_stream.cu -- convolve_2d with a kernel whose taps are all the same weight w
// (np.ones((k, k)) / k**2: the mean filter of the reference's docs and of benchmarks/; also any
// rectangular kh x kw): out = w * (sum of the window), reference convolution.py:285-313.
//
// A separable RUNNING box on the CTA-wide TMA pipeline of stencil3.cuh: O(1) work per cell for every k
// (the round-1 kernel built a summed-area table per tile in shared memory and was bound by
// shared-memory bandwidth at 0.27-0.33 of the HBM roofline; this one reaches 0.31-0.43 on B200 and is
// bound by the latency of its float64 shuffle / add chains):
//   * one producer warp streams the tile's rows (one row per stage, 224-cell TMA boxes, NaN
//     out-of-raster fill) into a ring of kh + 1 + PREFETCH row slots; consumer warps march down;
//   * vertical: every lane keeps the running column sums V of its 4 columns over the last kh rows
//     in float64 registers:  V += row(y + ry)  ...emit row y...  V -= row(y - ry)   (the row that
//     leaves is still in the ring; a slot is handed back to the producer after it has left);
//   * horizontal: inclusive prefix P of V along the warp's 128 columns (3 adds + a 5-step warp scan),
//     then window = P[x + rx] - P[x - rx - 1]: both operands sit in other lanes' registers at
//     compile-time lane / index offsets (RX is a template parameter), 8 float64 shuffles per lane-row;
//   * each warp scans only its own 128 columns, so it emits the 128 - 2 * pad(rx) columns whose
//     windows it sees completely; neighbouring warps' input strips overlap in shared memory (free),
//     neighbouring tiles overlap by 2 * pad(rx) columns (L2).
// Numerics.  The reference accumulates fma(w, v, acc) tap by tap in float64; here the window sum is
// formed in float64 (sums of float32 cells: rounding ~1e-16 of the window's magnitude) and scaled
// once -- far inside the 1e-5 bar of the float32 result.  Running sums are only trustworthy while
// every cell that entered them is "ordinary": a cell that is NaN, infinite or huge (|v| >= 2^100,
// e.g. a 3.4e38 nodata sentinel, which would wipe out the float64 low bits of V for as long as it
// stays in the window and corrupt it for good when it leaves) is kept OUT of V and counted instead
// (two 16-bit running counts per column, same prefix machinery): a window holding a NaN is NaN like
// the reference's; a window holding an infinite / huge cell is recomputed tap by tap in the
// reference's order from global memory; all other windows never saw the bad cell.  Rows without any
// such cell -- the warp votes once per row -- skip the masking and the counts entirely.
#include "stencil3.cuh"

namespace xrs {

constexpr int kBsBoxW = 224;     // cells per TMA box (896 B rows: 128-byte aligned destinations)
constexpr int kBsWarps = 8;      // consumer warps per CTA
constexpr int kBsMaxK = 25;      // window rows / columns served (ring: kh + 1 + prefetch rows)
constexpr float kBsHuge = 1.2676506e30f;  // 2^100: cells at or above stay out of the running sums

struct BsGeom {
    int64_t H, W;
    int kh, ry;
    int n_tiles, n_segs, seg_rows;
    int ring;       // row slots
    double w;
};

template <int RX> struct BsShape {
    static constexpr int kPad = (RX + 3) / 4 * 4;               // columns a warp cannot emit on each side
    static constexpr int kOutW = kStripW - 2 * kPad;            // columns a warp emits
    static constexpr int kTileOutW = kBsWarps * kOutW;
    static constexpr int kTileInW = kTileOutW + 2 * kPad;
    static constexpr int kNBox = (kTileInW + kBsBoxW - 1) / kBsBoxW;
    static constexpr int kRowCells = kNBox * kBsBoxW;
};

__device__ __noinline__ float bs_direct(const float *__restrict__ in, int64_t pitch_elems, int64_t H, int64_t W,
                                        int64_t y, int64_t x, int kh, int kw, double w) {
    // the reference's tap-order float64 accumulation (convolution.py:303-308); only called for cells
    // off the NaN ring, so every tap is inside the raster
    double acc = 0.0;
    const float *p = in + (y - kh / 2) * pitch_elems + (x - kw / 2);
    for (int ky = 0; ky < kh; ++ky)
        for (int kx = 0; kx < kw; ++kx) acc = fma(w, (double)p[ky * pitch_elems + kx], acc);
    return (float)acc;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// inclusive prefix over the warp's 128 columns of a lane's 4 values (in place)
__device__ __forceinline__ void bs_prefix(double (&p)[4], int lane) {
    p[1] += p[0];
    p[2] += p[1];
    p[3] += p[2];
    double t = p[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
    }
    const double e = t - p[3];  // exclusive offset of this lane
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] += e;
}
// the same for R rows at once: the R scans are independent, so their shuffle / add chains overlap
template <int R> __device__ __forceinline__ void bs_prefix_rows(double (&p)[R][4], int lane) {
    double t[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        p[i][1] += p[i][0];
        p[i][2] += p[i][1];
        p[i][3] += p[i][2];
        t[i] = p[i][3];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double u[R];
#pragma unroll
        for (int i = 0; i < R; ++i) u[i] = __shfl_up_sync(0xffffffffu, t[i], o);
#pragma unroll
        for (int i = 0; i < R; ++i)
            if (lane >= o) t[i] += u[i];
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const double e = t[i] - p[i][3];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[i][j] += e;
    }
}
__device__ __forceinline__ void bs_prefix_u(unsigned (&p)[4], int lane) {
    p[1] += p[0];
    p[2] += p[1];
    p[3] += p[2];
    unsigned t = p[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
    }
    const unsigned e = t - p[3];
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] += e;
}

// window[j] = P[c_j + RX] - P[c_j - RX - 1] for the lane's columns c_j = 4 * lane + j
template <int RX, typename T>
__device__ __forceinline__ void bs_window(const T (&p)[4], int lane, T (&win)[4]) {
    constexpr int q = RX / 4, m = RX % 4;
    T hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        // upper end: column 4 (lane + q) + j + m
        const int ju = j + m;
        const int src_u = lane + q + (ju >= 4 ? 1 : 0);
        const T vu = p[ju & 3];
        // lower end: column 4 (lane - q) + j - m - 1
        const int jl = j - m - 1;  // in [-4, 2]
        const int src_l = lane - q - (jl < 0 ? 1 : 0);
        const T vl = p[(jl + 4) & 3];
        if constexpr (sizeof(T) == 8) {
            hi[j] = shfl_d(vu, src_u);
            lo[j] = shfl_d(vl, src_l);
        } else {
            hi[j] = __shfl_sync(0xffffffffu, vu, src_u);
            lo[j] = __shfl_sync(0xffffffffu, vl, src_l);
        }
        // src_l < 0 (the prefix before the warp's first column, = 0) only reaches an emitting lane when
        // RX is a multiple of 4: lane q, cell 0.  Every other lane with src_l < 0 sits in the pad.
        if constexpr (m == 0) {
            if (j == 0 && src_l < 0) lo[j] = T(0);
        }
        win[j] = hi[j] - lo[j];
    }
}

template <int RX, int KR>
__global__ void __launch_bounds__((kBsWarps + 1) * 32, 2)
box_stream_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ in, int64_t in_pitch_elems,
                  float *__restrict__ out, int64_t out_pitch_elems, const BsGeom g) {
    using S = BsShape<RX>;
    constexpr int kw = 2 * RX + 1;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)g.ring * S::kRowCells * sizeof(float));
    uint64_t *empty = full + g.ring;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < g.ring; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kBsWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t n_tasks = (int64_t)g.n_tiles * g.n_segs;
    const int kh = g.kh, ry = g.ry;

    if (warp == kBsWarps) {
        // ---- producer: rows y0 - ry .. y1 - 1 + ry of every task, one row per slot
        if (lane == 0) {
            int slot = 0;
            uint32_t lap = 0;  // parity of the current lap around the ring
            for (int64_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
                const int seg = (int)(task / g.n_tiles), tile = (int)(task % g.n_tiles);
                const int64_t y0 = (int64_t)seg * g.seg_rows, y1 = min(y0 + (int64_t)g.seg_rows, g.H);
                const int bx = tile * S::kTileOutW - S::kPad;
                const int n_rows = (int)(y1 - y0) + kh - 1;
                for (int r = 0; r < n_rows; ++r) {
                    mbar_wait(&empty[slot], lap ^ 1u);  // a fresh barrier passes the first lap
                    mbar_arrive_expect_tx(&full[slot], (uint32_t)(S::kRowCells * sizeof(float)));
                    float *dst = ring + (size_t)slot * S::kRowCells;
#pragma unroll
                    for (int b = 0; b < S::kNBox; ++b)
                        tma_load_2d(dst + b * kBsBoxW, &tmap, &full[slot], bx + b * kBsBoxW, (int)(y0 - ry) + r);
                    if (++slot == g.ring) { slot = 0; lap ^= 1u; }
                }
            }
        }
        return;
    }

    // ---- consumers
    const int col = warp * S::kOutW + 4 * lane;                         // first of the lane's columns in the row slot
    const int off = (col / kBsBoxW) * kBsBoxW + (col % kBsBoxW);        // == col: boxes of one row are contiguous
    const bool emits = (4 * lane >= S::kPad) && (4 * lane < kStripW - S::kPad);
    int slot_new = 0, slot_old = 0;     // ring positions of the next row to add / to retire
    uint32_t lap_new = 0;
    for (int64_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const int seg = (int)(task / g.n_tiles), tile = (int)(task % g.n_tiles);
        const int64_t y0 = (int64_t)seg * g.seg_rows, y1 = min(y0 + (int64_t)g.seg_rows, g.H);
        const int64_t x = (int64_t)tile * S::kTileOutW - S::kPad + col;  // raster column of the lane's first cell
        const bool store_ok = emits && x >= 0 && x < g.W;                 // W % 4 == 0: all four cells in or out
        float *optr = out + y0 * out_pitch_elems + x;

        double V[4] = {0.0, 0.0, 0.0, 0.0};
        unsigned C[4] = {0u, 0u, 0u, 0u};   // lo 16 bits: NaN cells in the column window, hi 16: infinite / huge
        uint32_t dirty = 0;                  // bit i: the row added i steps ago held a NaN / inf / huge cell (this warp)
        const int n_rows = (int)(y1 - y0) + kh - 1;
        for (int r = 0; r < n_rows; ++r) {
            // ---- fast path: kRows rows at a time once the window is full, as long as neither the rows
            // entering nor the rows in the window hold a NaN / inf / huge cell.  The kRows prefix scans
            // are independent and run interleaved: one scan per row left the float64 units idle behind a
            // ~600-cycle shuffle / add chain (measured 0.18 of the HBM roofline at k = 25).
            constexpr int kRows = KR;
            while (r >= kh - 1 && r + kRows <= n_rows &&
                   (dirty & (((kh - 1) >= 32) ? 0xffffffffu : ((1u << (kh - 1)) - 1u))) == 0u) {
                float nv[kRows][4], ov[kRows][4];
                int sn = slot_new;
                uint32_t ln = lap_new;
                float amax = 0.f;
                bool anynan = false;
#pragma unroll
                for (int i = 0; i < kRows; ++i) {
                    mbar_wait(&full[sn], ln);
                    const float4 q = *reinterpret_cast<const float4 *>(ring + (size_t)sn * S::kRowCells + off);
                    nv[i][0] = q.x; nv[i][1] = q.y; nv[i][2] = q.z; nv[i][3] = q.w;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        amax = fmaxf(amax, fabsf(nv[i][j]));
                        anynan = anynan || (nv[i][j] != nv[i][j]);
                    }
                    if (++sn == g.ring) { sn = 0; ln ^= 1u; }
                }
                if (__any_sync(0xffffffffu, anynan || !(amax < kBsHuge))) break;   // row by row below
                int so = slot_old;
#pragma unroll
                for (int i = 0; i < kRows; ++i) {
                    const float4 q = *reinterpret_cast<const float4 *>(ring + (size_t)so * S::kRowCells + off);
                    ov[i][0] = q.x; ov[i][1] = q.y; ov[i][2] = q.z; ov[i][3] = q.w;
                    if (++so == g.ring) so = 0;
                }
                // window sums of row i: V_i = V_{i-1} - old_{i-1} + new_i  (the rows in the window are clean:
                // they were added on an unmasked path, so unmasked removal is exact bookkeeping)
                double P[kRows][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double acc = V[j];
#pragma unroll
                    for (int i = 0; i < kRows; ++i) {
                        const double d = (i == 0) ? (double)nv[0][j] : ((double)nv[i][j] - (double)ov[i - 1][j]);
                        acc += d;
                        P[i][j] = acc;
                    }
                    V[j] = acc - (double)ov[kRows - 1][j];
                }
                bs_prefix_rows<kRows>(P, lane);
#pragma unroll
                for (int i = 0; i < kRows; ++i) {
                    double win[4];
                    bs_window<RX, double>(P[i], lane, win);
                    if (store_ok)
                        __stcs(reinterpret_cast<float4 *>(optr),
                               make_float4((float)fma(g.w, win[0], 0.0), (float)fma(g.w, win[1], 0.0),
                                           (float)fma(g.w, win[2], 0.0), (float)fma(g.w, win[3], 0.0)));
                    optr += out_pitch_elems;
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kRows; ++i) {
                    if (lane == 0) mbar_arrive(&empty[slot_old]);
                    if (++slot_old == g.ring) slot_old = 0;
                }
                slot_new = sn;
                lap_new = ln;
                dirty <<= kRows;
                r += kRows;
            }
            if (r >= n_rows) break;
            // ---- add input row y0 - ry + r
            mbar_wait(&full[slot_new], lap_new);
            const float4 q = *reinterpret_cast<const float4 *>(ring + (size_t)slot_new * S::kRowCells + off);
            const float v[4] = {q.x, q.y, q.z, q.w};
            const float amax = fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3])));
            // fmaxf drops NaN operands, so test them separately: v != v for any of the four
            const bool bad_lane = !(amax < kBsHuge) || (v[0] != v[0]) || (v[1] != v[1]) || (v[2] != v[2]) || (v[3] != v[3]);
            const bool row_dirty = __any_sync(0xffffffffu, bad_lane);
            dirty = (dirty << 1) | (row_dirty ? 1u : 0u);
            if (!row_dirty) {
#pragma unroll
                for (int j = 0; j < 4; ++j) V[j] += (double)v[j];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool isnan_ = v[j] != v[j];
                    const bool big = !isnan_ && !(fabsf(v[j]) < kBsHuge);
                    V[j] += (isnan_ || big) ? 0.0 : (double)v[j];
                    C[j] += isnan_ ? 1u : (big ? 0x10000u : 0u);
                }
            }
            if (++slot_new == g.ring) { slot_new = 0; lap_new ^= 1u; }
            if (r < kh - 1) continue;   // window not complete yet

            // ---- emit output row y = y0 + r - (kh - 1)
            const bool win_dirty = (dirty & ((kh >= 32) ? 0xffffffffu : ((1u << kh) - 1u))) != 0u;
            double P[4] = {V[0], V[1], V[2], V[3]};
            bs_prefix(P, lane);
            double win[4];
            bs_window<RX, double>(P, lane, win);
            float res[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j] = (float)fma(g.w, win[j], 0.0);
            if (win_dirty) {   // warp-uniform
                unsigned PC[4] = {C[0], C[1], C[2], C[3]};
                bs_prefix_u(PC, lane);
                unsigned wc[4];
                bs_window<RX, unsigned>(PC, lane, wc);
                const int64_t y = y0 + r - (kh - 1);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (wc[j] & 0xffffu) res[j] = nan_of<float>();
                    else if (wc[j] >> 16) {
                        if (store_ok) res[j] = bs_direct(in, in_pitch_elems, g.H, g.W, y, x + j, kh, kw, g.w);
                    }
                }
            }
            if (store_ok) __stcs(reinterpret_cast<float4 *>(optr), make_float4(res[0], res[1], res[2], res[3]));
            optr += out_pitch_elems;

            // ---- retire input row y - ry (the oldest row of the window), then hand its slot back
            const float4 qo = *reinterpret_cast<const float4 *>(ring + (size_t)slot_old * S::kRowCells + off);
            const float vo[4] = {qo.x, qo.y, qo.z, qo.w};
            const bool old_dirty = (dirty >> (kh - 1)) & 1u;
            if (!old_dirty) {
#pragma unroll
                for (int j = 0; j < 4; ++j) V[j] -= (double)vo[j];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool isnan_ = vo[j] != vo[j];
                    const bool big = !isnan_ && !(fabsf(vo[j]) < kBsHuge);
                    V[j] -= (isnan_ || big) ? 0.0 : (double)vo[j];
                    C[j] -= isnan_ ? 1u : (big ? 0x10000u : 0u);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot_old]);
            if (++slot_old == g.ring) slot_old = 0;
        }
        // the last kh - 1 rows of the task were added but never retired: hand their slots back
        for (int i = 0; i < kh - 1; ++i) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot_old]);
            if (++slot_old == g.ring) slot_old = 0;
        }
    }
}

template <int RX, int KR>
static int launch_box_stream(const CUtensorMap &tmap, const float *in, int64_t in_pitch, float *out, int64_t out_pitch,
                             int64_t H, int64_t W, int kh, double w, cudaStream_t s) {
    using S = BsShape<RX>;
    BsGeom g;
    g.H = H; g.W = W; g.kh = kh; g.ry = kh / 2; g.w = w;
    g.n_tiles = (int)((W + S::kTileOutW - 1) / S::kTileOutW);
    // ring: the window's kh rows, one being retired, and up to 16 rows (~32 KB per CTA) in flight --
    // but never more than 110 KB per CTA, so that two CTAs (16 consumer warps) share an SM: the kernel
    // is latency-bound (float64 shuffle / add chains), and the second CTA's warps matter more than a
    // deeper ring (k = 25 with one 8-warp CTA per SM: 0.18 of the HBM roofline)
    const int row_bytes = S::kRowCells * 4;
    int prefetch = 16, max_ctas = 2;
    if (const char *e = getenv("XRS_BOX_PREFETCH")) prefetch = atoi(e);
    if (const char *e = getenv("XRS_BOX_CTAS")) max_ctas = atoi(e);
    const size_t cap = (size_t)(220 * 1024) / max_ctas;
    g.ring = kh + 1 + prefetch;
    size_t smem = (size_t)g.ring * row_bytes + (size_t)2 * g.ring * sizeof(uint64_t);
    while (smem > cap && g.ring > kh + 5) {
        --g.ring;
        smem = (size_t)g.ring * row_bytes + (size_t)2 * g.ring * sizeof(uint64_t);
    }
    auto kern = box_stream_kernel<RX, KR>;
    XRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    XRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (kBsWarps + 1) * 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > max_ctas) per_sm = max_ctas;
    const int64_t resident = (int64_t)sm_count() * per_sm;
    // segments: ~4 tasks per CTA, but tall enough that the kh - 1 warm-up rows stay a small overhead
    int64_t want_segs = (resident * 4 + g.n_tiles - 1) / g.n_tiles;
    int64_t seg_rows = (H + want_segs - 1) / (want_segs > 0 ? want_segs : 1);
    if (seg_rows < 16 * kh) seg_rows = 16 * kh;
    if (seg_rows > H) seg_rows = H;
    g.seg_rows = (int)seg_rows;
    g.n_segs = (int)((H + seg_rows - 1) / seg_rows);
    const int64_t n_tasks = (int64_t)g.n_tiles * g.n_segs;
    const int64_t grid = resident < n_tasks ? resident : n_tasks;
    kern<<<(unsigned)grid, (kBsWarps + 1) * 32, smem, s>>>(tmap, in, in_pitch / 4, out, out_pitch / 4, g);
    last_launch_info() = {3, (int)grid, (kBsWarps + 1) * 32, (int)smem};
    XRS_CUDA(cudaGetLastError());
    return XRS_OK;
}

// ============================================================================================
// Second generation (round 2, second session): the same running box, reorganised around what the
// ncu / SASS reading of the kernel above showed (36 instructions per cell, a third of them register
// moves and the 5-step float64 warp scan; issue slots 35 % busy behind long shuffle / add chains; for
// k = 25 a ring that holds the whole window leaves 4 rows of prefetch):
//   * LANE SUMS instead of a prefix scan.  A lane forms the inclusive prefix `pre` and suffix `suf` of
//     its own 4 column sums and its total; the window of column 4l + j then is
//         suf[.] of the lane its left end falls in + pre[.] of the lane its right end falls in
//         + the totals of the whole lanes in between,
//     every operand at a compile-time lane distance: 4 (k = 5) ... 11 (k = 25) float64 shuffles per
//     lane-row instead of 13, no 5-step dependent scan, no prefix differences.  A NaN that lives in
//     one lane's sums reaches exactly the windows that contain that lane's columns, so the columns
//     beyond the raster's left / right edge (NaN from the TMA unit) need no masking: edge tiles run
//     the fast path and their border windows come out NaN like the reference's.
//   * TWO STREAMS per stage: the 4 rows that enter the window and the 4 rows that leave it (re-read
//     through L2, they were fetched kh rows earlier by the same CTA) arrive as ONE stage with one
//     full / one empty mbarrier -- the ring no longer holds the window, so k = 25 gets the same 3
//     stages of prefetch as k = 5, and a consumer warp waits and arrives once per 4 rows.
//   * one 3-input NaN-propagating |max| chain (FMNMX3.NAN) finds NaN / inf / huge cells.
//   * segments are chosen so that no CTA runs one task more than the others (pick_seg_rows).
constexpr int kB2Rows = 4;              // rows per stage half (entering / leaving)
constexpr int kBoxNotTaken = -12345;

// XRS_BOX_ALGO=1 selects the first-generation (prefix-scan, window-in-ring) kernel, for A/B measurements
static int box_algo() {
    const char *e = getenv("XRS_BOX_ALGO");
    return (e && atoi(e) == 1) ? 1 : 2;
}

template <int RX, int NW> struct B2Shape {
    static constexpr int kPad = (RX + 3) / 4 * 4;               // columns a warp cannot emit on each side
    static constexpr int kOutW = kStripW - 2 * kPad;            // columns a warp emits
    static constexpr int kTileOutW = NW * kOutW;
    static constexpr int kTileInW = kTileOutW + 2 * kPad;
    static constexpr int kNBox = (kTileInW + 255) / 256;
    static constexpr int kBoxW = ((kTileInW + kNBox - 1) / kNBox + 31) / 32 * 32;   // cells: 128-byte multiples
    static constexpr int kBoxCells = kB2Rows * kBoxW;           // one TMA box: 4 rows x kBoxW cells
    static constexpr int kHalfCells = kNBox * kBoxCells;        // the entering (or leaving) rows of a stage
    static constexpr uint32_t kHalfBytes = kHalfCells * 4;
    static_assert(kBoxW <= 256 && kBoxW % 32 == 0 && kNBox * kBoxW >= kTileInW, "TMA box geometry");
};

struct B2Geom {
    int64_t H, W;
    int kh, ry;
    int n_tiles, n_segs, seg_rows;
    int stages;
    double w;
};

// max(|a|, |b|, |c|), NaN if any operand is NaN (FMNMX3.NAN with |.| operand modifiers)
__device__ __forceinline__ float bs_amax3(float a, float b, float c) {
    float r;
    asm("{\n.reg .f32 x, y, z;\nabs.f32 x, %1;\nabs.f32 y, %2;\nabs.f32 z, %3;\nmax.NaN.f32 %0, x, y, z;\n}"
        : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float bs_amax4(const float (&v)[4]) {
    return bs_amax3(bs_amax3(v[0], v[1], v[2]), v[3], v[3]);
}

// the value lane (l + D) holds; lanes past the warp's ends get their own (they sit in the pad)
template <int D, typename T> __device__ __forceinline__ T bs_from(T x) {
    if constexpr (D == 0) return x;
    else if constexpr (D > 0) return __shfl_down_sync(0xffffffffu, x, D);
    else return __shfl_up_sync(0xffffffffu, x, -D);
}

// window sum of column 4 l + J (radius RX) from the lane-local prefix / suffix sums, the lane total,
// `core` = the totals of lanes l-q+1 .. l+q-1, `tl` / `tr` = the totals of lanes l-q / l+q  (q = RX / 4)
template <int RX, int J, typename T>
__device__ __forceinline__ T bs_cell(const T (&pre)[4], const T (&suf)[4], T tot, T core, T tl, T tr) {
    constexpr int q = RX / 4, m = RX % 4;
    constexpr int a = J - m, b = J + m;          // window = [4 (l - q) + a, 4 (l + q) + b]
    if constexpr (q == 0 && a >= 0 && b <= 3) {  // inside the lane (RX = 1 only)
        static_assert(a == 0 || b == 3, "in-lane window");
        if constexpr (a == 0) return pre[b]; else return suf[a];
    } else {
        constexpr int dl = -q - (a < 0 ? 1 : 0), ia = (a + 4) & 3;
        constexpr int dh = q + (b >= 4 ? 1 : 0), ib = b & 3;
        const T ends = bs_from<dl>(suf[ia]) + bs_from<dh>(pre[ib]);
        if constexpr (q == 0) {
            if constexpr (a < 0 && b >= 4) return ends + tot; else return ends;
        } else {
            T full = core;
            if constexpr (a < 0) full = full + tl;
            if constexpr (b >= 4) full = full + tr;
            return ends + full;
        }
    }
}

// win[i][j] = sum of v[i] over columns 4 l + j - RX .. 4 l + j + RX, for R independent rows (their
// shuffle / add chains interleave)
template <int RX, int R, typename T>
__device__ __forceinline__ void bs_lanesum(const T (&v)[R][4], T (&win)[R][4]) {
    static_assert(RX >= 1 && RX <= 12, "window radius");
    constexpr int q = RX / 4, m = RX % 4;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        T pre[4], suf[4];
        pre[0] = v[i][0];
        pre[1] = pre[0] + v[i][1];
        pre[2] = pre[1] + v[i][2];
        pre[3] = pre[2] + v[i][3];
        suf[3] = v[i][3];
        suf[2] = v[i][2] + suf[3];
        suf[1] = v[i][1] + suf[2];
        suf[0] = pre[3];
        const T tot = pre[3];
        T core = tot, tl = T(0), tr = T(0);
        if constexpr (q == 2) core = (bs_from<-1>(tot) + tot) + bs_from<1>(tot);
        if constexpr (q == 3) {
            const T pair = tot + bs_from<1>(tot);                       // lanes l, l + 1
            core = (bs_from<-2>(pair) + pair) + bs_from<2>(tot);        // lanes l - 2 .. l + 2
        }
        if constexpr (q >= 1 && m > 0) {
            tl = bs_from<-q>(tot);
            tr = bs_from<q>(tot);
        }
        win[i][0] = bs_cell<RX, 0, T>(pre, suf, tot, core, tl, tr);
        win[i][1] = bs_cell<RX, 1, T>(pre, suf, tot, core, tl, tr);
        win[i][2] = bs_cell<RX, 2, T>(pre, suf, tot, core, tl, tr);
        win[i][3] = bs_cell<RX, 3, T>(pre, suf, tot, core, tl, tr);
    }
}

__device__ __forceinline__ uint32_t bs_bits(int n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }

template <int RX, int NW>
__global__ void __launch_bounds__((NW + 1) * 32, 2)
box_stream2_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ in, int64_t in_pitch_elems,
                   float *__restrict__ out, int64_t out_pitch_elems, const B2Geom g) {
    using S = B2Shape<RX, NW>;
    constexpr int kw = 2 * RX + 1;
    constexpr int kStageCells = 2 * S::kHalfCells;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)g.stages * kStageCells * sizeof(float));
    uint64_t *empty = full + g.stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t n_tasks = (int64_t)g.n_tiles * g.n_segs;
    const int kh = g.kh, ry = g.ry;

    if (warp == NW) {
        // ---- producer: per task the batches of 4 entering rows e0 .. e0 + 3 (row e = raster row
        // y0 - ry + e) and, once rows leave the window, the 4 rows e0 - (kh - 1) .. that leave with them
        if (lane == 0) {
            int stage = 0;
            uint32_t lap = 0;
            for (int64_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
                const int seg = (int)(task / g.n_tiles), tile = (int)(task % g.n_tiles);
                const int64_t y0 = (int64_t)seg * g.seg_rows, y1 = min(y0 + (int64_t)g.seg_rows, g.H);
                const int bx = tile * S::kTileOutW - S::kPad;
                const int ytop = (int)y0 - ry;
                const int n_in = (int)(y1 - y0) + kh - 1;
                for (int e0 = 0; e0 < n_in; e0 += kB2Rows) {
                    mbar_wait(&empty[stage], lap ^ 1u);  // a fresh barrier passes the first lap
                    const bool leaving = e0 + kB2Rows - 1 >= kh - 1;
                    mbar_arrive_expect_tx(&full[stage], leaving ? 2u * S::kHalfBytes : S::kHalfBytes);
                    float *dst = ring + (size_t)stage * kStageCells;
#pragma unroll
                    for (int b = 0; b < S::kNBox; ++b)
                        tma_load_2d(dst + b * S::kBoxCells, &tmap, &full[stage], bx + b * S::kBoxW, ytop + e0);
                    if (leaving) {
#pragma unroll
                        for (int b = 0; b < S::kNBox; ++b)
                            tma_load_2d(dst + S::kHalfCells + b * S::kBoxCells, &tmap, &full[stage], bx + b * S::kBoxW,
                                        ytop + e0 - (kh - 1));
                    }
                    if (++stage == g.stages) { stage = 0; lap ^= 1u; }
                }
            }
        }
        return;
    }

    // ---- consumers
    const int col = warp * S::kOutW + 4 * lane;                         // first of the lane's columns in the tile row
    const int off = (col / S::kBoxW) * S::kBoxCells + (col % S::kBoxW);  // + i * kBoxW for row i of the half
    const bool emits = (4 * lane >= S::kPad) && (4 * lane < kStripW - S::kPad);
    const uint32_t win_mask = bs_bits(kh - 1);   // the rows in the window before a row is added
    int stage = 0;
    uint32_t lap = 0;
    for (int64_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const int seg = (int)(task / g.n_tiles), tile = (int)(task % g.n_tiles);
        const int64_t y0 = (int64_t)seg * g.seg_rows, y1 = min(y0 + (int64_t)g.seg_rows, g.H);
        const int64_t x = (int64_t)tile * S::kTileOutW - S::kPad + col;  // raster column of the lane's first cell
        const bool in_raster = x >= 0 && x < g.W;                          // W % 4 == 0: all four cells in or out
        const bool store_ok = emits && in_raster;
        float *optr = out + y0 * out_pitch_elems + x;

        double V[4] = {0.0, 0.0, 0.0, 0.0};
        unsigned C[4] = {0u, 0u, 0u, 0u};   // lo 16 bits: NaN cells in the column window, hi 16: infinite / huge
        uint32_t dirty = 0;                  // bit i: the row added i rows ago held a NaN / inf / huge cell (this warp)
        const int n_in = (int)(y1 - y0) + kh - 1;
        for (int e0 = 0; e0 < n_in; e0 += kB2Rows) {
            mbar_wait(&full[stage], lap);
            const float *se = ring + (size_t)stage * kStageCells + off;   // entering rows
            const float *sl = se + S::kHalfCells;                          // leaving rows
            bool done = false;
            if (e0 >= kh - 1 && e0 + kB2Rows <= n_in && (dirty & win_mask) == 0u) {
                // ---- fast path: window complete, 4 rows enter, 4 rows are emitted, 4 rows leave; neither
                // the window nor the entering rows hold a NaN / inf / huge cell inside the raster
                float nv[kB2Rows][4];
#pragma unroll
                for (int i = 0; i < kB2Rows; ++i) {
                    const float4 q = *reinterpret_cast<const float4 *>(se + i * S::kBoxW);
                    nv[i][0] = q.x; nv[i][1] = q.y; nv[i][2] = q.z; nv[i][3] = q.w;
                }
                const float amax = bs_amax3(bs_amax3(bs_amax4(nv[0]), nv[1][0], nv[1][1]),
                                            bs_amax3(bs_amax4(nv[2]), nv[1][2], nv[1][3]), bs_amax4(nv[3]));
                if (!__any_sync(0xffffffffu, in_raster && !(amax < kBsHuge))) {
                    float ov[kB2Rows][4];
#pragma unroll
                    for (int i = 0; i < kB2Rows; ++i) {
                        const float4 q = *reinterpret_cast<const float4 *>(sl + i * S::kBoxW);
                        ov[i][0] = q.x; ov[i][1] = q.y; ov[i][2] = q.z; ov[i][3] = q.w;
                    }
                    // column sums of output row i: V_i = V_{i-1} - old_{i-1} + new_i
                    double P[kB2Rows][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double acc = V[j];
#pragma unroll
                        for (int i = 0; i < kB2Rows; ++i) {
                            const double d = (i == 0) ? (double)nv[0][j] : ((double)nv[i][j] - (double)ov[i > 0 ? i - 1 : 0][j]);
                            acc += d;
                            P[i][j] = acc;
                        }
                        V[j] = acc - (double)ov[kB2Rows - 1][j];
                    }
                    double win[kB2Rows][4];
                    bs_lanesum<RX, kB2Rows, double>(P, win);
#pragma unroll
                    for (int i = 0; i < kB2Rows; ++i) {
                        if (store_ok)
                            __stcs(reinterpret_cast<float4 *>(optr),
                                   make_float4((float)fma(g.w, win[i][0], 0.0), (float)fma(g.w, win[i][1], 0.0),
                                               (float)fma(g.w, win[i][2], 0.0), (float)fma(g.w, win[i][3], 0.0)));
                        optr += out_pitch_elems;
                    }
                    dirty <<= kB2Rows;
                    done = true;
                }
            }
            if (!done) {
                // ---- row by row: lead-in rows (nothing to emit yet), the last rows of a task, and windows
                // or entering rows with NaN / inf / huge cells (kept out of V, counted in C)
                for (int i = 0; i < kB2Rows; ++i) {
                    const int e = e0 + i;
                    if (e >= n_in) break;
                    const float4 q = *reinterpret_cast<const float4 *>(se + i * S::kBoxW);
                    const float v[4] = {q.x, q.y, q.z, q.w};
                    const bool row_dirty = __any_sync(0xffffffffu, in_raster && !(bs_amax4(v) < kBsHuge));
                    dirty = (dirty << 1) | (row_dirty ? 1u : 0u);
                    if (!row_dirty) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) V[j] += (double)v[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const bool isnan_ = v[j] != v[j];
                            const bool big = !isnan_ && !(fabsf(v[j]) < kBsHuge);
                            V[j] += (isnan_ || big) ? 0.0 : (double)v[j];
                            C[j] += isnan_ ? 1u : (big ? 0x10000u : 0u);
                        }
                    }
                    if (e < kh - 1) continue;   // window not complete yet

                    // emit output row y = y0 + e - (kh - 1)
                    const bool win_dirty = (dirty & bs_bits(kh)) != 0u;
                    double P1[1][4] = {{V[0], V[1], V[2], V[3]}}, w1[1][4];
                    bs_lanesum<RX, 1, double>(P1, w1);
                    float res[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) res[j] = (float)fma(g.w, w1[0][j], 0.0);   // + 0.0: an all-zero window is +0 like the reference's
                    if (win_dirty) {   // warp-uniform
                        unsigned C1[1][4] = {{C[0], C[1], C[2], C[3]}}, wc[1][4];
                        bs_lanesum<RX, 1, unsigned>(C1, wc);
                        const int64_t y = y0 + e - (kh - 1);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (wc[0][j] & 0xffffu) res[j] = nan_of<float>();
                            else if (wc[0][j] >> 16) {
                                if (store_ok) res[j] = bs_direct(in, in_pitch_elems, g.H, g.W, y, x + j, kh, kw, g.w);
                            }
                        }
                    }
                    if (store_ok) __stcs(reinterpret_cast<float4 *>(optr), make_float4(res[0], res[1], res[2], res[3]));
                    optr += out_pitch_elems;

                    // retire input row e - (kh - 1), the oldest row of the window
                    const float4 qo = *reinterpret_cast<const float4 *>(sl + i * S::kBoxW);
                    const float vo[4] = {qo.x, qo.y, qo.z, qo.w};
                    const bool old_dirty = (dirty >> (kh - 1)) & 1u;
                    if (!old_dirty) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) V[j] -= (double)vo[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const bool isnan_ = vo[j] != vo[j];
                            const bool big = !isnan_ && !(fabsf(vo[j]) < kBsHuge);
                            V[j] -= (isnan_ || big) ? 0.0 : (double)vo[j];
                            C[j] -= isnan_ ? 1u : (big ? 0x10000u : 0u);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == g.stages) { stage = 0; lap ^= 1u; }
        }
    }
}

template <int RX, int NW>
static int launch_box_stream2(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H, int64_t W,
                              int kh, double w, cudaStream_t s) {
    using S = B2Shape<RX, NW>;
    CUtensorMap tmap;
    if (!make_tensor_map_2d(&tmap, in, in_pitch, H, W, 4, S::kBoxW, kB2Rows)) return kBoxNotTaken;
    B2Geom g;
    g.H = H; g.W = W; g.kh = kh; g.ry = kh / 2; g.w = w;
    g.n_tiles = (int)((W + S::kTileOutW - 1) / S::kTileOutW);
    int stages = RX <= 2 ? 3 : 4, max_ctas = 2, want = 4;   // B200 sweep: profiles/r02s2_box_sweep.txt
    if (const char *e = getenv("XRS_BOX_STAGES")) stages = atoi(e);
    if (const char *e = getenv("XRS_BOX_CTAS")) max_ctas = atoi(e);
    if (const char *e = getenv("XRS_BOX_WAVES")) want = atoi(e);
    if (max_ctas < 1) max_ctas = 1;
    if (want < 1) want = 1;
    const size_t stage_bytes = (size_t)2 * S::kHalfBytes;
    // shared memory of an SM: 228 KB, 1 KB of it reserved per resident CTA
    const size_t cap = ((size_t)228 * 1024 - (size_t)max_ctas * 1024) / max_ctas - 256;
    if (stages < 2) stages = 2;
    while (stages > 2 && (size_t)stages * stage_bytes + (size_t)2 * stages * sizeof(uint64_t) > cap) --stages;
    g.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + (size_t)2 * stages * sizeof(uint64_t);
    auto kern = box_stream2_kernel<RX, NW>;
    XRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    XRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (NW + 1) * 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > max_ctas) per_sm = max_ctas;
    const int64_t resident = (int64_t)sm_count() * per_sm;
    // segments: tall enough that the kh - 1 lead-in rows stay a small overhead, (rows + kh - 1) a
    // multiple of 4 so that only the raster's last segment ends in a partial batch
    const int64_t seg_rows = pick_seg_rows(H, g.n_tiles, resident, 12 * (int64_t)kh, kh - 1, kB2Rows, want);
    g.seg_rows = (int)seg_rows;
    g.n_segs = (int)((H + seg_rows - 1) / seg_rows);
    const int64_t n_tasks = (int64_t)g.n_tiles * g.n_segs;
    const int64_t grid = resident < n_tasks ? resident : n_tasks;
    kern<<<(unsigned)grid, (NW + 1) * 32, smem, s>>>(tmap, in, in_pitch / 4, out, out_pitch / 4, g);
    last_launch_info() = {3, (int)grid, (NW + 1) * 32, (int)smem};
    XRS_CUDA(cudaGetLastError());
    return XRS_OK;
}

// true when the streaming box kernel took the job (*rc = its status)
bool try_box_stream(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H, int64_t W,
                    const double *kernel, int kh, int kw, cudaStream_t s, int *rc) {
    if (kh > kBsMaxK || kw > kBsMaxK || kh < 1 || kw < 3 || (kw & 1) == 0 || (kh & 1) == 0) return false;
    const double w = kernel[0];
    if (!(fabs(w) <= 1.7976931348623157e308)) return false;
    for (int i = 1; i < kh * kw; ++i)
        if (memcmp(&kernel[i], &w, sizeof(double)) != 0) return false;
    if (W % 4 != 0 || out_pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
    if (H >= (1LL << 31) - 64 || W >= (1LL << 31) - 4096) return false;
    if (box_algo() == 2) {
        // 7 consumer warps + the producer = 256 threads: 128 registers per thread at two CTAs per SM (no
        // spills; 8 + 1 warps are capped at 96 and spill in the fast path: 0.79 -> 0.87 of HBM at k = 9)
        int nw = 7;
        if (const char *e = getenv("XRS_BOX_WARPS")) nw = atoi(e);
        switch (kw / 2) {
#define XRS_BS(R) case R: { const int r2 = nw == 8 ? launch_box_stream2<R, 8>(in, in_pitch, out, out_pitch, H, W, kh, w, s)  \
                                                     : launch_box_stream2<R, 7>(in, in_pitch, out, out_pitch, H, W, kh, w, s); \
                            if (r2 != kBoxNotTaken) { *rc = r2; return true; } } break;
            XRS_BS(1) XRS_BS(2) XRS_BS(3) XRS_BS(4) XRS_BS(5) XRS_BS(6) XRS_BS(7) XRS_BS(8) XRS_BS(9) XRS_BS(10) XRS_BS(11) XRS_BS(12)
#undef XRS_BS
        }
    }
    CUtensorMap tmap;
    if (!make_tensor_map_2d(&tmap, in, in_pitch, H, W, 4, kBsBoxW, 1)) return false;
    switch (kw / 2) {
#define XRS_BS(R) case R: *rc = launch_box_stream<R, 4>(tmap, in, in_pitch, out, out_pitch, H, W, kh, w, s); return true;
        XRS_BS(1) XRS_BS(2) XRS_BS(3) XRS_BS(4) XRS_BS(5) XRS_BS(6) XRS_BS(7) XRS_BS(8) XRS_BS(9) XRS_BS(10) XRS_BS(11) XRS_BS(12)
#undef XRS_BS
    }
    return false;
}

}  // namespace xrs
