// conv.cu -- general k x k neighbourhood operators: convolve_2d (convolution.py:285-313) and
// the masked focal statistics behind focal.apply / focal_stats (focal.py:305-326, 268-302).
//
// One CTA (256 threads = 32 x 8) per 128 x 32 output tile, persistent over tiles.  The
// (tile + halo) block is brought into shared memory by TMA 2-D bulk loads with NaN
// out-of-bounds fill: for convolve_2d that produces the reference's NaN ring for free
// (NaN * w = NaN, even for w = 0), for the focal statistics out-of-raster cells are skipped
// exactly like NaN cells, which is the reference's clamped window.
//
// convolve_2d: the tile is widened to float64 once in shared memory; every thread owns a
// 4 x 4 block of outputs and accumulates in float64 with FMA (the reference accumulates
// `kernel(f64) * data(f32)` in a float64 `num`); weights are read from the kernel-parameter
// constant bank.  Bound: FP64 FMA rate for k >= 5 (2*k*k flop/cell), HBM for k = 3.
#include "common.cuh"

namespace xrs {

constexpr int kMaxTaps = 2401;  // up to 49 x 49
constexpr int kTileW = 128, kTileH = 32;

struct ConvWeights {
    double w[kMaxTaps];
};
struct MaskBits {
    unsigned char m[kMaxTaps];
};

struct TileGeom {
    int64_t H, W;
    int kh, kw, ry, rx;
    int pad;       // cells loaded left of the tile: rx rounded up to 4 (TMA box starts must stay
                   // 16-byte aligned in the innermost dimension)
    int off;       // pad - rx: column offset of tap 0 inside the shared tile
    int sw;        // shared tile width (cells), multiple of 4
    int sh;        // shared tile height = kTileH + kh - 1
    int tiles_x, tiles_y;
    int box_h;     // rows per TMA box (sh is loaded in ceil(sh / box_h) boxes)
};

__device__ __forceinline__ void load_tile_tma(const CUtensorMap *tmap, float *tile32, uint64_t *bar,
                                              const TileGeom &g, int tx0, int ty0, uint32_t parity) {
    // one elected thread issues the boxes; everyone waits on the mbarrier
    if (threadIdx.x == 0) {
        const int nbox = (g.sh + g.box_h - 1) / g.box_h;
        mbar_arrive_expect_tx(bar, (uint32_t)(nbox * g.box_h * g.sw * sizeof(float)));
        for (int b = 0; b < nbox; ++b)
            tma_load_2d(tile32 + (size_t)b * g.box_h * g.sw, tmap, bar, tx0 - g.pad, ty0 - g.ry + b * g.box_h);
    }
    mbar_wait(bar, parity);
}

// ----------------------------------------------------------------------------- convolve
// 128 x 64 output tile per CTA, 256 threads = 16 (x) x 16 (y); a thread owns 4 rows x (4 + 4)
// columns: columns 4tx .. 4tx+3 and 64+4tx .. 64+4tx+3, so that consecutive lanes read
// consecutive 16-byte pieces of a shared-memory row (conflict-free LDS.128).  The float32 tile
// is read directly and widened in registers (16 F2F per 128 DFMA); 32 independent float64
// accumulators per thread give the FP64 pipe enough parallelism at 2-3 CTAs per SM.
constexpr int kConvTileH = 64;

__global__ void __launch_bounds__(256)
conv2d_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ConvWeights cw,
              float *__restrict__ out, int64_t out_pitch_elems, const TileGeom g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int nbox = (g.sh + g.box_h - 1) / g.box_h;
    const size_t tile_cells = (size_t)nbox * g.box_h * g.sw;
    float *tile32 = reinterpret_cast<float *>(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + tile_cells * sizeof(float));

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t n_tiles = (int64_t)g.tiles_x * g.tiles_y;
    uint32_t parity = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int tile_y = (int)(t / g.tiles_x), tile_x = (int)(t % g.tiles_x);
        const int x0 = tile_x * kTileW, y0 = tile_y * kConvTileH;
        load_tile_tma(&tmap, tile32, bar, g, x0, y0, parity);
        parity ^= 1u;

        double acc[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;

        // Input row j of the thread's window feeds output row r with kernel row ky = j - r.
        // Rows 3 .. kh-1 feed all four output rows, and 4-tap chunks lying fully inside the
        // kernel need no tap test: that common case is straight-line DFMA code; the few edge
        // rows / edge chunks take the generic, predicated path.  Taps are indexed from the
        // 16-byte aligned tile origin: tap kx sits at column kx + off.
        const int rows_in = 4 + g.kh - 1;
        const int n_taps = g.off + g.kw;
        for (int j = 0; j < rows_in; ++j) {
            const float *rowp = tile32 + (size_t)(ty * 4 + j) * g.sw + 4 * tx;
            const bool full_rows = (j >= 3) && (j < g.kh);
            for (int kb = 0; kb < n_taps; kb += 4) {
                const float4 a0 = *reinterpret_cast<const float4 *>(rowp + kb);
                const float4 a1 = *reinterpret_cast<const float4 *>(rowp + kb + 4);
                const float4 b0 = *reinterpret_cast<const float4 *>(rowp + 64 + kb);
                const float4 b1 = *reinterpret_cast<const float4 *>(rowp + 64 + kb + 4);
                const double va[8] = {(double)a0.x, (double)a0.y, (double)a0.z, (double)a0.w,
                                      (double)a1.x, (double)a1.y, (double)a1.z, (double)a1.w};
                const double vb[8] = {(double)b0.x, (double)b0.y, (double)b0.z, (double)b0.w,
                                      (double)b1.x, (double)b1.y, (double)b1.z, (double)b1.w};
                const bool full_chunk = (kb >= g.off) && (kb + 4 <= n_taps);
                if (full_chunk && full_rows) {
                    const double *w0 = cw.w + (j * g.kw + kb - g.off);  // kernel row j, taps kb-off ..
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const double *wr = w0 - r * g.kw;            // kernel row j - r
#pragma unroll
                        for (int tt = 0; tt < 4; ++tt) {
                            const double wv = wr[tt];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                acc[r][c] = fma(wv, va[c + tt], acc[r][c]);
                                acc[r][4 + c] = fma(wv, vb[c + tt], acc[r][4 + c]);
                            }
                        }
                    }
                } else if (full_chunk) {
                    // lead-in / lead-out rows of the window: some output rows have no kernel row here
                    const double *w0 = cw.w + (j * g.kw + kb - g.off);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (j - r >= 0 && j - r < g.kh) {  // warp-uniform
                            const double *wr = w0 - r * g.kw;        // kernel row j - r
#pragma unroll
                            for (int tt = 0; tt < 4; ++tt) {
                                const double wv = wr[tt];
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    acc[r][c] = fma(wv, va[c + tt], acc[r][c]);
                                    acc[r][4 + c] = fma(wv, vb[c + tt], acc[r][4 + c]);
                                }
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int ky = j - r;
                        if (ky >= 0 && ky < g.kh) {
#pragma unroll
                            for (int tt = 0; tt < 4; ++tt) {
                                const int kx = kb + tt - g.off;
                                if (kx >= 0 && kx < g.kw) {
                                    const double wv = cw.w[ky * g.kw + kx];
#pragma unroll
                                    for (int c = 0; c < 4; ++c) {
                                        acc[r][c] = fma(wv, va[c + tt], acc[r][c]);
                                        acc[r][4 + c] = fma(wv, vb[c + tt], acc[r][4 + c]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        const int64_t xa = (int64_t)x0 + 4 * tx, xb = xa + 64;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int64_t yo = (int64_t)y0 + ty * 4 + r;
            if (yo < g.H) {  // W % 4 == 0 on this path: a float4 is all-in or all-out
                float *orow = out + yo * out_pitch_elems;
                if (xa < g.W)
                    __stcs(reinterpret_cast<float4 *>(orow + xa),
                           make_float4((float)acc[r][0], (float)acc[r][1], (float)acc[r][2], (float)acc[r][3]));
                if (xb < g.W)
                    __stcs(reinterpret_cast<float4 *>(orow + xb),
                           make_float4((float)acc[r][4], (float)acc[r][5], (float)acc[r][6], (float)acc[r][7]));
            }
        }
        __syncthreads();  // the tile buffer is reused by the next iteration
    }
}

// Fallback for rasters TMA cannot describe: one thread per cell, bounds-checked loads.
__global__ void __launch_bounds__(256)
conv2d_direct_kernel(const float *__restrict__ in, int64_t in_pitch_elems, const __grid_constant__ ConvWeights cw,
                     float *__restrict__ out, int64_t out_pitch_elems, int64_t H, int64_t W, int kh, int kw) {
    const int64_t n = H * W;
    const int ry = kh / 2, rx = kw / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t y = i / W, x = i % W;
        double acc = 0.0;
        for (int ky = 0; ky < kh; ++ky)
            for (int kx = 0; kx < kw; ++kx) {
                const int64_t yy = y + ky - ry, xx = x + kx - rx;
                const double v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (double)in[yy * in_pitch_elems + xx]
                                                                            : nan_of<double>();
                acc = fma(cw.w[ky * kw + kx], v, acc);
            }
        out[y * out_pitch_elems + x] = (float)acc;
    }
}

// ----------------------------------------------------------------------------- focal statistics
// Reducers follow Numba's nan-functions (numba/np/arraymath.py) as used by focal.py:268-302:
// mean/var/std accumulate in f64 (var two-pass about the f64 mean), sum accumulates in f32 in
// row-major window order (bit-identical to np.nansum on the f32 scratch), min/max skip NaN.
template <typename Fetch>
__device__ __forceinline__ float focal_reduce(const Fetch &fetch, const MaskBits &mask, int kh, int kw, int stat) {
    if (stat == XRS_STAT_MEAN || stat == XRS_STAT_VAR || stat == XRS_STAT_STD) {
        double c = 0.0;
        int cnt = 0;
        for (int ky = 0; ky < kh; ++ky)
            for (int kx = 0; kx < kw; ++kx)
                if (mask.m[ky * kw + kx]) {
                    const float v = fetch(ky, kx);
                    if (v == v) { c += (double)v; ++cnt; }
                }
        const double m = c / (double)cnt;
        if (stat == XRS_STAT_MEAN) return (float)m;
        double ssd = 0.0;
        for (int ky = 0; ky < kh; ++ky)
            for (int kx = 0; kx < kw; ++kx)
                if (mask.m[ky * kw + kx]) {
                    const float v = fetch(ky, kx);
                    if (v == v) { const double d = (double)v - m; ssd += d * d; }
                }
        const double var = ssd / (double)cnt;
        return (float)(stat == XRS_STAT_VAR ? var : sqrt(var));
    } else if (stat == XRS_STAT_SUM) {
        float c = 0.f;
        for (int ky = 0; ky < kh; ++ky)
            for (int kx = 0; kx < kw; ++kx)
                if (mask.m[ky * kw + kx]) {
                    const float v = fetch(ky, kx);
                    if (v == v) c += v;
                }
        return c;
    } else {
        // nanmin / nanmax start from scratch[0] (NaN unless the first window cell takes part)
        float mn = nan_of<float>(), mx = nan_of<float>();
        for (int ky = 0; ky < kh; ++ky)
            for (int kx = 0; kx < kw; ++kx)
                if (mask.m[ky * kw + kx]) {
                    const float v = fetch(ky, kx);
                    if (v == v) {
                        if (!(mn < v)) mn = v;
                        if (!(mx > v)) mx = v;
                    }
                }
        return stat == XRS_STAT_MIN ? mn : stat == XRS_STAT_MAX ? mx : mx - mn;
    }
}

// Register-blocked like conv2d_kernel: a thread owns 4 x 4 outputs, walks the input rows of its
// window once (twice for var / std), loads 8 cells per 4-tap chunk with two LDS.128 and feeds
// every (output row, tap) pair whose mask bit is set.  Taps are visited in row-major window
// order for each output, so the float32 `sum` is bit-identical to np.nansum on the scratch.
template <typename F>
__device__ __forceinline__ void focal_sweep(const float *tile32, const MaskBits &mask, const TileGeom &g, int tx,
                                            int ty, F &&f) {
    const int rows_in = 4 + g.kh - 1;
    const int n_taps = g.off + g.kw;
    for (int j = 0; j < rows_in; ++j) {
        const float *rowp = tile32 + (size_t)(ty * 4 + j) * g.sw + 4 * tx;
        for (int kb = 0; kb < n_taps; kb += 4) {
            const float4 q0 = *reinterpret_cast<const float4 *>(rowp + kb);
            const float4 q1 = *reinterpret_cast<const float4 *>(rowp + kb + 4);
            const float v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            // NaN test and f64 widening once per loaded cell, not once per (output, tap) use
            bool ok[8];
            double dv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                ok[i] = (v[i] == v[i]);
                dv[i] = (double)v[i];
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int ky = j - r;
                if (ky >= 0 && ky < g.kh) {
#pragma unroll
                    for (int tt = 0; tt < 4; ++tt) {
                        const int kx = kb + tt - g.off;
                        if (kx >= 0 && kx < g.kw && mask.m[ky * g.kw + kx]) {  // warp-uniform
#pragma unroll
                            for (int c = 0; c < 4; ++c) f(r, c, v[c + tt], dv[c + tt], ok[c + tt]);
                        }
                    }
                }
            }
        }
    }
}

template <int STAT>
__global__ void __launch_bounds__(256)
focal_stat_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ MaskBits mask,
                  float *__restrict__ out, int64_t out_pitch_elems, const TileGeom g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int nbox = (g.sh + g.box_h - 1) / g.box_h;
    const size_t tile_cells = (size_t)nbox * g.box_h * g.sw;
    float *tile32 = reinterpret_cast<float *>(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + tile_cells * sizeof(float));
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t n_tiles = (int64_t)g.tiles_x * g.tiles_y;
    uint32_t parity = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int tile_y = (int)(t / g.tiles_x), tile_x = (int)(t % g.tiles_x);
        const int x0 = tile_x * kTileW, y0 = tile_y * kTileH;
        load_tile_tma(&tmap, tile32, bar, g, x0, y0, parity);
        parity ^= 1u;
        float res[4][4];
        if constexpr (STAT == XRS_STAT_MEAN || STAT == XRS_STAT_VAR || STAT == XRS_STAT_STD) {
            double sum[4][4];
            int cnt[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) sum[r][c] = 0.0, cnt[r][c] = 0;
            focal_sweep(tile32, mask, g, tx, ty, [&](int r, int c, float, double d, bool ok) {
                sum[r][c] += ok ? d : 0.0;
                cnt[r][c] += ok ? 1 : 0;
            });
            if constexpr (STAT == XRS_STAT_MEAN) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) res[r][c] = (float)(sum[r][c] / (double)cnt[r][c]);
            } else {
                double ssd[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) sum[r][c] = sum[r][c] / (double)cnt[r][c], ssd[r][c] = 0.0;
                focal_sweep(tile32, mask, g, tx, ty, [&](int r, int c, float, double dvv, bool ok) {
                    const double d = dvv - sum[r][c];
                    ssd[r][c] += ok ? d * d : 0.0;
                });
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double var = ssd[r][c] / (double)cnt[r][c];
                        res[r][c] = (float)(STAT == XRS_STAT_VAR ? var : sqrt(var));
                    }
            }
        } else if constexpr (STAT == XRS_STAT_SUM) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) res[r][c] = 0.f;
            focal_sweep(tile32, mask, g, tx, ty, [&](int r, int c, float v, double, bool ok) { res[r][c] += ok ? v : 0.f; });
        } else {
            float mn[4][4], mx[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) mn[r][c] = mx[r][c] = nan_of<float>();
            focal_sweep(tile32, mask, g, tx, ty, [&](int r, int c, float v, double, bool ok) {
                if (ok) {
                    if (!(mn[r][c] < v)) mn[r][c] = v;
                    if (!(mx[r][c] > v)) mx[r][c] = v;
                }
            });
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    res[r][c] = STAT == XRS_STAT_MIN ? mn[r][c] : STAT == XRS_STAT_MAX ? mx[r][c] : mx[r][c] - mn[r][c];
        }
        const int64_t xo = (int64_t)x0 + 4 * tx;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int64_t yo = (int64_t)y0 + ty * 4 + r;
            if (yo < g.H && xo < g.W)  // W % 4 == 0 on this path
                __stcs(reinterpret_cast<float4 *>(out + yo * out_pitch_elems + xo),
                       make_float4(res[r][0], res[r][1], res[r][2], res[r][3]));
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
focal_stat_direct_kernel(const float *__restrict__ in, int64_t in_pitch_elems, const __grid_constant__ MaskBits mask,
                         float *__restrict__ out, int64_t out_pitch_elems, int64_t H, int64_t W, int kh, int kw,
                         int stat) {
    const int64_t n = H * W;
    const int ry = kh / 2, rx = kw / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t y = i / W, x = i % W;
        auto fetch = [=](int ky, int kx) {
            const int64_t yy = y + ky - ry, xx = x + kx - rx;
            return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? in[yy * in_pitch_elems + xx] : nan_of<float>();
        };
        out[y * out_pitch_elems + x] = focal_reduce(fetch, mask, kh, kw, stat);
    }
}

static int check_common(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H, int64_t W,
                        const double *kernel, int kh, int kw) {
    XRS_REQUIRE(in && out && kernel, "NULL pointer");
    XRS_REQUIRE((const void *)in != (const void *)out, "in and out must not alias");
    XRS_REQUIRE(kh >= 1 && kw >= 1 && (kh & 1) && (kw & 1), "kernel dimensions must be odd");
    XRS_REQUIRE(kh * kw <= kMaxTaps && kh <= 63 && kw <= 63, "kernel too large (max 49x49 taps, 63 per side)");
    XRS_REQUIRE(in_pitch % 4 == 0 && in_pitch >= W * 4 && out_pitch % 4 == 0 && out_pitch >= W * 4,
                "pitches must be multiples of 4 bytes and >= row bytes");
    XRS_REQUIRE(H < (1LL << 31) - 64 && W < (1LL << 31) - 256, "raster dimension too large");
    return XRS_OK;
}

static bool tile_geom(TileGeom &g, CUtensorMap *tmap, const float *in, int64_t in_pitch, float *out,
                      int64_t out_pitch, int64_t H, int64_t W, int kh, int kw, int tile_h) {
    g.H = H; g.W = W; g.kh = kh; g.kw = kw; g.ry = kh / 2; g.rx = kw / 2;
    g.pad = (g.rx + 3) / 4 * 4;
    g.off = g.pad - g.rx;
    // the 8-wide chunk loads of the convolution reach 4*31 + 4*((off+kw-1)/4) + 7 cells into a row
    g.sw = kTileW + ((g.off + kw + 3) / 4) * 4 + 4;
    g.sh = tile_h + kh - 1;
    g.tiles_x = (int)((W + kTileW - 1) / kTileW);
    g.tiles_y = (int)((H + tile_h - 1) / tile_h);
    g.box_h = g.sh <= 64 ? g.sh : 64;  // 64 % 8 == 0 keeps the following boxes 128-byte aligned
    if (g.sw > 256) return false;
    if (W % 4 != 0 || out_pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
    return make_tensor_map_2d(tmap, in, in_pitch, H, W, 4, g.sw, g.box_h);
}

}  // namespace xrs

using namespace xrs;

int xrs_conv3_strip(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H, int64_t W,
                    const double *kernel, cudaStream_t s);  // surface.cu

extern "C" {

int xrs_convolve2d_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H, int64_t W,
                       const double *kernel, int kh, int kw, xrs_stream_t s) {
    if (H <= 0 || W <= 0) return XRS_OK;
    const int rc = check_common(in, in_pitch, out, out_pitch, H, W, kernel, kh, kw);
    if (rc) return rc;
    if (kh == 3 && kw == 3) return xrs_conv3_strip(in, in_pitch, out, out_pitch, H, W, kernel, (cudaStream_t)s);
    static thread_local ConvWeights cw;
    for (int i = 0; i < kh * kw; ++i) cw.w[i] = kernel[i];
    TileGeom g;
    CUtensorMap tmap;
    const int sms = sm_count();
    if (tile_geom(g, &tmap, in, in_pitch, out, out_pitch, H, W, kh, kw, kConvTileH)) {
        const int nbox = (g.sh + g.box_h - 1) / g.box_h;
        const size_t cells = (size_t)nbox * g.box_h * g.sw;
        const size_t smem = cells * 4 + 16;
        if (smem <= 227 * 1024) {
            XRS_CUDA(cudaFuncSetAttribute(conv2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            XRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv2d_kernel, 256, smem));
            if (per_sm < 1) per_sm = 1;
            int64_t grid = (int64_t)sms * per_sm;
            const int64_t n_tiles = (int64_t)g.tiles_x * g.tiles_y;
            if (grid > n_tiles) grid = n_tiles;
            conv2d_kernel<<<(unsigned)grid, 256, smem, (cudaStream_t)s>>>(tmap, cw, out, out_pitch / 4, g);
            XRS_CUDA(cudaGetLastError());
            return XRS_OK;
        }
    }
    int64_t grid = (H * W + 255) / 256;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    conv2d_direct_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)s>>>(in, in_pitch / 4, cw, out, out_pitch / 4, H, W,
                                                                     kh, kw);
    XRS_CUDA(cudaGetLastError());
    return XRS_OK;
}

int xrs_focal_stat_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch, int64_t H, int64_t W,
                       const double *kernel, int kh, int kw, int stat, xrs_stream_t s) {
    if (H <= 0 || W <= 0) return XRS_OK;
    const int rc = check_common(in, in_pitch, out, out_pitch, H, W, kernel, kh, kw);
    if (rc) return rc;
    XRS_REQUIRE(stat >= XRS_STAT_MEAN && stat <= XRS_STAT_VAR, "unknown focal statistic");
    static thread_local MaskBits mask;
    for (int i = 0; i < kh * kw; ++i) mask.m[i] = (kernel[i] == 1.0) ? 1 : 0;  // focal.py:323
    TileGeom g;
    CUtensorMap tmap;
    const int sms = sm_count();
    if (tile_geom(g, &tmap, in, in_pitch, out, out_pitch, H, W, kh, kw, kTileH)) {
        const int nbox = (g.sh + g.box_h - 1) / g.box_h;
        const size_t smem = (size_t)nbox * g.box_h * g.sw * 4 + 16;
        if (smem <= 227 * 1024) {
            int per_sm = (int)((227 * 1024) / (smem + 1024));
            if (per_sm > 3) per_sm = 3;
            if (per_sm < 1) per_sm = 1;
            int64_t grid = (int64_t)sms * per_sm;
            const int64_t n_tiles = (int64_t)g.tiles_x * g.tiles_y;
            if (grid > n_tiles) grid = n_tiles;
#define XRS_FS(ST)                                                                                              \
    case ST:                                                                                                    \
        XRS_CUDA(cudaFuncSetAttribute(focal_stat_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                      (int)smem));                                                              \
        focal_stat_kernel<ST><<<(unsigned)grid, 256, smem, (cudaStream_t)s>>>(tmap, mask, out, out_pitch / 4, g); \
        break;
            switch (stat) {
                XRS_FS(XRS_STAT_MEAN) XRS_FS(XRS_STAT_SUM) XRS_FS(XRS_STAT_MIN) XRS_FS(XRS_STAT_MAX)
                XRS_FS(XRS_STAT_STD) XRS_FS(XRS_STAT_RANGE) XRS_FS(XRS_STAT_VAR)
            }
#undef XRS_FS
            XRS_CUDA(cudaGetLastError());
            return XRS_OK;
        }
    }
    int64_t grid = (H * W + 255) / 256;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    focal_stat_direct_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)s>>>(in, in_pitch / 4, mask, out, out_pitch / 4,
                                                                         H, W, kh, kw, stat);
    XRS_CUDA(cudaGetLastError());
    return XRS_OK;
}

}  // extern "C"
