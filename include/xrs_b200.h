/*
 * xrs_b200.h -- C ABI of libxrs_b200.so, the B200 (sm_100a) backend for the dense 2-D
 * stencil hot path of xarray-spatial.
 *
 * The reference has no FFI: its backend seam is the set of Python runner callables that
 * xrspatial.utils.ArrayTypeFunctionMapping (utils.py:117-143) selects by array type, e.g.
 * slope._run_cupy(data, cellsize_x, cellsize_y) (slope.py:145).  Each entry point below is
 * what a fifth, "b200" runner of that mapping binds through ctypes; the comment on every
 * function names the reference runner it replaces (file:line in /root/reference/xrspatial).
 * INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - plain C types only; rasters are row-major (H rows = y, W cols = x); pitches in BYTES.
 *  - `*_f32` device entry points take DEVICE pointers and a cudaStream_t (as void*), only
 *    enqueue work on that stream and never synchronise or allocate user-visible memory.
 *  - `xrs_host_*` entry points take HOST pointers (pinned or pageable), run the same kernels
 *    through an internal pipelined H2D / compute / D2H stripe engine, and return when the
 *    result is in `out`.  They are what a numpy-backed DataArray call uses.
 *  - return value: 0 on success, negative xrs_status otherwise; xrs_last_error_string()
 *    describes the last failure on the calling thread.  Nothing throws across the ABI.
 *  - inputs are never modified.
 */
#ifndef XRS_B200_H
#define XRS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XRS_ABI_VERSION 1

typedef void *xrs_stream_t; /* cudaStream_t */

enum xrs_status {
    XRS_OK = 0,
    XRS_EINVAL = -1,       /* bad argument (shape, null pointer, even kernel, ...) */
    XRS_ECUDA = -2,        /* CUDA runtime / driver error */
    XRS_EUNSUPPORTED = -3, /* valid request this build cannot serve */
    XRS_ENOMEM = -4
};

/* focal statistic ids (focal.py:782-790 `_function_mapping`) */
enum xrs_focal_stat {
    XRS_STAT_MEAN = 0, XRS_STAT_SUM = 1, XRS_STAT_MIN = 2, XRS_STAT_MAX = 3,
    XRS_STAT_STD = 4, XRS_STAT_RANGE = 5, XRS_STAT_VAR = 6
};

/* element types for zonal inputs */
enum xrs_dtype { XRS_F32 = 0, XRS_F64 = 1, XRS_I32 = 2, XRS_I64 = 3, XRS_I16 = 4, XRS_U16 = 5 };

int xrs_abi_version(void);
const char *xrs_last_error_string(void);
/* device 0..n-1 properties the Python layer needs to size launches / report rooflines */
int xrs_device_count(int *n);
int xrs_device_sm_count(int device, int *sm_count);

/* ------------------------------------------------------------------ surface (3x3)
 * All: in/out float32, 1-cell NaN ring, NaN inputs propagate to their 3x3 neighbourhood. */

/* slope._run_cupy (slope.py:145-160) / `_cpu` (slope.py:56-76): Horn slope in degrees. */
int xrs_slope_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch,
                  int64_t H, int64_t W, double cellsize_x, double cellsize_y, xrs_stream_t s);
/* aspect._run_cupy (aspect.py:139-147) / `_run_numpy` (aspect.py:56-90): compass degrees,
 * -1 on flats; follows the CPU path (no 359.999 clamp). */
int xrs_aspect_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch,
                   int64_t H, int64_t W, xrs_stream_t s);
/* curvature._run_cupy (curvature.py:81-95) / `_cpu` (curvature.py:31-41). */
int xrs_curvature_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch,
                      int64_t H, int64_t W, double cellsize, xrs_stream_t s);
/* hillshade._run_cupy (hillshade.py:78-100) / `_run_numpy` (hillshade.py:20-35);
 * output float32 as the reference's GPU path and docs promise. */
int xrs_hillshade_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch,
                      int64_t H, int64_t W, double azimuth, double angle_altitude,
                      xrs_stream_t s);
/* analytics.summarize_terrain (analytics.py:84-86) fused: one read, up to four writes.
 * Any output pointer may be NULL (that product is skipped); all outputs share out_pitch. */
int xrs_surface_suite_f32(const float *in, int64_t in_pitch, float *slope_out,
                          float *aspect_out, float *curvature_out, float *hillshade_out,
                          int64_t out_pitch, int64_t H, int64_t W, double cellsize_x,
                          double cellsize_y, double azimuth, double angle_altitude,
                          xrs_stream_t s);

/* slope / aspect with method='geodesic' (slope._run_cupy_geodesic slope.py:176-195,
 * aspect._run_cupy_geodesic aspect.py:179-198; arithmetic of geodesic.py:40-231): elevation float32
 * or float64 (elev_dtype XRS_F32 / XRS_F64), latitude / longitude in degrees as DEVICE float64
 * arrays -- lat[H], lon[W] for a regular grid (latlon_2d = 0) or (H, W) contiguous arrays for a
 * curvilinear one (latlon_2d = 1); z_factor converts the elevation unit to metres.  float32
 * output: slope in degrees, or compass aspect (-1 on flats) when want_aspect != 0. */
int xrs_geodesic(const void *elev, int elev_dtype, int64_t elev_pitch, const double *lat,
                 const double *lon, int latlon_2d, float *out, int64_t out_pitch, int64_t H, int64_t W,
                 double z_factor, int want_aspect, xrs_stream_t s);

/* Direct ingestion: slope / aspect / curvature / hillshade (op = XRS_OP_*) on a
 * raster of int16, uint16, int32 or float64 cells (in_dtype = XRS_I16 / XRS_U16 / XRS_I32 / XRS_F64),
 * converted to float32 in registers exactly like the reference's `.astype(np.float32)`
 * (slope.py:58,150) but without the extra pass.  p: slope {csx, csy}; curvature {cellsize};
 * hillshade {azimuth, altitude}.  Needs 16-byte aligned rows and W % 4 == 0, else
 * XRS_EUNSUPPORTED (cast and use the float32 entry points). */
int xrs_surface_typed(int op, const void *in, int in_dtype, int64_t in_pitch, float *out,
                      int64_t out_pitch, int64_t H, int64_t W, const double *p, xrs_stream_t s);

/* ------------------------------------------------------------------ focal / convolution */
/* focal._mean_cupy (focal.py:135-146) / `_mean_numpy` (focal.py:44-67): ONE pass of the 3x3
 * NaN-skipping mean with clamped windows; centre cells equal (NaN-aware) to one of
 * `excludes` (host array, n_ex <= 8) are copied through.  f32: float in/out (sums in f64);
 * f64: double in/out.  in and out must not alias. */
int xrs_focal_mean_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch,
                       int64_t H, int64_t W, const double *excludes, int n_ex, xrs_stream_t s);
int xrs_focal_mean_f64(const double *in, int64_t in_pitch, double *out, int64_t out_pitch,
                       int64_t H, int64_t W, const double *excludes, int n_ex, xrs_stream_t s);
/* float32 in, float64 out: what focal.mean's `agg.data.astype(float)` (focal.py:257) yields for
 * a float32 raster, without a separate widening pass */
int xrs_focal_mean_f32_f64(const float *in, int64_t in_pitch, double *out, int64_t out_pitch,
                           int64_t H, int64_t W, const double *excludes, int n_ex,
                           xrs_stream_t s);
/* convolution._convolve_2d_cupy (convolution.py:368-374) / `_convolve_2d_numpy` (:285-313):
 * correlation with a host float64 kernel (kh, kw odd, <= 63); NaN ring of (kh/2, kw/2). */
int xrs_convolve2d_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch,
                       int64_t H, int64_t W, const double *kernel, int kh, int kw,
                       xrs_stream_t s);
/* focal._focal_stats_cupy (focal.py:757-779) with the CPU semantics of `_apply_numpy`
 * (focal.py:305-326) + reducers (:268-302): cells where kernel == 1 participate, NaN and
 * out-of-raster cells are skipped.  `stat` is an xrs_focal_stat.  XRS_STAT_MEAN over a kernel of all
 * ones (np.ones((kh, kw)), the reference's FocalApply benchmark; odd kh, kw <= 25) runs on the running-box
 * kernel in NaN-skipping mode, O(1) work per cell; everything else on the tiled kernels. */
int xrs_focal_stat_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch,
                       int64_t H, int64_t W, const double *kernel, int kh, int kw, int stat,
                       xrs_stream_t s);
/* focal.focal_stats (focal.py:800-878: seven `apply` calls stacked by xr.concat) in one pass:
 * plane i of `out` (planes `plane_stride` bytes apart, rows `out_pitch` bytes apart) receives
 * statistic stats[i] (xrs_focal_stat ids, no duplicates, n_stats <= 7).  The tile is loaded
 * once and swept twice for all seven statistics; results are bit-identical to n_stats calls
 * of xrs_focal_stat_f32. */
int xrs_focal_stats_multi_f32(const float *in, int64_t in_pitch, float *out, int64_t out_pitch,
                              int64_t plane_stride, int64_t H, int64_t W, const double *kernel,
                              int kh, int kw, const int *stats, int n_stats, xrs_stream_t s);

/* focal.hotspots (focal.py:1050-1125) = convolve_2d with kernel / kernel.sum(), then
 * z = (mean - global_mean) / global_std classified into {0, +-90, +-95, +-99} (int8).
 * xrs_global_stats_f32: partial3[0..2] (device) <- count, sum(v - pivot), sum((v - pivot)^2) over
 * the non-NaN cells, from which np.nanmean / np.nanstd follow.  (focal.py:881-937) */
int xrs_global_stats_f32(const float *values, int64_t n, double pivot, double *partial3,
                         xrs_stream_t s);
int xrs_hotspots_classify_f32(const float *mean, int64_t n, double global_mean, double global_std,
                              int8_t *out, xrs_stream_t s);

/* ------------------------------------------------------------------ multispectral
 * Elementwise over n contiguous float32 cells; out is NaN where the denominator is 0. */
/* multispectral._run_normalized_ratio_cupy (:862) -- ndvi, nbr, nbr2, ndmi */
int xrs_normalized_ratio_f32(const float *a, const float *b, float *out, int64_t n, xrs_stream_t s);
/* multispectral._savi_cupy (:912) */
int xrs_savi_f32(const float *nir, const float *red, double soil_factor, float *out, int64_t n,
                 xrs_stream_t s);
/* multispectral._evi_cupy (:210) */
int xrs_evi_f32(const float *nir, const float *red, const float *blue, double c1, double c2,
                double soil_factor, double gain, float *out, int64_t n, xrs_stream_t s);
/* multispectral._arvi_cupy (:65) */
int xrs_arvi_f32(const float *nir, const float *red, const float *blue, float *out, int64_t n,
                 xrs_stream_t s);
/* multispectral._gci_cupy (:378) */
int xrs_gci_f32(const float *nir, const float *green, float *out, int64_t n, xrs_stream_t s);
/* multispectral._sipi_cupy (:1052) */
int xrs_sipi_f32(const float *nir, const float *red, const float *blue, float *out, int64_t n,
                 xrs_stream_t s);
/* multispectral._ebbi_cupy (:1195) */
int xrs_ebbi_f32(const float *red, const float *swir, const float *tir, float *out, int64_t n,
                 xrs_stream_t s);

/* ------------------------------------------------------------------ zonal.stats
 * zonal._stats_cupy (zonal.py:335-419) / `_stats_numpy` (:280-332) replaced by one streaming group-by that
 * discovers the zone ids and accumulates per-zone partials; mean / std / var are finalised from them by the
 * caller, and partials from several devices combine with sum / min / max (NCCL AllReduce).  (Round 1 also
 * exported a lookup-table kernel for a given id list, xrs_zonal_partials(_ex): 0.24 of the HBM roofline,
 * slower than the group-by it specialised; removed -- the float64 second pass now runs on the group-by
 * kernel, xrs_zonal_hash_second_pass.)
 *
 * zones (zones_dtype: I32, I64, F32, F64) and values (values_dtype: F32, F64) are device arrays of n
 * cells; a cell contributes to its zone if its value is finite and != nodata (when has_nodata). */

/* The group-by: aggregation into an open-addressing
 * hash table of `cap` slots (power of two >= 1024; device arrays keys/count/s1/s2/vmin/vmax of
 * `cap` entries, initialised by xrs_zonal_hash_init).  keys[slot] is the zone id (int64 for
 * integer zones, the bit pattern of the float64 value for float zones; INT64_MIN = empty),
 * s1/s2 are sums of (v - pivot) and (v - pivot)^2.  Every finite zone value present in the
 * raster gets a slot, also when none of its cells is valid (count 0).  *overflow (device
 * int) is set when the raster holds more than `cap` distinct zones.  row_len is the raster's
 * row length (n = rows * row_len): the scan walks down 128-column strips so that runs of equal
 * zone ids stay long. */
int xrs_zonal_hash_init(int64_t *keys, int64_t *count, double *s1, double *s2, double *vmin,
                        double *vmax, int cap, int *overflow, xrs_stream_t s);
int xrs_zonal_hash_accumulate(const void *values, int values_dtype, const void *zones,
                              int zones_dtype, int64_t n, int64_t row_len, double pivot, int has_nodata,
                              double nodata, int64_t *keys, int64_t *count, double *s1,
                              double *s2, double *vmin, double *vmax, int cap, int *overflow,
                              xrs_stream_t s);

/* The whole single-pass group-by behind one call, with no host round trip in the middle: samples the
 * pivot on the device (unless use_pivot_hint: row stripes must share one pivot), initialises the table,
 * accumulates, then compacts the used slots into `packed` (device, 3 + 6 * max_out doubles):
 *   packed[0] = used slots, packed[1] = table overflowed (grow `cap` and retry), packed[2] = pivot,
 *   then 6 rows of max_out: key bit patterns, count (int64 bit patterns), s1, s2, min, max, in any
 *   order.  Slots beyond max_out are dropped (packed[0] > max_out tells; read the table instead).
 * flags: 2 device ints of scratch.  Replaces the sort + per-zone reductions of zonal.py:280-332. */
int xrs_zonal_hash_run(const void *values, int values_dtype, const void *zones, int zones_dtype, int64_t n,
                       int64_t row_len, int has_nodata, double nodata, int use_pivot_hint, double pivot_hint,
                       int64_t *keys, int64_t *count, double *s1, double *s2, double *vmin, double *vmax, int cap,
                       double *packed, int max_out, int *flags, xrs_stream_t s);

/* Second pass for float64 rasters (numpy's two-pass variance, zonal.py:75-76 `ndarray.std / var`): the same
 * streaming group-by over the table xrs_zonal_hash_run left behind -- `keys` as populated by the first pass,
 * read only -- with sums taken about `zone_pivots[slot]` (device, `cap` doubles: the zone's mean from the
 * first pass) instead of one global pivot, so that s2 / n - (s1 / n)^2 does not cancel.  count / s1 / s2 /
 * vmin / vmax: a second set of `cap`-entry accumulators (reset here); `packed` / `flags` as above
 * (packed[2] = 0).  values_dtype must be XRS_F64.  */
int xrs_zonal_hash_second_pass(const void *values, int values_dtype, const void *zones, int zones_dtype, int64_t n,
                               int64_t row_len, int has_nodata, double nodata, const int64_t *keys,
                               const double *zone_pivots, int64_t *count, double *s1, double *s2, double *vmin,
                               double *vmax, int cap, double *packed, int max_out, int *flags, xrs_stream_t s);

/* `majority` (zonal.py:56-68): counts (zone, value) pairs of float32 values / int32 zones into
 * a hash table (keys/count of `cap` entries, initialised with xrs_zonal_hash_init; key =
 * (zone << 32) | float32 bits of the value, -0.0 folded into +0.0).  The caller picks the most
 * frequent value per zone (smallest value on ties). */
int xrs_zonal_pair_count(const float *values, const int32_t *zones, int64_t n, int64_t row_len,
                         int has_nodata, double nodata, int64_t *keys, int64_t *count, int cap,
                         int *overflow, xrs_stream_t s);

/* ------------------------------------------------------------------ host-buffer (end-to-end)
 * Same operators on HOST rasters: the library stripes the raster over rows, and overlaps
 * host->device copies, kernels and device->host copies on internal streams.  `op` selects
 * the operator; scalar parameters are passed in p[0..7] in the order of the device entry
 * point (e.g. slope: p[0]=cellsize_x, p[1]=cellsize_y).  aux/naux carry the excludes
 * (focal mean) or the kernel followed by kh,kw in p[0],p[1] (convolve / focal stat). */
enum xrs_op {
    XRS_OP_SLOPE = 0, XRS_OP_ASPECT = 1, XRS_OP_CURVATURE = 2, XRS_OP_HILLSHADE = 3,
    XRS_OP_FOCAL_MEAN = 4, XRS_OP_CONVOLVE = 5, XRS_OP_FOCAL_STAT = 6,
    XRS_OP_FOCAL_MEAN_F64 = 7,     /* in/out are double */
    XRS_OP_FOCAL_MEAN_F32_F64 = 8  /* in float, out double */
};
/* in/out: float32 rasters (see the two FOCAL_MEAN_F* ops for float64), contiguous rows.
 * p: slope {csx, csy}; curvature {cellsize}; hillshade {azimuth, altitude};
 *    convolve {kh, kw}; focal stat {kh, kw, stat}.  aux: excludes / kernel (host). */
int xrs_host_stencil(int op, const void *in, void *out, int64_t H, int64_t W, const double *p,
                     const double *aux, int naux, int device);
/* slope / aspect / curvature / hillshade on a HOST raster of int16 / uint16 / int32 / float64 cells
 * (xrs_surface_typed behind the same pipeline; needs W % 4 == 0): the raw cells cross PCIe. */
int xrs_host_surface_typed(int op, const void *in, int in_dtype, float *out, int64_t H, int64_t W,
                           const double *p, int device);
/* The same two entry points over SEVERAL devices: output rows are cut into one stripe per listed
 * device, each stripe runs the chunk pipeline on its own host thread and PCIe link; the halo rows
 * of a stripe come from the host raster, so there is no device-to-device traffic.  This is the
 * reference's dask.map_overlap(depth=r, boundary=nan) over row blocks (slope.py:94-97,
 * focal.py:72-75) for host rasters.  devices: CUDA ordinals, each at most once. */
int xrs_host_stencil_multi(int op, const void *in, void *out, int64_t H, int64_t W, const double *p,
                           const double *aux, int naux, const int *devices, int n_devices);
int xrs_host_surface_typed_multi(int op, const void *in, int in_dtype, float *out, int64_t H, int64_t W,
                                 const double *p, const int *devices, int n_devices);
/* frees the per-device staging buffers the host path keeps between calls */
int xrs_host_release(int device);
/* pinned host memory helpers (cudaHostAlloc / cudaFreeHost) for callers that want
 * full-speed DMA on the host path */
int xrs_host_alloc(void **ptr, int64_t bytes);
int xrs_host_free(void *ptr);

/* test / profiling hooks: which kernel the last launch on this thread chose -- 0 cp.async strip
 * kernel, 1 TMA strip kernel, 2 direct-ingest TMA kernel, 3 running-box kernel (uniform convolve_2d, focal.apply
 * mean over an all-ones window), 4 generic tiled convolve, 5 bounds-checked convolve fallback, 6 fused focal
 * statistics -- and with how many CTAs */
int xrs_debug_last_used_tma(void);
int xrs_debug_last_grid(void);
/* host-only test hook: the row-segment height the persistent kernels pick for a raster of H rows cut into
 * n_tiles column tiles, dealt round-robin to `resident` CTAs (at least min_rows, rows + lead a multiple of
 * quantum, about `want` tasks per CTA): minimises ceil(tasks / resident) x (rows + lead). */
int64_t xrs_debug_pick_seg_rows(int64_t H, int64_t n_tiles, int64_t resident, int64_t min_rows, int64_t lead,
                                int64_t quantum, int64_t want);

/* ------------------------------------------------------------------ synthetic inputs
 * Deterministic fBm-like terrain (value-noise octaves), a pure function of
 * (seed, global row, global col): stripes generated on different devices tile exactly.
 * Used by bench.py and the tests; not part of the reference API. */
int xrs_synth_terrain_f32(float *out, int64_t out_pitch, int64_t H, int64_t W,
                          int64_t row0, int64_t col0, uint64_t seed, float zmin, float zmax,
                          xrs_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* XRS_B200_H */
