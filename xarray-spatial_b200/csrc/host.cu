// host.cu -- host-buffer (end-to-end) entry point: the same operators on HOST rasters.
//
// This is what a numpy-backed DataArray call binds (the reference's `_run_numpy` slot,
// e.g. slope.py:79).  The raster is cut into row chunks of ~32 MiB; chunk c is copied
// host->device on a copy stream while chunk c-1 is computed and chunk c-2 travels back,
// using three device slots and events, so PCIe runs full duplex and the kernels hide behind
// it.  Each chunk carries `r` halo rows above and below (r = kernel radius); the operator
// treats the chunk view as a raster of its own, so its first/last r output rows are only
// correct when they coincide with the real raster edge -- interior halo rows are simply
// not copied back.  Host buffers may be pageable (works, slower) or pinned (xrs_host_alloc).
//
// Several GPUs (xrs_host_stencil_multi): the output rows are cut into one stripe per device and
// every device runs the pipeline above on its stripe from its own host thread -- one PCIe link
// each, no device-to-device traffic at all, because a stripe's halo rows come straight from the
// host raster like any chunk's.  This is the reference's `dask.map_overlap(depth=r)` over row
// blocks (slope.py:94-97) with the blocks being PCIe-sized.
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace xrs {

struct Slot {
    void *din = nullptr, *dout = nullptr;
    size_t cap_in = 0, cap_out = 0;
    cudaEvent_t in_done = nullptr, k_done = nullptr, out_done = nullptr;
};
struct HostCtx {
    bool init = false;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    Slot slot[3];
    std::mutex mu;
};
static HostCtx g_ctx[16];

static int ensure_ctx(HostCtx &c) {
    if (c.init) return XRS_OK;
    XRS_CUDA(cudaStreamCreateWithFlags(&c.s_in, cudaStreamNonBlocking));
    XRS_CUDA(cudaStreamCreateWithFlags(&c.s_k, cudaStreamNonBlocking));
    XRS_CUDA(cudaStreamCreateWithFlags(&c.s_out, cudaStreamNonBlocking));
    for (auto &s : c.slot) {
        XRS_CUDA(cudaEventCreateWithFlags(&s.in_done, cudaEventDisableTiming));
        XRS_CUDA(cudaEventCreateWithFlags(&s.k_done, cudaEventDisableTiming));
        XRS_CUDA(cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming));
    }
    c.init = true;
    return XRS_OK;
}
static int ensure_cap(void **p, size_t *cap, size_t need) {
    if (*cap >= need) return XRS_OK;
    if (*p) XRS_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    XRS_CUDA(cudaMalloc(p, need));
    *cap = need;
    return XRS_OK;
}

static int run_op(int op, int in_dtype, const void *din, void *dout, int64_t pitch, int64_t opitch, int64_t h,
                  int64_t W, const double *p, const double *aux, int naux, cudaStream_t s) {
    if (in_dtype != XRS_F32)  // raw int16 / uint16 / int32 / float64 cells: direct-ingest kernels
        return xrs_surface_typed(op, din, in_dtype, pitch, (float *)dout, opitch, h, W, p, s);
    const float *fi = (const float *)din;
    float *fo = (float *)dout;
    switch (op) {
        case XRS_OP_SLOPE: return xrs_slope_f32(fi, pitch, fo, opitch, h, W, p[0], p[1], s);
        case XRS_OP_ASPECT: return xrs_aspect_f32(fi, pitch, fo, opitch, h, W, s);
        case XRS_OP_CURVATURE: return xrs_curvature_f32(fi, pitch, fo, opitch, h, W, p[0], s);
        case XRS_OP_HILLSHADE: return xrs_hillshade_f32(fi, pitch, fo, opitch, h, W, p[0], p[1], s);
        case XRS_OP_FOCAL_MEAN: return xrs_focal_mean_f32(fi, pitch, fo, opitch, h, W, aux, naux, s);
        case XRS_OP_FOCAL_MEAN_F64:
            return xrs_focal_mean_f64((const double *)din, pitch, (double *)dout, opitch, h, W, aux, naux, s);
        case XRS_OP_FOCAL_MEAN_F32_F64:
            return xrs_focal_mean_f32_f64(fi, pitch, (double *)dout, opitch, h, W, aux, naux, s);
        case XRS_OP_CONVOLVE: return xrs_convolve2d_f32(fi, pitch, fo, opitch, h, W, aux, (int)p[0], (int)p[1], s);
        case XRS_OP_FOCAL_STAT:
            return xrs_focal_stat_f32(fi, pitch, fo, opitch, h, W, aux, (int)p[0], (int)p[1], (int)p[2], s);
    }
    set_error("unknown op %d", op);
    return XRS_EINVAL;
}

}  // namespace xrs

using namespace xrs;

// Output rows [y_begin, y_end) of the H x W raster on `device`.
static int host_pipeline(int op, int in_dtype, const void *in, void *out, int64_t H, int64_t W, const double *p,
                         const double *aux, int naux, int device, int64_t y_begin = 0, int64_t y_end = -1) {
    if (y_end < 0) y_end = H;
    if (H <= 0 || W <= 0 || y_end <= y_begin) return XRS_OK;
    XRS_REQUIRE(in && out, "NULL host pointer");
    XRS_REQUIRE(device >= 0 && device < 16, "device index out of range");
    XRS_REQUIRE(op >= XRS_OP_SLOPE && op <= XRS_OP_FOCAL_MEAN_F32_F64, "unknown op");
    int esz = (op == XRS_OP_FOCAL_MEAN_F64) ? 8 : 4;                                             // input
    if (in_dtype != XRS_F32) {
        XRS_REQUIRE(op <= XRS_OP_HILLSHADE, "typed input is served for slope, aspect, curvature, hillshade");
        XRS_REQUIRE(W % 4 == 0, "typed host input needs W % 4 == 0");
        esz = (in_dtype == XRS_F64) ? 8 : (in_dtype == XRS_I32 ? 4 : 2);
    }
    const int osz = (op == XRS_OP_FOCAL_MEAN_F64 || op == XRS_OP_FOCAL_MEAN_F32_F64) ? 8 : 4;  // output
    int radius = 1;
    if (op == XRS_OP_CONVOLVE || op == XRS_OP_FOCAL_STAT) {
        XRS_REQUIRE(p && aux, "kernel parameters missing");
        radius = (int)p[0] / 2;
    }
    if ((op == XRS_OP_SLOPE || op == XRS_OP_CURVATURE || op == XRS_OP_HILLSHADE) && !p) {
        set_error("scalar parameters missing");
        return XRS_EINVAL;
    }
    int prev = 0;
    XRS_CUDA(cudaGetDevice(&prev));
    XRS_CUDA(cudaSetDevice(device));
    HostCtx &c = g_ctx[device];
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = ensure_ctx(c);
    if (rc) { cudaSetDevice(prev); return rc; }

    // device row pitch: multiple of 16 bytes so the TMA kernels apply whenever W % 4 == 0
    const int64_t row_bytes = W * esz, orow_bytes = W * osz;
    const int64_t pitch = (row_bytes + 15) / 16 * 16, opitch = (orow_bytes + 15) / 16 * 16;
    const int64_t span = y_end - y_begin;
    int64_t rows = (32LL << 20) / pitch;
    if (rows < 8 * radius + 8) rows = 8 * radius + 8;
    if (rows > span) rows = span;
    const int64_t n_chunks = (span + rows - 1) / rows;
    const size_t cap = (size_t)(rows + 2 * radius) * pitch, ocap = (size_t)(rows + 2 * radius) * opitch;

    for (int64_t ci = 0; ci < n_chunks && rc == XRS_OK; ++ci) {
        Slot &s = c.slot[ci % 3];
        const int64_t r0 = y_begin + ci * rows, r1 = (r0 + rows < y_end) ? r0 + rows : y_end;
        const int64_t a0 = (r0 - radius > 0) ? r0 - radius : 0, a1 = (r1 + radius < H) ? r1 + radius : H;
        const int64_t h = a1 - a0;
        // the slot is free once its previous D2H has finished
        cudaError_t e = cudaEventSynchronize(s.out_done);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaEventSynchronize"); break; }
        rc = ensure_cap(&s.din, &s.cap_in, cap);
        if (rc) break;
        rc = ensure_cap(&s.dout, &s.cap_out, ocap);
        if (rc) break;
        e = cudaMemcpy2DAsync(s.din, pitch, (const char *)in + a0 * row_bytes, row_bytes, row_bytes, h,
                              cudaMemcpyHostToDevice, c.s_in);
        if (e == cudaSuccess) e = cudaEventRecord(s.in_done, c.s_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c.s_k, s.in_done, 0);
        if (e != cudaSuccess) { rc = cuda_fail(e, "H2D enqueue"); break; }
        rc = run_op(op, in_dtype, s.din, s.dout, pitch, opitch, h, W, p, aux, naux, c.s_k);
        if (rc) break;
        e = cudaEventRecord(s.k_done, c.s_k);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c.s_out, s.k_done, 0);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync((char *)out + r0 * orow_bytes, orow_bytes,
                                  (const char *)s.dout + (r0 - a0) * opitch, opitch, orow_bytes, r1 - r0,
                                  cudaMemcpyDeviceToHost, c.s_out);
        if (e == cudaSuccess) e = cudaEventRecord(s.out_done, c.s_out);
        if (e != cudaSuccess) { rc = cuda_fail(e, "D2H enqueue"); break; }
    }
    cudaError_t e = cudaStreamSynchronize(c.s_out);
    cudaError_t e2 = cudaStreamSynchronize(c.s_k);
    cudaError_t e3 = cudaStreamSynchronize(c.s_in);
    if (rc == XRS_OK && (e != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess))
        rc = cuda_fail(e != cudaSuccess ? e : (e2 != cudaSuccess ? e2 : e3), "pipeline synchronize");
    cudaSetDevice(prev);
    return rc;
}

extern "C" int xrs_host_stencil(int op, const void *in, void *out, int64_t H, int64_t W, const double *p,
                                const double *aux, int naux, int device) {
    return host_pipeline(op, XRS_F32, in, out, H, W, p, aux, naux, device);
}

// Row stripes over several devices, one host thread and one PCIe link per device.  Errors: the
// first failing stripe's status and message are returned (the message is copied out of the worker
// thread, whose thread-local error string the caller cannot see).
static int host_multi(int op, int in_dtype, const void *in, void *out, int64_t H, int64_t W, const double *p,
                      const double *aux, int naux, const int *devices, int n_devices) {
    XRS_REQUIRE(devices != nullptr && n_devices >= 1 && n_devices <= 16, "1 .. 16 devices expected");
    if (H <= 0 || W <= 0) return XRS_OK;
    int radius = 1;
    if (op == XRS_OP_CONVOLVE || op == XRS_OP_FOCAL_STAT) {
        XRS_REQUIRE(p != nullptr, "kernel parameters missing");
        radius = (int)p[0] / 2;
    }
    // stripes shorter than a few halos are not worth a device
    int n = n_devices;
    while (n > 1 && H / n < 8 * radius + 8) --n;
    if (n == 1) return host_pipeline(op, in_dtype, in, out, H, W, p, aux, naux, devices[0]);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) XRS_REQUIRE(devices[i] != devices[j], "device listed twice");
    std::vector<int> rc(n, XRS_OK);
    std::vector<std::string> msg(n);
    std::vector<std::thread> th;
    const int64_t base = H / n, rem = H % n;
    int64_t y = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t h = base + (i < rem ? 1 : 0);
        const int64_t y0 = y, y1 = y + h;
        y = y1;
        th.emplace_back([&, i, y0, y1] {
            rc[i] = host_pipeline(op, in_dtype, in, out, H, W, p, aux, naux, devices[i], y0, y1);
            if (rc[i] != XRS_OK) msg[i] = xrs_last_error_string();
        });
    }
    for (auto &t : th) t.join();
    for (int i = 0; i < n; ++i)
        if (rc[i] != XRS_OK) {
            set_error("device %d: %s", devices[i], msg[i].c_str());
            return rc[i];
        }
    return XRS_OK;
}

extern "C" int xrs_host_stencil_multi(int op, const void *in, void *out, int64_t H, int64_t W, const double *p,
                                      const double *aux, int naux, const int *devices, int n_devices) {
    return host_multi(op, XRS_F32, in, out, H, W, p, aux, naux, devices, n_devices);
}

extern "C" int xrs_host_surface_typed_multi(int op, const void *in, int in_dtype, float *out, int64_t H, int64_t W,
                                            const double *p, const int *devices, int n_devices) {
    XRS_REQUIRE(in_dtype == XRS_F64 || in_dtype == XRS_I32 || in_dtype == XRS_I16 || in_dtype == XRS_U16,
                "in_dtype must be int16, uint16, int32 or float64");
    return host_multi(op, in_dtype, in, out, H, W, p, nullptr, 0, devices, n_devices);
}

// slope / aspect / curvature / hillshade on a HOST raster of int16 / uint16 / int32 / float64 cells:
// the raw cells travel over PCIe (half the bytes for 16-bit DEMs) and are converted on the device.
extern "C" int xrs_host_surface_typed(int op, const void *in, int in_dtype, float *out, int64_t H, int64_t W,
                                      const double *p, int device) {
    XRS_REQUIRE(in_dtype == XRS_F64 || in_dtype == XRS_I32 || in_dtype == XRS_I16 || in_dtype == XRS_U16,
                "in_dtype must be int16, uint16, int32 or float64");
    return host_pipeline(op, in_dtype, in, out, H, W, p, nullptr, 0, device);
}

// release the per-device staging buffers (tests / interpreter shutdown)
extern "C" int xrs_host_release(int device) {
    XRS_REQUIRE(device >= 0 && device < 16, "device index out of range");
    HostCtx &c = g_ctx[device];
    std::lock_guard<std::mutex> lock(c.mu);
    if (!c.init) return XRS_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    for (auto &s : c.slot) {
        if (s.din) cudaFree(s.din);
        if (s.dout) cudaFree(s.dout);
        s.din = s.dout = nullptr;
        s.cap_in = s.cap_out = 0;
    }
    cudaSetDevice(prev);
    return XRS_OK;
}
