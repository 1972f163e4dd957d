#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 raster backend.

Metric (BASELINE.json): Mcells/s of slope + hillshade + focal.mean on a float32 DEM, with the
fraction of the HBM roofline, at 1/2/4/8 GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--raster R]

N = 1   : 32768 x 32768 synthetic fBm-like DEM resident on the GPU (BASELINE configs[1]).
N > 1   : launched by torchrun, one rank per GPU; a 65536 x 65536 DEM (configs[4]) is row-striped
          over the ranks, every step exchanges the 1-row halos with NCCL send/recv and runs the
          three operators on the stripe (strong scaling: the raster is fixed, Mcells/s is
          size-normalised so it compares directly with the N = 1 line).
A step = slope, hillshade and focal.mean (one pass) each once over the whole raster:
3 * H * W cells.  Inputs are 4-16 GiB (>> the 126 MB L2), so no L2 flush is needed.

JSON line: value (device-resident throughput, CUDA events, max over ranks), e2e (same three
operators through the public API on numpy/pinned HOST rasters: H2D + kernels + D2H inside
the timed region), roofline of the dominant kernel, cpu_baseline (the CPU oracle, i.e. the
reference's algorithm restated in C, on the box's host cores over a bounded sample), clocks.

--impl reference times that CPU oracle instead (all host threads, bounded sample per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "Mcells/s slope+hillshade+focal.mean f32"
RES = (30.0, 30.0)
ALG_BYTES_PER_CELL = 8.0  # 4 B read + 4 B written per cell for each 3x3 float32 operator


# ----------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons every 20 ms.  The process is started before the warm-up
    steps (its start-up takes longer than a short timed region); `mark()` is called when the timed
    region begins and `stop()` right after it ends, and only the rows printed in between are
    reported.  If the timed region is shorter than two samples, the warm-up rows (same kernels,
    same load) are included and `window` says so."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index
        self.i0 = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def wait_first(self, timeout=2.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        self.i0 = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        i1 = len(self.rows)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows, window = self.rows[self.i0:i1], "timed region"
        if len(rows) < 2:
            rows, window = self.rows[:max(i1, 1)], "warm-up + timed region (timed region shorter than two samples)"
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "window": window,
                "reasons": sorted(reasons)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- CPU arm
def cpu_steps(sample, threads, steps, warmup):
    """Time slope + hillshade + focal.mean of the oracle on `sample`; returns seconds per step."""
    import oracle
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.slope(sample, RES[0], RES[1], nthreads=threads)
        oracle.hillshade(sample, 225, 25, nthreads=threads)
        oracle.focal_mean(sample, nthreads=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times


def host_threads():
    """All host cores this process may use (torchrun pins OMP_NUM_THREADS=1, so ask the OS)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def synth_sample(rows, cols):
    """The top-left rows x cols window of the benchmark DEM (same generator, same seed as the GPU
    arm: only the INPUT is produced on the device, nothing of the timed path runs there); a
    host-generated stand-in when no GPU is usable."""
    try:
        import ctypes
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("no GPU")
        from xrspatial_b200 import _lib
        t = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
        _lib.call("xrs_synth_terrain_f32", ctypes.c_void_p(t.data_ptr()), cols * 4, rows, cols, 0, 0, 1235, 0.0, 4000.0,
                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        out = t.cpu().numpy()
        del t
        torch.cuda.empty_cache()
        return out, "top-left %d x %d window of the benchmark DEM (xrs_synth_terrain_f32, seed 1235)" % (rows, cols)
    except Exception:
        return host_sample(rows, cols), "%d x %d host-generated DEM (no GPU available for the synthetic terrain)" % (rows, cols)


def host_sample(rows, cols, seed=1234):
    """Cheap host-side DEM for the CPU arm when no GPU generated one (same statistics)."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((rows, cols), dtype=np.float32)
    z = np.cumsum(np.cumsum(z, axis=0, dtype=np.float64), axis=1)
    z = (z - z.min()) / (z.max() - z.min()) * 4000.0
    return z.astype(np.float32)


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is
    pure Python/Numba and cannot travel to the GPU box) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = host_threads()
    rows = cols = args.cpu_sample
    sample, sample_desc = synth_sample(rows, cols)
    times = cpu_steps(sample, threads, args.steps, args.warmup)
    dt = float(np.mean(times))
    value = 3.0 * rows * cols / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "slope+hillshade+focal.mean on a float32 fBm-like DEM, res=(30,30): CPU oracle "
                               "(C restatement of the reference's Numba/NumPy kernels, which cannot travel to "
                               "this box), bounded sample per step",
                   "sample": sample_desc, "arithmetic": "Horn sums in f64 as Numba promotes them"},
        "cpu_baseline": {"value": value, "unit": "Mcells/s", "cores": threads, "kind": "port",
                         "sample": sample_desc + ", all three operators, OpenMP over rows"},
        "e2e": {"value": value, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import ctypes
    import xrspatial_b200 as xb
    from xrspatial_b200 import _lib
    from xrspatial_b200.stripes import RowStripes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    side = args.raster or (32768 if n_gpus == 1 else 65536)
    H = W = side
    stripes = RowStripes(H, W, radius=1, device=dev)
    lib = _lib.lib()
    # synthetic DEM: pure function of (seed, global row, col) -> identical for any striping
    interior = stripes.interior
    _lib.call("xrs_synth_terrain_f32", ctypes.c_void_p(interior.data_ptr()), W * 4, stripes.h, W, stripes.y0, 0,
              1235, 0.0, 4000.0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()

    attrs = {"res": RES}
    hp = stripes.buf.shape[0]
    agg = xb.DataArray(stripes.buf, dims=("y", "x"), attrs=attrs)
    outs = {}

    def step(events=None):
        stripes.exchange()
        if events is not None:
            events[0].record()
        outs["slope"] = xb.slope(agg).data
        if events is not None:
            events[1].record()
        outs["hillshade"] = xb.hillshade(agg).data
        if events is not None:
            events[2].record()
        outs["mean"] = xb.mean(agg).data
        if events is not None:
            events[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if rank == 0 and n_gpus == 1 and not args.skip_host:
        parity = run_parity_gate(xb, stripes, attrs)     # before any timing
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()           # before the warm-up: nvidia-smi needs ~0.1 s to print its first row
        sampler.wait_first()
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    assert lib.xrs_debug_last_used_tma() == 1, "TMA kernels were not selected"

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    t_begin.record()
    for i in range(args.steps):
        step(ev[i])
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_begin.elapsed_time(t_end)
    kt = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(3)] for e in ev])  # ms per kernel
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        k = torch.tensor(kt.mean(0), dtype=torch.float64, device=dev)
        dist.all_reduce(k, op=dist.ReduceOp.MAX)
        kmean = k.cpu().numpy()
    else:
        kmean = kt.mean(0)
    ms_per_step = total_ms / args.steps
    cells_step = 3.0 * H * W
    value = cells_step / (ms_per_step * 1e-3) / 1e6

    # roofline of the dominant (slowest) kernel, per launch, algorithmic bytes only
    names = ["slope", "hillshade", "focal.mean"]
    dom = int(np.argmax(kmean))
    peak, peak_src = measured_peak_gbs()
    rows_launch = hp  # the kernel processes the padded stripe
    alg_bytes = ALG_BYTES_PER_CELL * rows_launch * W
    achieved = alg_bytes / (kmean[dom] * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu capture
    # (profiles/r01_bench_ncu_summary.md, taken on a 32768 x 32768 launch), scaled to the rows of
    # this launch
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
            traffic = json.load(f).get(names[dom])
        if traffic is not None:
            traffic = traffic * (float(rows_launch) * W) / (32768.0 * 32768.0)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "per_kernel_ms": dict(zip(names, [float(x) for x in kmean])),
                "per_kernel_frac": dict(zip(names, [float(alg_bytes / (x * 1e-3) / 1e9 / peak) for x in kmean]))}

    e2e = None
    cpu = None
    if rank == 0 and n_gpus == 1:
        e2e = None if args.skip_host else run_e2e(xb, stripes, H, W, attrs, args)
        cpu = None if args.skip_host else run_cpu_baseline(stripes, args)
    elif rank == 0:
        e2e = {"value": None, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "host-buffer path is measured at N=1 (it drives one GPU per call)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "slope+hillshade+focal.mean on a %d x %d float32 fBm-like DEM, res=(30,30)%s"
                                   % (H, W, "" if n_gpus == 1 else ", row-striped over %d GPUs with 1-row NCCL "
                                      "halo exchange per step" % n_gpus),
                       "raster": [H, W], "cells_per_step": cells_step, "parallelism": "rows/%d" % n_gpus,
                       "arithmetic": "f32 in/out; Horn sums and focal sums in f64 like the reference's Numba kernels",
                       "l2": "inputs (%.1f GiB per GPU) are larger than L2, no flush" % (hp * W * 4 / 2 ** 30)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "parity_gate": parity,
            "gpu_launches": 3 * args.steps,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(xb, stripes, H, W, attrs, args):
    """Same three operators through the public API on HOST (pinned) rasters."""
    import torch
    from xrspatial_b200 import _hostmem
    steps = max(1, min(args.steps, args.e2e_steps))
    eh = H
    z = None
    while eh >= 1024:
        try:
            z = _hostmem.empty((eh, W), np.float32)
            break
        except Exception:
            eh //= 2
    if z is None:
        return {"value": None, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "could not allocate pinned host memory"}
    torch.from_numpy(z).copy_(stripes.interior[:eh])
    torch.cuda.synchronize()
    hagg = xb.DataArray(z, dims=("y", "x"), attrs=attrs)

    def one():
        a = xb.slope(hagg).data
        b = xb.hillshade(hagg).data
        c = xb.mean(hagg).data
        return a, b, c

    res = one()  # warm-up: allocates the pinned result blocks and the device slots
    assert isinstance(res[0], np.ndarray) and res[2].dtype == np.float64
    del res
    t0 = time.perf_counter()
    for _ in range(steps):
        res = one()
        del res
    dt = (time.perf_counter() - t0) / steps
    cells = 3.0 * eh * W
    return {"value": cells / dt / 1e6, "unit": "Mcells/s", "steps": steps, "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": int(3 * eh * W * 4), "d2h_bytes_per_step": int(eh * W * (4 + 4 + 8)),
            "raster": [eh, W],
            "api": "xrspatial_b200.slope/hillshade/mean on numpy DataArrays in pinned host memory "
                   "(xrs_host_stencil: chunked H2D -> kernel -> D2H pipeline); focal.mean returns float64 "
                   "like the reference's numpy path"}


def run_parity_gate(xb, stripes, attrs, n=2048):
    """parity gate before timing -- the three benchmark operators on the top-left
    n x n window of the benchmark DEM against the CPU oracle (|gpu - ref| <= 1e-5 |ref| + 1e-6,
    identical NaN masks)."""
    import oracle
    oracle.build()
    n = int(min(n, stripes.h, stripes.W))
    win = stripes.interior[:n, :n].contiguous()
    host = win.cpu().numpy()
    agg = xb.DataArray(win, dims=("y", "x"), attrs=attrs)
    threads = host_threads()
    worst = 0.0
    for name, got, ref in (
            ("slope", xb.slope(agg).data, oracle.slope(host, RES[0], RES[1], nthreads=threads)),
            ("hillshade", xb.hillshade(agg).data, oracle.hillshade(host, 225, 25, nthreads=threads)),
            ("focal.mean", xb.mean(agg).data, oracle.focal_mean(host, nthreads=threads))):
        g = got.cpu().numpy().astype(np.float64)
        r = np.asarray(ref, dtype=np.float64)
        if not np.array_equal(np.isnan(g), np.isnan(r)):
            raise AssertionError("parity gate: NaN mask of %s differs from the oracle" % name)
        m = ~np.isnan(r)
        e = float((np.abs(g[m] - r[m]) / (1e-5 * np.abs(r[m]) + 1e-6)).max())
        if e > 1.0:
            raise AssertionError("parity gate: %s is %.3g x outside the tolerance" % (name, e))
        worst = max(worst, e)
    return {"checked": True, "window": [n, n], "operators": ["slope", "hillshade", "focal.mean"],
            "tolerance": "|gpu-ref| <= 1e-5*|ref| + 1e-6, NaN masks identical", "worst_err_over_tol": worst}


def run_cpu_baseline(stripes, args):
    import oracle
    oracle.build()
    n = min(args.cpu_sample, stripes.h, stripes.W)
    sample = stripes.interior[:n, :n].contiguous().cpu().numpy()
    threads = host_threads()
    t_all = float(np.mean(cpu_steps(sample, threads, 2, 1)))
    t_one = float(np.mean(cpu_steps(sample[: max(256, n // 4)], 1, 1, 1)))
    cells = 3.0 * n * n
    return {"value": cells / t_all / 1e6, "unit": "Mcells/s", "cores": threads, "kind": "port",
            "sample": "top-left %d x %d window of the benchmark DEM, slope+hillshade+focal.mean" % (n, n),
            "single_thread_value": 3.0 * max(256, n // 4) * n / t_one / 1e6,
            "note": "oracle/xrs_oracle.c (C restatement of the reference's Numba/NumPy kernels); "
                    "single_thread_value is what a stock numpy-backed xrspatial call does (ngjit is serial)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--raster", type=int, default=0, help="raster side (default 32768 at N=1, 65536 at N>1)")
    ap.add_argument("--cpu-sample", type=int, default=8192, help="side of the CPU-arm sample window")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--skip-host", action="store_true",
                    help="profiling runs only: skip the e2e (host-buffer) and CPU-baseline legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
