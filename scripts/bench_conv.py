"""Time convolve_2d for uniform (box path) and mixed-weight kernels at one raster size."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import xrspatial_b200 as xb  # noqa: E402
from xrspatial_b200.convolution import convolve_2d  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
import ctypes  # noqa: E402
from xrspatial_b200 import _lib  # noqa: E402
dem = torch.empty((side, side), dtype=torch.float32, device="cuda")
_lib.call("xrs_synth_terrain_f32", ctypes.c_void_p(dem.data_ptr()), side * 4, side, side, 0, 0, 1235, 0.0, 4000.0,
          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


rng = np.random.default_rng(7)
for k in (5, 9, 15, 25):
    ms = timeit(lambda: convolve_2d(dem, np.ones((k, k)) / (k * k)))
    print("uniform k=%2d : %8.3f ms  %7.1f Gcells/s  path=%d" % (k, ms, dem.numel() / ms / 1e6,
                                                                 xb._lib.lib().xrs_debug_last_used_tma()), flush=True)
sub = dem[: side // 8]
for k in (5, 7, 9, 11, 13, 15, 25):
    kern = rng.standard_normal((k, k))
    ms = timeit(lambda: convolve_2d(sub, kern), n=3)
    print("mixed   k=%2d : %8.3f ms  %7.1f Gcells/s  path=%d" % (k, ms, sub.numel() / ms / 1e6,
                                                                 xb._lib.lib().xrs_debug_last_used_tma()), flush=True)

from xrspatial_b200 import focal  # noqa: E402
sub = dem[: side // 4]
agg = xb.DataArray(sub, dims=("y", "x"))
k5 = np.ones((5, 5))
for s in ("mean", "sum", "min", "std"):
    ms = timeit(lambda: focal.apply(agg, k5, func=s), n=3)
    print("focal.apply %-4s 5x5 : %8.3f ms  %7.1f Gcells/s" % (s, ms, sub.numel() / ms / 1e6), flush=True)
ms = timeit(lambda: focal.focal_stats(agg, k5), n=3)
print("focal_stats x7  5x5 : %8.3f ms  %7.1f Gcells/s  path=%d" % (ms, sub.numel() / ms / 1e6,
                                                                  xb._lib.lib().xrs_debug_last_used_tma()), flush=True)
