"""One uniform k x k convolve_2d (and optionally one mixed-weight one) for ncu."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import xrspatial_b200 as xb
from xrspatial_b200 import _lib
from xrspatial_b200.convolution import convolve_2d
side = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ks = [int(k) for k in sys.argv[2].split(",")] if len(sys.argv) > 2 else [9, 25]
t = torch.empty((side, side), dtype=torch.float32, device="cuda")
_lib.call("xrs_synth_terrain_f32", ctypes.c_void_p(t.data_ptr()), side * 4, side, side, 0, 0, 1235, 0.0, 4000.0,
          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
for k in ks:
    kern = np.ones((k, k)) / (k * k)
    for _ in range(2):
        convolve_2d(t, kern)
torch.cuda.synchronize()
print("done")
