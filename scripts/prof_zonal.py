import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import xrspatial_b200 as xb
from xrspatial_b200 import _lib, zonal as Z
from xrspatial_b200.convolution import convolve_2d
side = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
t = torch.empty((side, side), dtype=torch.float32, device="cuda")
_lib.call("xrs_synth_terrain_f32", ctypes.c_void_p(t.data_ptr()), side * 4, side, side, 0, 0, 1235, 0.0, 4000.0,
          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
yy = torch.arange(side, device="cuda", dtype=torch.int32)[:, None] // (side // 32)
xx = torch.arange(side, device="cuda", dtype=torch.int32)[None, :] // (side // 32)
zones = (yy * 32 + xx).contiguous()
sel = np.arange(1024, dtype=np.int32)
for _ in range(2):
    Z.hash_partials(zones, t)
    convolve_2d(t[: side // 4], np.ones((9, 9)) / 81.0)
    xb.aspect(xb.DataArray(t))
torch.cuda.synchronize()
print("done")
