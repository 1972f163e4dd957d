"""The (zone, value) pair histogram on the banded benchmark DEM and the stats kernel on noisy zones, for ncu."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from xrspatial_b200 import _lib, zonal as Z
side = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
t = torch.empty((side, side), dtype=torch.float32, device="cuda")
_lib.call("xrs_synth_terrain_f32", ctypes.c_void_p(t.data_ptr()), side * 4, side, side, 0, 0, 1235, 0.0, 4000.0,
          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
yy = torch.arange(side, device="cuda", dtype=torch.int32)[:, None] // (side // 32)
xx = torch.arange(side, device="cuda", dtype=torch.int32)[None, :] // (side // 32)
zones = (yy * 32 + xx).contiguous()
cats = (t * (16.0 / 4000.0)).floor_().clamp_(0, 15)
zirr = (t * (64.0 / 4000.0)).floor_().clamp_(0, 63).to(torch.int32)
for _ in range(2):
    Z.pair_counts(zones, cats)
    Z.hash_partials(zirr, t)
torch.cuda.synchronize()
print("done")
