"""Runs slope / hillshade / focal.mean (and optionally more) once each after warm-up, for ncu."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import xrspatial_b200 as xb
from xrspatial_b200 import _lib
side = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ops = sys.argv[3].split(",") if len(sys.argv) > 3 else ["slope", "hillshade", "mean"]
t = torch.empty((side, side), dtype=torch.float32, device="cuda")
_lib.call("xrs_synth_terrain_f32", ctypes.c_void_p(t.data_ptr()), side * 4, side, side, 0, 0, 1235, 0.0, 4000.0,
          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
agg = xb.DataArray(t, dims=("y", "x"), attrs={"res": (30.0, 30.0)})
fns = {"slope": xb.slope, "hillshade": xb.hillshade, "mean": xb.mean, "aspect": xb.aspect, "curvature": xb.curvature, "suite": xb.surface_suite}
for _ in range(reps):
    for o in ops:
        fns[o](agg)
torch.cuda.synchronize()
print("done")
