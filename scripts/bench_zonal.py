import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import xrspatial_b200 as xb
from xrspatial_b200 import _lib, zonal as Z
side = 32768
t = torch.empty((side, side), dtype=torch.float32, device="cuda")
_lib.call("xrs_synth_terrain_f32", ctypes.c_void_p(t.data_ptr()), side * 4, side, side, 0, 0, 1235, 0.0, 4000.0,
          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
yy = torch.arange(side, device="cuda", dtype=torch.int32)[:, None] // (side // 32)
xx = torch.arange(side, device="cuda", dtype=torch.int32)[None, :] // (side // 32)
zones = (yy * 32 + xx).contiguous()
P = lambda x: ctypes.c_void_p(x.data_ptr())
cap = 1 << 16
keys = torch.empty(cap, dtype=torch.int64, device="cuda"); count = torch.empty(cap, dtype=torch.int64, device="cuda")
s1 = torch.empty(cap, dtype=torch.float64, device="cuda"); s2 = torch.empty_like(s1); mn = torch.empty_like(s1); mx = torch.empty_like(s1)
ovf = torch.empty(1, dtype=torch.int32, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    _lib.call("xrs_zonal_hash_init", P(keys), P(count), P(s1), P(s2), P(mn), P(mx), cap, P(ovf), st)
    _lib.call("xrs_zonal_hash_accumulate", P(t), 0, P(zones), 2, t.numel(), side, 2000.0, 0, 0.0, P(keys), P(count), P(s1), P(s2), P(mn), P(mx), cap, P(ovf), st)
for _ in range(3): run()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
e[0].record()
for i in range(10):
    run(); e[i+1].record()
torch.cuda.synchronize()
ms = np.median([e[i].elapsed_time(e[i+1]) for i in range(10)])
print("zonal hash kernel (+init): %.3f ms  %.0f GB/s  frac %.3f" % (ms, side*side*8/ms/1e6, side*side*8/ms/1e6/6569.6))
t0=time.perf_counter(); Z.hash_partials(zones, t); torch.cuda.synchronize(); print("hash_partials wall %.3f ms" % ((time.perf_counter()-t0)*1e3))
df = xb.zonal_stats(xb.DataArray(zones), xb.DataArray(t)); print(df.head(3))
