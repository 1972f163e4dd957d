import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import xrspatial_b200 as xb
from xrspatial_b200 import focal
from xrspatial_b200.convolution import convolve_2d
import oracle as o
z = (np.random.default_rng(0).standard_normal((260, 384)).cumsum(0).cumsum(1)).astype(np.float32)
t = torch.from_numpy(z).cuda()
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("focal", "both"):
    k = np.ones((3, 3))
    r = focal.apply(xb.DataArray(t), k).data
    torch.cuda.synchronize()
    print("focal stat tma ok", xb._lib.lib().xrs_debug_last_used_tma(), np.allclose(r.cpu().numpy(), o.focal_apply(z, k, "mean"), rtol=1e-6, equal_nan=True))
if which in ("conv", "both"):
    for k in (3, 9, 25):
        kern = np.ones((k, k)) / (k * k)
        r = convolve_2d(t, kern)
        torch.cuda.synchronize()
        ref = o.convolve_2d(z, kern)
        print("conv", k, np.allclose(r.cpu().numpy(), ref, rtol=1e-5, atol=1e-3, equal_nan=True))
