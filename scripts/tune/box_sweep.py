"""Sweep of the streaming box kernel's tuning knobs (rows per scan batch, prefetch rows, CTAs per SM) on
one GPU.  usage: box_sweep.py [side]   -> prints ms / Gcells/s per configuration; outputs are compared
bit for bit with the first configuration of each k."""
import ctypes, itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from xrspatial_b200 import _lib
from xrspatial_b200.convolution import convolve_2d

side = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
t = torch.empty((side, side), dtype=torch.float32, device="cuda")
_lib.call("xrs_synth_terrain_f32", ctypes.c_void_p(t.data_ptr()), side * 4, side, side, 0, 0, 1235, 0.0, 4000.0,
          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
cfgs = [(2, 16), (2, 8), (3, 8), (3, 4), (4, 6), (1, 16)]
for k in (5, 9, 15, 25):
    kern = np.ones((k, k)) / (k * k)
    ref = None
    for rows, (ctas, pre) in itertools.product((4, 6, 8), cfgs):
        os.environ["XRS_BOX_ROWS"] = str(rows)
        os.environ["XRS_BOX_CTAS"] = str(ctas)
        os.environ["XRS_BOX_PREFETCH"] = str(pre)
        out = convolve_2d(t, kern)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for i in range(3):
            ev[i].record()
            out = convolve_2d(t, kern)
        ev[3].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(3))[1]
        li = _lib.lib().xrs_debug_last_launch
        same = None
        if ref is None:
            ref = out.clone()
        else:
            same = bool(torch.equal(out.view(torch.int32), ref.view(torch.int32)))
        print("k=%2d rows=%d ctas<=%d prefetch=%2d : %.3f ms  %.1f Gcells/s  same=%s" %
              (k, rows, ctas, pre, ms, side * side / ms / 1e6, same), flush=True)
        del out
