"""Single-GPU emulation of bench.py's N-rank stripe-boundary parity gate (no NCCL): the 65536-wide benchmark
DEM is cut into `world` row stripes, every stripe (with its 1-row halos, as the exchange would deliver
them) is processed on this GPU, and every boundary band is recomputed as one raster and compared bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import xrspatial_b200 as xb
from xrspatial_b200.stripes import split_rows

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
W = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
B = 2048
attrs = {"res": (30.0, 30.0)}
ops = (("slope", xb.slope), ("hillshade", xb.hillshade), ("mean", xb.mean))
rows = split_rows(H, world)
bad = 0
prev = None
for r, (y0, y1) in enumerate(rows):
    top, bot = (1 if r > 0 else 0), (1 if r < world - 1 else 0)
    buf = torch.empty((top + (y1 - y0) + bot, W), dtype=torch.float32, device="cuda")
    bench.synth_into(buf, y0 - top)
    agg = xb.DataArray(buf, dims=("y", "x"), attrs=attrs)
    h = y1 - y0
    own = {k: f(agg).data[top:top + h] for k, f in ops}
    inter = buf[top:top + h]
    if prev is not None:
        pin, pown, py1 = prev
        band = torch.cat([pin[-(B + 1):], inter[:B + 1]], dim=0)
        bagg = xb.DataArray(band, dims=("y", "x"), attrs=attrs)
        for k, f in ops:
            got = f(bagg).data[1:2 * B + 1]
            exp = torch.cat([pown[k][-B:], own[k][:B]], dim=0)
            same = torch.equal(got.view(torch.int32), exp.view(torch.int32))
            if not same:
                d = (got.view(torch.int32) != exp.view(torch.int32))
                nanboth = torch.isnan(got) & torch.isnan(exp)
                d &= ~nanboth
                idx = torch.nonzero(d)
                print("boundary %d/%d op %s: %d cells differ; rows (band coords) %s cols %s" % (
                    r - 1, r, k, int(d.sum()), idx[:, 0].unique()[:10].tolist(), idx[:, 1].unique()[:10].tolist()), flush=True)
                if idx.numel():
                    i, j = idx[0].tolist()
                    print("   first: got %r exp %r" % (got[i, j].item(), exp[i, j].item()))
                bad += 1
        del band
    prev = (inter[-(B + 1):].clone(), {k: v[-B:].clone() for k, v in own.items()}, y1)
    del buf, own
    torch.cuda.empty_cache()
print("world %d: %d boundary/op mismatches" % (world, bad))
