"""torchrun diagnostic: are the exchanged halo rows right, and where do striped outputs differ from a band
recomputed as one raster?  (bench.py's parity gate with a detailed report.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
import xrspatial_b200 as xb
from xrspatial_b200.stripes import RowStripes

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
st = RowStripes(H, W, radius=1, device=dev)
bench.synth_into(st.interior, st.y0)
torch.cuda.synchronize()
st.buf[:st.top].fill_(-1.0)
st.buf[st.top + st.h:].fill_(-2.0)
st.exchange()
torch.cuda.synchronize()
exp = torch.empty_like(st.buf)
bench.synth_into(exp, st.y0 - st.top)
torch.cuda.synchronize()
ok_top = bool(torch.equal(st.buf[:st.top], exp[:st.top]))
ok_bot = bool(torch.equal(st.buf[st.top + st.h:], exp[st.top + st.h:]))
ok_int = bool(torch.equal(st.interior, exp[st.top:st.top + st.h]))
print("rank %d: halo top %s bottom %s interior %s (top=%d bot=%d h=%d)" % (rank, ok_top, ok_bot, ok_int, st.top, st.bot, st.h), flush=True)
attrs = {"res": (30.0, 30.0)}
agg = xb.DataArray(st.buf, dims=("y", "x"), attrs=attrs)
outs = {"slope": xb.slope(agg).data, "hillshade": xb.hillshade(agg).data, "mean": xb.mean(agg).data}
# reference for THIS rank's stripe from the regenerated padded stripe (no communication at all)
eagg = xb.DataArray(exp, dims=("y", "x"), attrs=attrs)
for k, f in (("slope", xb.slope), ("hillshade", xb.hillshade), ("mean", xb.mean)):
    ref = f(eagg).data
    same = torch.equal(ref[st.top:st.top + st.h].view(torch.int32), outs[k][st.top:st.top + st.h].view(torch.int32))
    print("rank %d: %s on exchanged stripe == on regenerated stripe: %s" % (rank, k, same), flush=True)
try:
    res = bench.run_stripe_parity_gate(xb, st, outs, attrs, dist, dev)
    print("rank %d: gate ok" % rank, flush=True)
except AssertionError as e:
    print("rank %d: gate FAILED: %s" % (rank, e), flush=True)
dist.barrier()
dist.destroy_process_group()
