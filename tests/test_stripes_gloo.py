"""World-size-2/3 tests of the multi-GPU plumbing on CPU with the gloo backend: the halo
exchange fills every stripe's halo rows with the neighbouring stripes' rows, and zonal
partials combined with all_reduce finalise to the single-raster statistics."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, radius, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xrspatial_b200.stripes import RowStripes
        from xrspatial_b200 import zonal
        full = torch.arange(H * W, dtype=torch.float32).reshape(H, W)
        st = RowStripes(H, W, radius=radius, device="cpu")
        st.buf.fill_(-1.0)
        st.interior.copy_(full[st.y0:st.y1])
        st.exchange()
        lo, hi = st.y0 - st.top, st.y1 + st.bot
        ok_halo = bool(torch.equal(st.buf, full[lo:hi]))
        # zonal partials of this stripe (computed with numpy here; the CUDA kernel is tested on
        # the GPU) -> all_reduce -> finalize == whole-raster statistics
        vals = (full[st.y0:st.y1] % 17).double().numpy()
        zones = ((torch.arange(H)[:, None] // 3) * 2 + (torch.arange(W)[None, :] // (W // 2))).numpy()[st.y0:st.y1]
        ids = np.unique(((np.arange(H)[:, None] // 3) * 2 + (np.arange(W)[None, :] // (W // 2))))
        piv = 5.0
        cnt = torch.tensor([(zones == z).sum() for z in ids], dtype=torch.int64)
        s1 = torch.tensor([(vals[zones == z] - piv).sum() for z in ids], dtype=torch.float64)
        s2 = torch.tensor([((vals[zones == z] - piv) ** 2).sum() for z in ids], dtype=torch.float64)
        mn = torch.tensor([vals[zones == z].min() if (zones == z).any() else np.inf for z in ids], dtype=torch.float64)
        mx = torch.tensor([vals[zones == z].max() if (zones == z).any() else -np.inf for z in ids], dtype=torch.float64)
        for t, op in ((cnt, dist.ReduceOp.SUM), (s1, dist.ReduceOp.SUM), (s2, dist.ReduceOp.SUM),
                      (mn, dist.ReduceOp.MIN), (mx, dist.ReduceOp.MAX)):
            dist.all_reduce(t, op=op)
        cols = zonal.finalize(dict(count=cnt.numpy(), s1=s1.numpy(), s2=s2.numpy(), min=mn.numpy(), max=mx.numpy()),
                              np.full(len(ids), piv), ["mean", "var", "count", "min", "max", "sum"])
        # the product's own combiner: per-stripe tables with DIFFERENT id sets -> all-gather of
        # the ids + dense AllReduce (zonal.allreduce_tables)
        present = np.unique(zones)
        loc = dict(count=np.array([(zones == z).sum() for z in present], dtype=np.int64),
                   s1=np.array([(vals[zones == z] - piv).sum() for z in present]),
                   s2=np.array([((vals[zones == z] - piv) ** 2).sum() for z in present]),
                   min=np.array([vals[zones == z].min() for z in present]),
                   max=np.array([vals[zones == z].max() for z in present]))
        uids, upart = zonal.allreduce_tables(present, loc, torch.device("cpu"), dist.group.WORLD)
        ok_tab = bool(np.array_equal(uids, ids) and np.array_equal(upart["count"], cnt.numpy())
                      and np.allclose(upart["s1"], s1.numpy()) and np.allclose(upart["s2"], s2.numpy())
                      and np.array_equal(upart["min"], mn.numpy()) and np.array_equal(upart["max"], mx.numpy()))
        allv = (full % 17).double().numpy()
        allz = (np.arange(H)[:, None] // 3) * 2 + (np.arange(W)[None, :] // (W // 2))
        ok_z = True
        for i, z in enumerate(ids):
            v = allv[allz == z]
            ok_z &= bool(np.isclose(cols["mean"][i], v.mean()) and np.isclose(cols["var"][i], v.var())
                         and cols["count"][i] == v.size and cols["min"][i] == v.min() and cols["max"][i] == v.max()
                         and np.isclose(cols["sum"][i], v.sum()))
        q.put((rank, ok_halo, ok_z and ok_tab, (st.y0, st.y1, st.top, st.bot)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,radius", [(2, 1), (3, 2)])
def test_halo_exchange_and_zonal_allreduce(world, radius):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    H, W = 13, 8
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, radius, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_halo, ok_z, geom in sorted(res):
        assert ok_halo, "halo rows wrong on rank %d %r" % (rank, geom)
        assert ok_z, "zonal all_reduce wrong on rank %d" % rank
        y0, y1, top, bot = geom
        assert top == (radius if rank > 0 else 0) and bot == (radius if rank < world - 1 else 0)
