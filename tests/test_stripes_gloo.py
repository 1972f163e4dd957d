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


def _apply_worker(rank, world, port, H, W, radius, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch.nn.functional as F
        from xrspatial_b200._xr import DataArray
        from xrspatial_b200.stripes import RowStripes
        k = 2 * radius + 1

        def box(agg, scale=1.0):       # any translation-invariant operator of radius `radius`
            x = agg.data.double()
            y = F.conv2d(x[None, None], torch.ones(1, 1, k, k, dtype=torch.float64), padding=radius)[0, 0] * scale
            return DataArray(y.float(), dims=agg.dims, attrs=agg.attrs)

        g = torch.Generator().manual_seed(5)
        full = torch.randint(-50, 50, (H, W), generator=g).float()      # small integers: every sum is exact
        whole = box(DataArray(full, dims=("y", "x")), 2.0).data
        st = RowStripes(H, W, radius=radius, device="cpu")
        st.buf.fill_(float("nan"))
        st.interior.copy_(full[st.y0:st.y1])
        a = st.apply(box, 2.0, overlap=True)
        st.buf[:st.top] = float("nan")          # stale halos: apply() must refresh them itself
        st.buf[st.top + st.h:] = float("nan")
        b = st.apply(box, 2.0, overlap=False)
        st.exchange()
        c = st.apply(box, 2.0, exchange=False)
        ref = whole[st.y0:st.y1]
        # multi-pass with one exchange per pass (the structure of RowStripes.mean)
        cur, ref2 = st.buf.clone(), full
        for _ in range(3):
            st.exchange(cur)
            cur = box(DataArray(cur, dims=("y", "x")), 1.0 / 64).data
            ref2 = box(DataArray(ref2, dims=("y", "x")), 1.0 / 64).data
        ok_multi = bool(torch.equal(cur[st.top:st.top + st.h], ref2[st.y0:st.y1]))
        q.put((rank, bool(torch.equal(a, ref)), bool(torch.equal(b, ref)), bool(torch.equal(c, ref)), ok_multi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,radius,H", [(2, 1, 23), (3, 3, 40), (3, 12, 130)])
def test_apply_overlapped_and_per_pass_exchange_are_partition_invariant(world, radius, H):
    """general_checks.py:124-131 (numpy == dask) for the stripe plumbing itself: the overlapped
    apply (owned rows first, boundary bands patched after the exchange), the plain apply and a 3-pass
    chain with one exchange per pass all reproduce the single-raster result exactly, for radii 1, 3
    and 12 (the k = 25 halo), with a stand-in operator on the CPU."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_apply_worker, args=(r, world, port, H, 16, radius, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_a, ok_b, ok_c, ok_multi in sorted(res):
        assert ok_a, "overlapped apply differs on rank %d" % rank
        assert ok_b, "plain apply differs on rank %d" % rank
        assert ok_c, "apply(exchange=False) differs on rank %d" % rank
        assert ok_multi, "3 passes with per-pass exchange differ on rank %d" % rank
