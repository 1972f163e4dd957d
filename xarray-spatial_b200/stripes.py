"""Row-striped rasters across GPUs (one process per GPU, torch.distributed).

The reference scales out with dask.map_overlap(depth=r, boundary=nan) (slope.py:94-97; one
map_overlap PER PASS for focal.mean, focal.py:72-75 with the loop at :258-259): every chunk is
padded with r rows of its neighbours, the operator runs on the padded chunk and the halo is
trimmed.  Here rank g owns rows [y0, y1) of the raster, keeps them in a buffer with r extra rows
above/below and fills those rows from the neighbouring ranks with NCCL send/recv (`exchange`).
Raster-edge stripes have no halo on that side, so the kernels' NaN out-of-bounds fill supplies
the reference's raster-edge rule.  The result is identical, bit for bit, to the single-GPU
result (same per-cell arithmetic, exact halos).

`apply` overlaps the exchange with the arithmetic: the send/recv group is started first, the
operator runs on the rows this rank owns while the halo rows travel (every output row at least
r rows away from a stripe boundary is already final), and only after the current stream has been
made to wait for the exchange are the two boundary bands (3r rows each, r outputs) recomputed
with their halos and patched in.  `mean(passes=n)` exchanges before EVERY pass -- the halo rows
of a pass's output are not the neighbour's output rows -- and `convolve` / any operator of
radius r works the same way with `radius=r` stripes.

The exchange itself is backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from ._xr import DataArray


def split_rows(H, world):
    """Row ranges [(y0, y1)] of `world` near-equal stripes."""
    base, rem = divmod(H, world)
    out, y = [], 0
    for r in range(world):
        h = base + (1 if r < rem else 0)
        out.append((y, y + h))
        y += h
    return out


class RowStripes(object):
    def __init__(self, H, W, radius=1, group=None, device=None, dtype=torch.float32):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.H, self.W, self.radius = H, W, radius
        self.y0, self.y1 = split_rows(H, self.world)[self.rank]
        self.h = self.y1 - self.y0
        if self.world > 1 and min(b - a for a, b in split_rows(H, self.world)) < radius:
            raise ValueError("stripes must be at least `radius` rows tall")
        self.top = radius if self.rank > 0 else 0
        self.bot = radius if self.rank < self.world - 1 else 0
        self.buf = torch.empty((self.top + self.h + self.bot, W), dtype=dtype, device=device)

    # views ---------------------------------------------------------------------------
    @property
    def interior(self):
        return self.buf[self.top:self.top + self.h]

    def _global_rank(self, r):
        if self.group is None:
            return r
        return dist.get_global_rank(self.group, r)

    # halo exchange -------------------------------------------------------------------
    def start_exchange(self, buf=None):
        """Start filling the halo rows of `buf` (default: the stripe buffer; any tensor with the
        same padded layout) from the neighbouring stripes: one batched send/recv group.  Returns
        the pending requests; `finish_exchange` makes the CURRENT STREAM wait for them (NCCL: no
        host block), so work enqueued in between overlaps the transfer."""
        if self.world == 1:
            return []
        buf = self.buf if buf is None else buf
        if tuple(buf.shape) != tuple(self.buf.shape):
            raise ValueError("exchange needs a tensor with the padded stripe layout %r" % (tuple(self.buf.shape),))
        r = self.radius
        ops = []
        if self.rank > 0:
            up = self._global_rank(self.rank - 1)
            ops.append(dist.P2POp(dist.isend, buf[self.top:self.top + r], up, self.group))
            ops.append(dist.P2POp(dist.irecv, buf[0:self.top], up, self.group))
        if self.rank < self.world - 1:
            dn = self._global_rank(self.rank + 1)
            ops.append(dist.P2POp(dist.isend, buf[self.top + self.h - r:self.top + self.h], dn, self.group))
            ops.append(dist.P2POp(dist.irecv, buf[self.top + self.h:], dn, self.group))
        return dist.batch_isend_irecv(ops)

    @staticmethod
    def finish_exchange(reqs):
        for req in reqs:
            req.wait()

    def exchange(self, buf=None):
        """Fill the halo rows from the neighbouring stripes (one batched send/recv group)."""
        self.finish_exchange(self.start_exchange(buf))

    # operators -----------------------------------------------------------------------
    def apply(self, fn, *args, attrs=None, overlap=True, exchange=True, **kwargs):
        """Run a single-raster operator `fn(DataArray, ...)` of radius `self.radius` on this stripe
        and return the rows this rank owns (h x W tensor).

        exchange=True refreshes the halos first; with overlap=True the operator runs on the owned
        rows while the halos travel and the two boundary bands are patched in afterwards (see the
        module docstring).  exchange=False assumes the halos are current (after `exchange()`)."""
        r, h, top, bot = self.radius, self.h, self.top, self.bot
        attrs = attrs or {}

        def run(t):
            return fn(DataArray(t, dims=("y", "x"), attrs=attrs), *args, **kwargs).data

        if self.world == 1 or not exchange or not overlap or h < 3 * r:
            if exchange:
                self.exchange()
            return run(self.buf)[top:top + h]
        reqs = self.start_exchange()
        out = run(self.interior)                    # rows [r, h - r) are final (and true raster edges)
        self.finish_exchange(reqs)
        if top:
            band = run(self.buf[0:top + 2 * r])     # halo + first 2r owned rows -> outputs of rows [0, r)
            out[0:r] = band[r:2 * r]
        if bot:
            band = run(self.buf[top + h - 2 * r:top + h + bot])
            out[h - r:h] = band[r:2 * r]
        return out

    def mean(self, passes=1, excludes=(np.nan,), attrs=None):
        """focal.mean(passes=n) over the striped raster: one halo exchange before every pass
        (focal.py:72-75, 258-259).  Returns the owned rows after the last pass."""
        from .focal import mean as _mean
        if self.radius < 1:
            raise ValueError("focal.mean needs stripes of radius >= 1")
        cur = self.buf
        for _ in range(int(passes)):
            self.exchange(cur)
            cur = _mean(DataArray(cur, dims=("y", "x"), attrs=attrs or {}), passes=1, excludes=list(excludes)).data
        return cur[self.top:self.top + self.h]

    def convolve(self, kernel, attrs=None, overlap=True):
        """convolve_2d over the striped raster; the stripes must have radius >= kernel rows // 2."""
        from .convolution import convolution_2d
        kernel = np.asarray(kernel)
        if kernel.shape[0] // 2 > self.radius:
            raise ValueError("kernel of %d rows needs stripes of radius >= %d" % (kernel.shape[0], kernel.shape[0] // 2))
        return self.apply(convolution_2d, kernel, attrs=attrs, overlap=overlap)
