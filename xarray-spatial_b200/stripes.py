"""Row-striped rasters across GPUs (one process per GPU, torch.distributed).

The reference scales out with dask.map_overlap(depth=r, boundary=nan) (slope.py:94-97): every
chunk is padded with r rows of its neighbours, the operator runs on the padded chunk and the
halo is trimmed.  Here rank g owns rows [y0, y1) of the raster, keeps them in a buffer with r
extra rows above/below, fills those rows from the neighbouring ranks with NCCL send/recv
(`exchange`), runs the single-GPU operator on the padded buffer and returns its own rows.
Raster-edge stripes have no halo on that side, so the kernels' NaN out-of-bounds fill supplies
the reference's raster-edge rule.  The result is identical, bit for bit, to the single-GPU
result (same per-cell arithmetic, exact halos).

The exchange itself is backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
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
        if self.world > 1 and self.h < radius:
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

    def exchange(self):
        """Fill the halo rows from the neighbouring stripes (one batched send/recv group)."""
        if self.world == 1:
            return
        r = self.radius
        ops = []
        if self.rank > 0:
            up = self._global_rank(self.rank - 1)
            ops.append(dist.P2POp(dist.isend, self.buf[self.top:self.top + r], up, self.group))
            ops.append(dist.P2POp(dist.irecv, self.buf[0:self.top], up, self.group))
        if self.rank < self.world - 1:
            dn = self._global_rank(self.rank + 1)
            ops.append(dist.P2POp(dist.isend, self.buf[self.top + self.h - r:self.top + self.h], dn, self.group))
            ops.append(dist.P2POp(dist.irecv, self.buf[self.top + self.h:], dn, self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def apply(self, fn, *args, attrs=None, **kwargs):
        """Run a single-raster operator `fn(DataArray, ...)` on the padded stripe and return the
        rows this rank owns (tensor view, no copy)."""
        agg = DataArray(self.buf, dims=("y", "x"), attrs=attrs or {})
        out = fn(agg, *args, **kwargs).data
        return out[self.top:self.top + self.h]
