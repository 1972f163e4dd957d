"""xrspatial.zonal.stats on the B200 backend (reference: zonal.py:422-667).

One streaming pass (xrs_zonal_hash_accumulate) discovers the zone ids and produces per-zone count /
sum / sum-of-squares / min / max partials; mean, std (ddof=0) and var are finalised from them in
float64.  `majority` counts (zone, value) pairs in a second pass (hash table; a device sort when
the values are too varied for the table).  Custom callables (`stats_funcs` given as a dict, exactly
like the reference: zonal.py:640-642) run per zone on the zone's valid values, grouped on the device
by one sort.  With `comm` (a torch.distributed process group) the partials of row-striped rasters
are combined with AllReduce before finalisation.
"""
import ctypes

import numpy as np
import pandas as pd

from . import _lib
from ._xr import DataArray, Dataset
from .utils import (ArrayTypeFunctionMapping, as_device_tensor, like_container, stream_ptr,
                    validate_arrays)

_DEFAULT_STATS = ("mean", "max", "min", "sum", "std", "var", "count", "majority")
_PARTIAL_STATS = ("mean", "max", "min", "sum", "std", "var", "count")


def _dtype_code(t):
    import torch
    return {torch.float32: 0, torch.float64: 1, torch.int32: 2, torch.int64: 3}.get(t.dtype)


def _prepare(t, allow):
    """Cast a device tensor to a dtype the kernel reads natively."""
    import torch
    if t.dtype in allow:
        return t.contiguous()
    if t.dtype.is_floating_point:
        return t.to(torch.float64 if t.dtype == torch.float64 else torch.float32).contiguous()
    if t.dtype in (torch.int8, torch.int16, torch.uint8, torch.bool):
        return t.to(torch.int32).contiguous()
    return t.to(torch.int64).contiguous()


_EMPTY_KEY = -(1 << 63)


def _sample_pivot(values_t, comm=None):
    """One global shift p keeps sum((v-p)^2) well conditioned; the mean of a small strided sample
    is enough (one tiny device-to-host copy)."""
    import torch
    flat = values_t.reshape(-1)
    step = max(1, flat.numel() // 4096)
    sample = flat[::step][:4096].to(torch.float64).cpu().numpy()
    sample = sample[np.isfinite(sample)]
    p0 = float(sample.mean()) if sample.size else 0.0
    if comm is not None:
        import torch.distributed as dist
        pt = torch.tensor([p0], dtype=torch.float64, device=values_t.device)
        dist.broadcast(pt, src=dist.get_global_rank(comm, 0), group=comm)
        p0 = float(pt.item())
    return p0


_MAX_OUT = 4096     # zones returned by the one-copy fast path of hash_partials


def hash_partials(zones_t, values_t, nodata_values=None, comm=None, cap=1 << 16, table=None):
    """One streaming pass that discovers the zone ids and accumulates their partials
    (xrs_zonal_hash_run: pivot sampling, table init, accumulation and compaction are enqueued
    back to back; ONE device-to-host copy -- a 3-double header + 6 x 4096 doubles -- is the only
    synchronisation).  Returns (ids, part, pivot): ids = sorted unique finite zone values present
    in the raster (numpy, in the zones dtype), part = dict of numpy arrays aligned with ids
    (count int64; s1, s2, min, max float64), pivot = the scalar shift of s1/s2.
    With `comm`, the tables of all ranks are merged by id (and share one pivot).
    `table`: a dict that receives the device-side hash table (keys, cap) for `second_pass_partials`."""
    import torch
    dev = values_t.device
    hint = _sample_pivot(values_t, comm) if comm is not None else None
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    if values_t.numel() == 0:
        zdt = np.float32 if zones_t.dtype == torch.float32 else np.float64 if zones_t.dtype == torch.float64 else \
            np.int32 if zones_t.dtype == torch.int32 else np.int64
        e = np.zeros(0)
        return np.zeros(0, zdt), dict(count=np.zeros(0, np.int64), s1=e, s2=e.copy(), min=e.copy(), max=e.copy()), 0.0
    while True:
        # one blob: rows = keys, count (int64 bit patterns), s1, s2, min, max
        blob = torch.empty((6, cap), dtype=torch.float64, device=dev)
        packed = torch.empty(3 + 6 * _MAX_OUT, dtype=torch.float64, device=dev)
        flags = torch.empty(2, dtype=torch.int32, device=dev)
        keys, count = blob[0].view(torch.int64), blob[1].view(torch.int64)
        with torch.cuda.device(dev):
            _lib.call("xrs_zonal_hash_run", P(values_t), _dtype_code(values_t), P(zones_t), _dtype_code(zones_t),
                      values_t.numel(), int(values_t.shape[-1]) if values_t.dim() else 1,
                      0 if nodata_values is None else 1, 0.0 if nodata_values is None else float(nodata_values),
                      0 if hint is None else 1, 0.0 if hint is None else float(hint),
                      P(keys), P(count), P(blob[2]), P(blob[3]), P(blob[4]), P(blob[5]), cap,
                      P(packed), _MAX_OUT, P(flags), stream_ptr(values_t))
        host = packed.cpu().numpy()                    # the one synchronising copy
        n_used, overflow, pivot = int(host[0]), int(host[1]), float(host[2])
        if overflow == 0:
            break
        if cap >= (1 << 24):
            raise NotImplementedError("more than 16M distinct zones are not supported")
        cap *= 16
    if n_used <= _MAX_OUT:
        rows = host[3:].reshape(6, _MAX_OUT)[:, :n_used]
    else:                                               # many zones: gather the used slots of the table itself
        used = torch.nonzero(keys != _EMPTY_KEY).reshape(-1)
        rows = blob[:, used].cpu().numpy()
    k = np.ascontiguousarray(rows[0]).view(np.int64)
    part = dict(count=np.ascontiguousarray(rows[1]).view(np.int64).copy(), s1=rows[2].copy(), s2=rows[3].copy(),
                min=rows[4].copy(), max=rows[5].copy())
    ids = _keys_to_ids(k, zones_t.dtype)
    if table is not None:
        table.update(keys=keys, cap=cap)
    if comm is not None:
        ids, part = allreduce_tables(ids, part, dev, comm)
    order = np.argsort(ids, kind="stable")
    return ids[order], {n: a[order] for n, a in part.items()}, pivot


def _keys_to_ids(k, zones_dtype):
    """int64 table keys -> zone ids in the zones' dtype (float zones: the key is the float64 bit pattern)"""
    import torch
    if zones_dtype.is_floating_point:
        return k.view(np.float64).astype(np.float32 if zones_dtype == torch.float32 else np.float64)
    return k.astype(np.int32 if zones_dtype == torch.int32 else np.int64)


def second_pass_partials(zones_t, values_t, table, ids, means, nodata_values=None, comm=None):
    """numpy's two-pass variance for float64 rasters: a second streaming pass over the hash table
    `hash_partials(..., table=...)` left on the device, with the sums taken about every zone's own
    mean (xrs_zonal_hash_second_pass; the means are scattered to the table's slots on the device, no
    host round trip beyond the one the first pass needed).  `ids`: sorted zone ids (numpy), `means`:
    float64, aligned -- over row stripes the merged ids / global means.  Returns the partials
    (count, s1, s2, min, max about `means`) aligned with `ids`."""
    import torch
    dev = values_t.device
    nz = len(ids)
    part = dict(count=np.zeros(nz, np.int64), s1=np.zeros(nz), s2=np.zeros(nz),
                min=np.full(nz, np.inf), max=np.full(nz, -np.inf))
    if nz and values_t.numel():
        keys, cap = table["keys"], table["cap"]
        fz = zones_t.dtype.is_floating_point
        ids_t = torch.as_tensor(np.asarray(ids, dtype=np.float64 if fz else np.int64), device=dev)
        means_t = torch.as_tensor(np.asarray(means, dtype=np.float64), device=dev)
        slot_ids = keys.view(torch.float64) if fz else keys
        idx = torch.searchsorted(ids_t, slot_ids).clamp_(max=nz - 1)
        pivots = means_t[idx].contiguous()                 # empty slots get some zone's mean: never read
        blob = torch.empty((5, cap), dtype=torch.float64, device=dev)
        packed = torch.empty(3 + 6 * _MAX_OUT, dtype=torch.float64, device=dev)
        flags = torch.empty(2, dtype=torch.int32, device=dev)
        count = blob[0].view(torch.int64)
        P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        with torch.cuda.device(dev):
            _lib.call("xrs_zonal_hash_second_pass", P(values_t), _dtype_code(values_t), P(zones_t),
                      _dtype_code(zones_t), values_t.numel(), int(values_t.shape[-1]) if values_t.dim() else 1,
                      0 if nodata_values is None else 1, 0.0 if nodata_values is None else float(nodata_values),
                      P(keys), P(pivots), P(count), P(blob[1]), P(blob[2]), P(blob[3]), P(blob[4]), cap,
                      P(packed), _MAX_OUT, P(flags), stream_ptr(values_t))
        host = packed.cpu().numpy()
        n_used = int(host[0])
        if n_used <= _MAX_OUT:
            rows = host[3:].reshape(6, _MAX_OUT)[:, :n_used]
        else:
            used = torch.nonzero(keys != _EMPTY_KEY).reshape(-1)
            rows = torch.cat([keys[used].view(torch.float64)[None], blob[:, used]]).cpu().numpy()
        local = _keys_to_ids(np.ascontiguousarray(rows[0]).view(np.int64), zones_t.dtype)
        pos = np.searchsorted(ids, local)
        part["count"][pos] = np.ascontiguousarray(rows[1]).view(np.int64)
        for i, n in enumerate(("s1", "s2", "min", "max")):
            part[n][pos] = rows[2 + i]
    if comm is not None and nz:
        import torch.distributed as dist
        t = {n: torch.as_tensor(a, device=dev) for n, a in part.items()}
        dist.all_reduce(t["count"], op=dist.ReduceOp.SUM, group=comm)
        dist.all_reduce(t["s1"], op=dist.ReduceOp.SUM, group=comm)
        dist.all_reduce(t["s2"], op=dist.ReduceOp.SUM, group=comm)
        dist.all_reduce(t["min"], op=dist.ReduceOp.MIN, group=comm)
        dist.all_reduce(t["max"], op=dist.ReduceOp.MAX, group=comm)
        part = {n: v.cpu().numpy() for n, v in t.items()}
    return part


class _PairTableOverflow(Exception):
    """more distinct (zone, value) pairs than the hash table is allowed to grow to"""


def pair_counts(zones_t, values_t, nodata_values=None, comm=None, cap=1 << 20, max_cap=1 << 26):
    """(zone ids int64, values float64, counts int64) of every distinct valid (zone, value) pair:
    one pass of xrs_zonal_pair_count over (int32 zone, float32 value) pairs."""
    import torch
    dev = values_t.device
    if zones_t.dtype != torch.int32:
        if zones_t.dtype.is_floating_point:
            zi = zones_t.to(torch.int32)
            if not bool((zi.to(zones_t.dtype) == zones_t)[torch.isfinite(zones_t)].all()):
                raise NotImplementedError("'majority' needs integer-valued zone ids")
        else:
            zi = zones_t.to(torch.int32)
            if not bool((zi.to(zones_t.dtype) == zones_t).all()):
                raise NotImplementedError("'majority' needs zone ids that fit in int32")
        finite_zone = torch.isfinite(zones_t) if zones_t.dtype.is_floating_point else None
    else:
        zi, finite_zone = zones_t, None
    # float32 rasters are read in place (the kernel itself folds -0.0 into 0.0, one value for np.unique):
    # a `+ 0.0` copy here cost a full extra read + write of the raster
    vf = values_t if values_t.dtype == torch.float32 else values_t.to(torch.float32)
    if values_t.dtype != torch.float32 and not bool(((vf.to(values_t.dtype) == values_t) | ~torch.isfinite(values_t)).all()):
        raise NotImplementedError("'majority' needs values that are exact in float32")
    if finite_zone is not None:
        vf = torch.where(finite_zone, vf, torch.full_like(vf, float("nan")))  # cells of NaN zones drop out
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    while True:
        keys = torch.empty(cap, dtype=torch.int64, device=dev)
        count = torch.empty(cap, dtype=torch.int64, device=dev)
        dummy = torch.empty((4, cap), dtype=torch.float64, device=dev)
        ovf = torch.empty(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            st = stream_ptr(vf)
            _lib.call("xrs_zonal_hash_init", P(keys), P(count), P(dummy[0]), P(dummy[1]), P(dummy[2]), P(dummy[3]),
                      cap, P(ovf), st)
            _lib.call("xrs_zonal_pair_count", P(vf.contiguous()), P(zi.contiguous()), vf.numel(),
                      int(vf.shape[-1]) if vf.dim() else 1, 0 if nodata_values is None else 1,
                      0.0 if nodata_values is None else float(nodata_values), P(keys), P(count), cap, P(ovf), st)
        if int(ovf.item()) == 0:
            break
        if cap >= max_cap:
            raise _PairTableOverflow()
        cap = min(cap * 16, max_cap)
    used = torch.nonzero(keys != _EMPTY_KEY).reshape(-1)
    k = keys[used].cpu().numpy()
    c = count[used].cpu().numpy()
    if comm is not None:
        import torch.distributed as dist
        gathered = [None] * dist.get_world_size(comm)
        dist.all_gather_object(gathered, (k, c), group=comm)
        k = np.concatenate([g[0] for g in gathered])
        c = np.concatenate([g[1] for g in gathered])
        k, inv = np.unique(k, return_inverse=True)
        c = np.bincount(inv, weights=c).astype(np.int64)
    zone = (k >> 32).astype(np.int64)
    val = (k & 0xFFFFFFFF).astype(np.uint32).view(np.float32).astype(np.float64)
    return zone, val, c


def _total_order_bits(v32):
    """float32 tensor -> int64 keys in [0, 2^32) whose integer order is the float order."""
    import torch
    bits = v32.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return torch.where(bits >= 0x80000000, 0xFFFFFFFF - bits, bits + 0x80000000)


def _majority_by_sort(zidx, vals, n_zones):
    """Majority per zone by one device sort.  zidx: int64 zone index per valid cell (0..n_zones-1),
    vals: the cells' values (float32: keyed by their order-preserving bit pattern; float64: keyed by
    their rank among the distinct values).  Returns a float64 numpy array of length n_zones (NaN for
    zones without cells): the most frequent value, smallest on ties (zonal.py:56-68)."""
    import torch
    out = np.full(n_zones, np.nan)
    if zidx.numel() == 0:
        return out
    if vals.dtype == torch.float32:
        vkey, table = _total_order_bits(vals), None
        span = 1 << 32
    else:
        table, vkey = torch.unique(vals, return_inverse=True)   # sorted distinct values, ranks
        span = int(table.numel())
    keys = zidx * span + vkey
    del vkey
    keys = torch.sort(keys).values
    uniq, cnt = torch.unique_consecutive(keys, return_counts=True)
    del keys
    zone = torch.div(uniq, span, rounding_mode="floor")
    zid, zinv = torch.unique_consecutive(zone, return_inverse=True)
    best = torch.zeros(zid.numel(), dtype=cnt.dtype, device=cnt.device).scatter_reduce_(0, zinv, cnt, "amax")
    pos = torch.arange(uniq.numel(), device=uniq.device)
    pos = torch.where(cnt == best[zinv], pos, torch.full_like(pos, uniq.numel()))
    first = torch.full((zid.numel(),), uniq.numel(), dtype=pos.dtype, device=pos.device).scatter_reduce_(0, zinv, pos, "amin")
    vk = uniq[first] - zid * span
    if table is None:
        b = torch.where(vk >= 0x80000000, vk - 0x80000000, 0xFFFFFFFF - vk)
        win = b.cpu().numpy().astype(np.uint32).view(np.float32).astype(np.float64)
    else:
        win = table[vk].to(torch.float64).cpu().numpy()
    out[zid.cpu().numpy()] = win
    return out


def majority_by_zone(zones_t, values_t, unique_zones, nodata_values=None, comm=None):
    """float64 numpy array aligned with `unique_zones` (the sorted distinct finite zone ids): per zone
    the most frequent valid value, smallest on ties, NaN for zones without valid cells (zonal.py:56-68
    `_stats_majority` = np.unique + argmax).

    int32 zones with float32-exact values take the (zone, value) pair-count kernel; anything else
    (non-integer or wide zone ids, float64 values that float32 cannot hold, more distinct pairs than
    the table may grow to) is grouped by one device sort."""
    import torch
    uz = np.asarray(unique_zones)
    out = np.full(len(uz), np.nan)
    if len(uz) == 0:
        return out
    vf = values_t.to(torch.float32)
    exact32 = values_t.dtype == torch.float32 or \
        bool(((vf.to(values_t.dtype) == values_t) | ~torch.isfinite(values_t)).all())
    if zones_t.dtype == torch.int32 and exact32:
        try:
            zone, val, c = pair_counts(zones_t, values_t, nodata_values, comm, max_cap=1 << 24)
            order = np.lexsort((val, -c, zone))      # per zone: highest count first, then smallest value
            zone, val = zone[order], val[order]
            first = np.r_[True, zone[1:] != zone[:-1]]
            pos = np.searchsorted(uz.astype(np.int64), zone[first])
            ok = (pos < len(uz)) & (uz.astype(np.int64)[np.minimum(pos, len(uz) - 1)] == zone[first])
            out[pos[ok]] = val[first][ok]
            return out
        except _PairTableOverflow:
            pass
    if comm is not None:
        raise NotImplementedError("'majority' over row stripes needs int32 zones and categorical float32-exact "
                                  "values (the sort-based path is single-GPU)")
    z = zones_t.reshape(-1)
    v = (vf if exact32 else values_t.to(torch.float64)).reshape(-1) + 0.0   # -0.0 and 0.0 are one value
    ok = torch.isfinite(v)
    if nodata_values is not None:
        ok &= v != float(nodata_values)
    ids_t = torch.as_tensor(uz.astype(np.float64), device=z.device)
    zf = z.to(torch.float64)
    if z.dtype.is_floating_point:
        ok &= torch.isfinite(z)
    idx = torch.searchsorted(ids_t, zf).clamp_(max=len(uz) - 1)
    ok &= ids_t[idx] == zf
    return _majority_by_sort(idx[ok], v[ok], len(uz))


def allreduce_tables(ids, part, dev, comm):
    """Combine the per-stripe tables of all ranks: the id lists are all-gathered (a few KB), every
    rank scatters its partials into dense arrays over the sorted union of ids, and the dense
    arrays are combined with NCCL AllReduce (SUM for count / sums, MIN, MAX) -- SURVEY.md 8e."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(comm)
    n_local = torch.tensor([len(ids)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n_local, group=comm)
    nmax = int(max(int(c.item()) for c in counts))
    pad = np.zeros(nmax, dtype=np.float64)
    pad[:len(ids)] = np.asarray(ids, dtype=np.float64)
    mine = torch.as_tensor(pad, device=dev)
    gathered = [torch.empty(nmax, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(gathered, mine, group=comm)
    union = np.unique(np.concatenate([g[:int(c.item())].cpu().numpy() for g, c in zip(gathered, counts)]))
    pos = np.searchsorted(union, np.asarray(ids, dtype=np.float64))
    nz = len(union)
    dense = dict(count=torch.zeros(nz, dtype=torch.int64, device=dev),
                 s1=torch.zeros(nz, dtype=torch.float64, device=dev),
                 s2=torch.zeros(nz, dtype=torch.float64, device=dev),
                 min=torch.full((nz,), float("inf"), dtype=torch.float64, device=dev),
                 max=torch.full((nz,), float("-inf"), dtype=torch.float64, device=dev))
    if len(ids):
        idx = torch.as_tensor(pos, device=dev)
        for k in dense:
            dense[k][idx] = torch.as_tensor(part[k], device=dev)
    if nz:
        dist.all_reduce(dense["count"], op=dist.ReduceOp.SUM, group=comm)
        dist.all_reduce(dense["s1"], op=dist.ReduceOp.SUM, group=comm)
        dist.all_reduce(dense["s2"], op=dist.ReduceOp.SUM, group=comm)
        dist.all_reduce(dense["min"], op=dist.ReduceOp.MIN, group=comm)
        dist.all_reduce(dense["max"], op=dist.ReduceOp.MAX, group=comm)
    return union.astype(np.asarray(ids).dtype), {k: v.cpu().numpy() for k, v in dense.items()}


def finalize(part, pivot, stats_funcs):
    """dict stat -> float64 column; zones without valid cells are NaN (zonal.py:153-162)."""
    cnt = part["count"].astype(np.float64)
    ok = cnt > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        m1 = part["s1"] / cnt
        cols = {}
        for s in stats_funcs:
            if s == "mean":
                c = pivot + m1
            elif s == "sum":
                c = pivot * cnt + part["s1"]
            elif s in ("var", "std"):
                c = np.maximum(part["s2"] / cnt - m1 * m1, 0.0)
                if s == "std":
                    c = np.sqrt(c)
            elif s == "count":
                c = cnt.copy()
            elif s == "max":
                c = part["max"].copy()
            elif s == "min":
                c = part["min"].copy()
            else:
                raise ValueError("Invalid stat name. %s option not supported." % s)
            c = np.where(ok, c, np.nan)
            cols[s] = c
    return cols


def _stats_device(zones, values, zone_ids, stats_funcs, nodata_values, return_type='pandas.DataFrame',
                  comm=None):
    """Device runner (replaces zonal.py:335 `_stats_cupy`)."""
    import torch
    zt = _prepare(as_device_tensor(zones), (torch.int32, torch.int64, torch.float32, torch.float64))
    vt = _prepare(as_device_tensor(values), (torch.float32, torch.float64))
    if vt.dtype not in (torch.float32, torch.float64):
        vt = vt.to(torch.float64)
    if len(vt.shape) > 2:
        raise TypeError('3D inputs not supported for the device backend')
    names = list(stats_funcs)
    table = {}
    unique_zones, part_all, pivot0 = hash_partials(zt, vt, nodata_values, comm=comm, table=table)
    zdtype = unique_zones.dtype
    if zone_ids is None:
        sel = unique_zones
        part = part_all
    else:
        sel = np.array([z for z in np.unique(zone_ids) if z in unique_zones], dtype=zdtype)
        pos = np.searchsorted(unique_zones, sel)
        part = {n: a[pos] for n, a in part_all.items()}
    pivot = np.full(len(sel), pivot0)
    cols = finalize(part, pivot, [s for s in names if s not in ("std", "var", "majority")])
    if "majority" in names:
        maj = majority_by_zone(zt, vt, unique_zones, nodata_values, comm=comm)
        cols["majority"] = maj if zone_ids is None else maj[pos]
    sv = [s for s in names if s in ("std", "var")]
    if sv:
        if vt.dtype == torch.float64 and len(sel):
            # second pass about the per-zone means: numpy's two-pass variance, to ~1e-15
            cnt = part_all["count"].astype(np.float64)
            with np.errstate(invalid="ignore", divide="ignore"):
                means = np.where(cnt > 0, pivot0 + part_all["s1"] / cnt, 0.0)
            part2 = second_pass_partials(zt, vt, table, unique_zones, means, nodata_values, comm=comm)
            if zone_ids is not None:
                part2 = {n: a[pos] for n, a in part2.items()}
                means = means[pos]
            cols.update(finalize(part2, means, sv))
        else:
            cols.update(finalize(part, pivot, sv))
    if return_type == 'pandas.DataFrame':
        d = {"zone": sel}
        for s in names:
            d[s] = cols[s]
        return pd.DataFrame(d)
    out = _broadcast_back(cols, names, sel, zt, vt.shape, vt.device)
    return like_container(out, values)


def _broadcast_back(cols, names, sel, zt, shape, device):
    """(len(names), H, W) float64 tensor: every statistic broadcast onto its zone's cells, NaN
    elsewhere (zonal.py:313-331)."""
    import torch
    H, W = shape
    out = torch.full((len(names), H * W), float("nan"), dtype=torch.float64, device=device)
    if len(sel):
        ids_t = torch.as_tensor(np.asarray(sel, dtype=np.float64), device=device)
        zf = zt.reshape(-1).to(torch.float64)
        idx = torch.searchsorted(ids_t, zf).clamp_(max=len(sel) - 1)
        hit = ids_t[idx] == zf
        for i, s in enumerate(names):
            table = torch.as_tensor(np.asarray(cols[s], dtype=np.float64), device=device)
            out[i] = torch.where(hit, table[idx], out[i])
    return out.reshape(len(names), H, W)


def _stats_custom(zones, values, zone_ids, stats_funcs, nodata_values, return_type='pandas.DataFrame',
                  host=False):
    """`stats_funcs` given as a dict of callables (zonal.py:640-642): the reference groups the cells
    by zone with an argsort and calls every function on the zone's valid values
    (`_calc_stats`, zonal.py:144-163).  Same here, with the grouping done by one stable device
    sort; the callables receive the zone's values in the raster's dtype -- numpy arrays for numpy
    rasters (`host`), device tensors otherwise -- and must return a scalar."""
    import torch
    zt = _prepare(as_device_tensor(zones), (torch.int32, torch.int64, torch.float32, torch.float64))
    vt = as_device_tensor(values).contiguous()
    if len(vt.shape) > 2:
        raise TypeError('3D inputs not supported for the device backend')
    for name, f in stats_funcs.items():
        if not callable(f):
            raise ValueError(name)
    z = zt.reshape(-1)
    v = vt.reshape(-1)
    zfinite = torch.isfinite(z) if z.dtype.is_floating_point else None
    unique_zones = torch.unique(z[zfinite] if zfinite is not None else z).cpu().numpy()
    if zone_ids is None:
        sel = unique_zones
    else:
        sel = np.array([zz for zz in np.unique(zone_ids) if zz in unique_zones], dtype=unique_zones.dtype)
    ok = torch.isfinite(v) if v.dtype.is_floating_point else torch.ones_like(v, dtype=torch.bool)
    if nodata_values is not None:
        ok &= v != nodata_values
    if zfinite is not None:
        ok &= zfinite
    zk, vk = z[ok], v[ok]
    order = torch.argsort(zk, stable=True)
    zs, vs = zk[order], vk[order]
    present, counts = torch.unique_consecutive(zs, return_counts=True)
    present, counts = present.cpu().numpy(), counts.cpu().numpy()
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]]) if len(counts) else np.zeros(0, np.int64)
    where = {zz: (int(a), int(a + c)) for zz, a, c in zip(present.tolist(), starts, counts)}
    groups = vs.cpu().numpy() if host else vs
    cols = {}
    for name, func in stats_funcs.items():
        col = np.full(len(sel), np.nan)
        for i, zz in enumerate(sel.tolist()):
            if zz in where:
                a, b = where[zz]
                col[i] = float(func(groups[a:b]))
        cols[name] = col
    names = list(stats_funcs.keys())
    if return_type == 'pandas.DataFrame':
        d = {"zone": sel}
        d.update({n: cols[n] for n in names})
        return pd.DataFrame(d)
    out = _broadcast_back(cols, names, sel, zt, vt.shape, vt.device)
    return out.cpu().numpy() if host else like_container(out, values)


def _stats_host(zones, values, zone_ids, stats_funcs, nodata_values, return_type='pandas.DataFrame'):
    """numpy runner (replaces zonal.py:280 `_stats_numpy`): upload, same device pass."""
    import torch
    zt = torch.from_numpy(np.ascontiguousarray(zones)).cuda()
    vt = torch.from_numpy(np.ascontiguousarray(values)).cuda()
    if isinstance(stats_funcs, dict):
        res = _stats_custom(zt, vt, zone_ids, stats_funcs, nodata_values, return_type, host=True)
    else:
        res = _stats_device(zt, vt, zone_ids, stats_funcs, nodata_values, return_type)
        if return_type != 'pandas.DataFrame':
            res = res.cpu().numpy()
    if return_type == 'pandas.DataFrame':
        res["zone"] = res["zone"].astype(np.asarray(zones).dtype)
    return res


def stats(zones, values, zone_ids=None,
          stats_funcs=["mean", "max", "min", "sum", "std", "var", "count", "majority"],
          nodata_values=None, return_type='pandas.DataFrame', comm=None):
    """Per-zone summary statistics (zonal.py:422-667), same signature and defaults.

    `stats_funcs` as a list names built-in statistics: they come from one streaming pass over the
    raster (`majority` adds a second pass, so leave it out of the list when it is not needed).
    `stats_funcs` as a dict maps column names to callables, which are run per zone like the
    reference does (no built-in is ever substituted for a callable).  `comm` (optional
    torch.distributed group, not in the reference) marks `zones` and `values` as this rank's row
    stripe of a larger raster; custom callables are not available over stripes.
    """
    if isinstance(values, Dataset):
        if return_type != 'pandas.DataFrame':
            raise ValueError("return_type must be 'pandas.DataFrame' when values is a Dataset")
        dfs = []
        for var_name in values.data_vars:
            df = stats(zones, values[var_name], zone_ids, stats_funcs, nodata_values, 'pandas.DataFrame')
            df = df.rename(columns={c: f'{var_name}_{c}' for c in df.columns if c != 'zone'})
            dfs.append(df)
        result = dfs[0]
        for df in dfs[1:]:
            result = result.merge(df, on='zone', how='outer')
        return result

    validate_arrays(zones, values)
    for nm, arr in (("zones", zones), ("values", values)):
        dt = arr.data.dtype
        kind = getattr(dt, "kind", None)
        if kind is None:  # torch dtype
            ok = dt.is_floating_point or "int" in str(dt)
        else:
            ok = kind in "iuf"
        if not ok:
            raise ValueError("`%s` must be an array of integers or floats." % nm)

    if isinstance(stats_funcs, list):
        for s in stats_funcs:
            if s not in _DEFAULT_STATS:
                raise ValueError(f"Invalid stat name. {s} option not supported.")
        funcs = list(stats_funcs)
        names = funcs
    elif isinstance(stats_funcs, dict):
        funcs = dict(stats_funcs)
        names = list(funcs.keys())
        if comm is not None:
            raise NotImplementedError("custom statistics cannot be combined across row stripes; "
                                      "name built-in statistics in a list when `comm` is given")
    else:  # the reference falls through to an UnboundLocalError here; say what is wrong instead
        raise TypeError("`stats_funcs` must be a list of statistic names or a dict of callables")

    if comm is not None:
        result = _stats_device(zones.data, values.data, zone_ids, funcs, nodata_values, return_type, comm)
    else:
        device = _stats_custom if isinstance(funcs, dict) else _stats_device
        mapper = ArrayTypeFunctionMapping(
            numpy_func=lambda *a: _stats_host(*a, return_type=return_type),
            cupy_func=lambda *a: device(*a, return_type=return_type))
        result = mapper(values)(zones.data, values.data, zone_ids, funcs, nodata_values)

    if return_type == 'xarray.DataArray':
        coords = {'stats': names}
        coords.update(values.coords)
        return DataArray(result, coords=coords, dims=('stats',) + tuple(values.dims), attrs=values.attrs)
    return result


TOTAL_COUNT = '_total_count'


def _pivot_pairs(sel, cats, pz, pv, pc):
    """(zone, value, count) pairs -> per-zone totals (float32, like the reference) and a
    (len(cats), len(sel)) count matrix.  `cats` is sorted; zones outside `sel` are dropped.
    zonal.py:719-727: a selected category also collects the unselected categories between it and
    the previous selected one (the reference's `cat_start` only advances at selected categories);
    with all categories selected this is the plain per-category count."""
    total = np.zeros(len(sel), dtype=np.float32)
    counts = np.zeros((len(cats), len(sel)), dtype=np.int64)
    if len(sel) == 0 or len(pz) == 0:
        return total, counts
    sel_f = np.asarray(sel, dtype=np.float64)
    order = np.argsort(sel_f, kind="stable")
    pos = np.searchsorted(sel_f[order], pz.astype(np.float64), side="right") - 1   # last of equal ids
    pos[pos < 0] = 0
    hit = sel_f[order][pos] == pz
    row = order[pos]
    np.add.at(total, row[hit], pc[hit].astype(np.float32))      # unbuffered: pair order, like a loop
    if len(cats):
        bounds = np.asarray([float(c) for c in cats], dtype=np.float64)
        col = np.searchsorted(bounds, pv, side="left")
        keep = hit & (col < len(cats))
        np.add.at(counts, (col[keep], row[keep]), pc[keep])
    return total, counts


def _crosstab_3d(zones, values, zone_ids, cat_ids, layer, agg, nodata_values, comm):
    """3-D `values` (zonal.py:1096-1116, `_single_zone_crosstab_3d` :734-745): the categories are the
    coordinate values of dimension `layer`, and cell (zone, category) is statistic `agg` of that
    category's 2-D layer over the zone -- i.e. one zonal.stats pass per selected layer."""
    import torch
    if agg not in _DEFAULT_STATS:
        raise ValueError("`agg` method for 3D numpy backed data array must be one of following %s"
                         % (list(_DEFAULT_STATS),))
    if layer is None:
        layer = 0
    try:
        ldim = values.dims[layer]
        if ldim not in values.coords:
            raise KeyError(ldim)
        unique_cats = np.asarray(getattr(values.coords[ldim], "values", values.coords[ldim]))
    except (IndexError, KeyError, TypeError):
        raise ValueError("Invalid `layer`")
    axis = list(values.dims).index(ldim)
    host = isinstance(values.data, np.ndarray)
    vt = torch.from_numpy(np.ascontiguousarray(values.data)).cuda() if host else as_device_tensor(values.data)
    zt = torch.from_numpy(np.ascontiguousarray(zones.data)).cuda() if isinstance(zones.data, np.ndarray) \
        else as_device_tensor(zones.data)
    vt = torch.movedim(vt, axis, 0)
    if tuple(zt.shape) != tuple(vt.shape[1:]):
        raise ValueError("Incompatible shapes")
    zt = _prepare(zt, (torch.int32, torch.int64, torch.float32, torch.float64))
    if cat_ids is None:
        cats = list(unique_cats.tolist())
    else:
        cats = [c for c in cat_ids if c in unique_cats]
    cat_pos = {c: j for j, c in enumerate(unique_cats.tolist())}
    zf = zt.reshape(-1)
    zfin = zf[torch.isfinite(zf)] if zf.dtype.is_floating_point else zf
    unique_zones = torch.unique(zfin).cpu().numpy()
    if zone_ids is None:
        sel = unique_zones
    else:
        sel = np.array([z for z in zone_ids if z in unique_zones], dtype=unique_zones.dtype)
    d = {"zone": sel}
    for c in cats:
        lt = vt[cat_pos[c]].contiguous()
        lt = lt if lt.dtype in (torch.float32, torch.float64) else lt.to(torch.float64)
        table = {}
        ids, part, pivot0 = hash_partials(zt, lt, nodata_values, comm=comm, table=table)
        pos = np.searchsorted(ids, sel)
        pos = np.clip(pos, 0, max(len(ids) - 1, 0))
        hit = (ids[pos] == sel) if len(ids) else np.zeros(len(sel), bool)
        if agg == "majority":
            col_all = majority_by_zone(zt, lt, ids, nodata_values, comm=comm)
        elif agg in ("std", "var") and lt.dtype == torch.float64 and len(ids):
            cnt = part["count"].astype(np.float64)
            with np.errstate(invalid="ignore", divide="ignore"):
                means = np.where(cnt > 0, pivot0 + part["s1"] / cnt, 0.0)
            part2 = second_pass_partials(zt, lt, table, ids, means, nodata_values, comm=comm)
            col_all = finalize(part2, means, [agg])[agg]
        else:
            col_all = finalize(part, np.full(len(ids), pivot0), [agg])[agg]
        col = np.where(hit, col_all[pos] if len(ids) else np.nan, np.nan)
        if agg == "count":          # np.ma.count of an empty selection is 0, and the column is integer
            col = np.where(np.isnan(col), 0, col).astype(np.int64)
        d[c] = col
    return pd.DataFrame(d)


def crosstab(zones, values, zone_ids=None, cat_ids=None, layer=None, agg="count", nodata_values=None, comm=None):
    """Cross-tabulation of a `values` raster by zone (zonal.py:922-1155): a `zone` column and one column
    per category.  2-D values: cell counts (or percentages) of the categorical values per zone, built on
    the (zone, value) pair histogram of xrs_zonal_pair_count.  3-D values: statistic `agg` of every
    category layer per zone (`layer` names the category dimension)."""
    if not isinstance(zones, DataArray):
        raise TypeError("zones must be instance of DataArray")
    if not isinstance(values, DataArray):
        raise TypeError("values must be instance of DataArray")
    if zones.ndim != 2:
        raise ValueError("zones must be 2D")
    if values.ndim not in (2, 3):
        raise ValueError("`values` must use either 2D or 3D coordinates.")
    if values.ndim == 3:
        return _crosstab_3d(zones, values, zone_ids, cat_ids, layer, agg, nodata_values, comm)
    validate_arrays(zones, values)
    if agg not in ("percentage", "count"):
        raise ValueError("`agg` method for 2D data array must be one of following ['percentage', 'count']")
    import torch
    if isinstance(values.data, np.ndarray):
        zt = torch.from_numpy(np.ascontiguousarray(zones.data)).cuda()
        vt = torch.from_numpy(np.ascontiguousarray(values.data)).cuda()
    else:
        zt, vt = as_device_tensor(zones.data), as_device_tensor(values.data)
    zt = _prepare(zt, (torch.int32, torch.int64, torch.float32, torch.float64))
    vt_f = vt.contiguous() if vt.dtype in (torch.float32, torch.float64) else vt.to(torch.float64)
    unique_zones, _, _ = hash_partials(zt, vt_f, None, comm=comm)
    try:
        pz, pv, pc = pair_counts(zt, vt_f, nodata_values, comm=comm)
    except _PairTableOverflow:
        raise NotImplementedError("crosstab: more than 64M distinct (zone, category) pairs")
    vdtype = np.dtype(str(vt.dtype).replace("torch.", "")) if not isinstance(values.data, np.ndarray) else values.data.dtype
    unique_cats = np.unique(pv).astype(vdtype)
    if cat_ids is None:
        cats = unique_cats
    else:
        cats = [c for c in cat_ids if c in unique_cats]
    if zone_ids is None:
        sel = unique_zones
    else:
        sel = np.array([z for z in zone_ids if z in unique_zones], dtype=unique_zones.dtype)
    cats = sorted(cats)
    total, counts = _pivot_pairs(sel, cats, pz, pv, pc)
    table = {c: counts[j] for j, c in enumerate(cats)}
    d = {"zone": sel}
    if agg == "percentage":
        total[total == 0] = np.nan
        for c in cats:
            d[c] = table[c] / total * 100
    else:
        for c in cats:
            d[c] = table[c]
    return pd.DataFrame(d)
