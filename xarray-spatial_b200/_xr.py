"""xarray access point.

The package is written against the xarray API (DataArray / Dataset / concat).  Two container
families can be live at once:

* the real `xarray` classes, used whenever xarray is importable AND the payload is something
  xarray can hold (numpy arrays, lists, scalars): a numpy-backed `xarray.DataArray` in gives a
  numpy-backed `xarray.DataArray` out, exactly like the reference;
* a small stand-in (`ShimDataArray` / `ShimDataset`) for device payloads -- `torch.Tensor` has
  neither `__array_function__` nor `__array_namespace__`, so `xarray.DataArray(cuda tensor)` would
  call `np.asarray` on it and fail -- and for environments without xarray (this container and
  the GPU boxes do not ship it).

`DataArray`, `Dataset` and `concat` below are facades: calling them builds whichever family fits
the payload, and `isinstance(x, DataArray)` is true for both families, so operator code never
needs to know which one it holds.  The stand-in is NOT a general xarray replacement, but where
it implements a member it is as strict as xarray (read-only `data_vars`, default integer index
for a dimension without coordinate, dimension-count checks).
"""
from collections import OrderedDict
from types import MappingProxyType

import numpy as np

try:  # pragma: no cover - depends on the environment
    import xarray as _xarray
    HAVE_XARRAY = True
except ImportError:
    _xarray = None
    HAVE_XARRAY = False


def _shape_of(data):
    return tuple(data.shape)


def is_torch_tensor(x):
    return type(x).__module__.split(".")[0] == "torch" and hasattr(x, "data_ptr")


def _unwrap(v):
    """Payload of a coordinate / variable given as any DataArray-like."""
    return v.data if hasattr(v, "dims") and hasattr(v, "data") else v


class ShimDataArray(object):
    """Minimal stand-in for xarray.DataArray (data + dims + coords + attrs + name)."""

    def __init__(self, data=None, coords=None, dims=None, name=None, attrs=None):
        if hasattr(data, "dims") and hasattr(data, "coords"):   # any DataArray-like
            coords = data.coords if coords is None else coords
            dims = data.dims if dims is None else dims
            attrs = data.attrs if attrs is None else attrs
            name = data.name if name is None else name
            data = data.data
        if isinstance(data, (list, tuple)):
            data = np.asarray(data)
        self.data = data
        nd = len(_shape_of(data))
        if dims is None:
            dims = tuple("dim_%d" % i for i in range(nd))
        if isinstance(dims, str):
            dims = (dims,)
        self.dims = tuple(dims)
        if len(self.dims) != nd:
            raise ValueError("different number of dimensions on data and dims: %d vs %d"
                             % (nd, len(self.dims)))
        self.coords = OrderedDict()
        if coords is not None:
            items = coords.items() if hasattr(coords, "items") else zip(self.dims, coords)
            for k, v in items:
                v = _unwrap(v)
                v = v if is_torch_tensor(v) else np.asarray(v)
                if k in self.dims and len(_shape_of(v)) == 1 and \
                        _shape_of(v)[0] != self.shape[self.dims.index(k)]:
                    raise ValueError("conflicting sizes for dimension %r: length %d on the data but length "
                                     "%d on coordinate %r" % (k, self.shape[self.dims.index(k)],
                                                              _shape_of(v)[0], k))
                self.coords[k] = v
        self.attrs = dict(attrs) if attrs is not None else {}
        self.name = name

    # --- array protocol ---------------------------------------------------------
    @property
    def shape(self):
        return _shape_of(self.data)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def values(self):
        d = self.data
        if isinstance(d, np.ndarray):
            return d
        if hasattr(d, "detach"):  # torch tensor
            return d.detach().cpu().numpy()
        return np.asarray(d)

    def __array__(self, dtype=None, copy=None):
        v = self.values
        return v.astype(dtype) if dtype is not None else v

    # --- coordinates ------------------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, str):
            if key in self.coords:
                c = self.coords[key]
                cd = (key,) if len(_shape_of(c)) == 1 else tuple(self.dims[-len(_shape_of(c)):]) if len(_shape_of(c)) else ()
                return ShimDataArray(c, dims=cd, name=key)
            if key in self.dims:  # xarray: a dimension without coordinate has the default integer index
                return ShimDataArray(np.arange(self.shape[self.dims.index(key)]), dims=(key,), name=key)
            raise KeyError(key)
        raise NotImplementedError("positional indexing is not part of the shim")

    def __setitem__(self, key, value):
        if not isinstance(key, str):
            raise NotImplementedError("only coordinate assignment is supported")
        value = _unwrap(value)
        self.coords[key] = value if is_torch_tensor(value) else np.asarray(value)

    def min(self):
        return ShimDataArray(np.asarray(np.nanmin(self.values)))

    def max(self):
        return ShimDataArray(np.asarray(np.nanmax(self.values)))

    def item(self):
        return self.values.item()

    def copy(self, deep=True):
        d = self.data
        if deep:
            d = d.clone() if hasattr(d, "clone") else d.copy()
        return ShimDataArray(d, coords=self.coords, dims=self.dims, name=self.name, attrs=self.attrs)

    def __repr__(self):
        return "<shim.DataArray %r %s %s>\n%r" % (self.name, dict(zip(self.dims, self.shape)),
                                                 self.dtype, self.data)


class ShimDataset(object):
    """Minimal stand-in for xarray.Dataset: an ordered mapping name -> DataArray.  Like xarray,
    `data_vars` is a read-only view; variables are added with `ds[name] = array`."""

    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._vars = OrderedDict()
        for k, v in (data_vars or {}).items():
            self[k] = v
        self.coords = OrderedDict(coords or {})
        self.attrs = dict(attrs) if attrs is not None else {}

    @property
    def data_vars(self):
        return MappingProxyType(self._vars)

    def __getitem__(self, key):
        return self._vars[key]

    def __setitem__(self, key, value):
        if not isinstance(key, str):
            raise NotImplementedError("only `ds[name] = array` is supported")
        if not (hasattr(value, "dims") and hasattr(value, "data")):
            value = ShimDataArray(value)
        for other in self._vars.values():   # xarray refuses conflicting dimension sizes
            for d, n in zip(value.dims, value.shape):
                if d in other.dims and other.shape[other.dims.index(d)] != n:
                    raise ValueError("conflicting sizes for dimension %r" % (d,))
        self._vars[key] = value

    def __contains__(self, key):
        return key in self._vars

    def __iter__(self):
        return iter(self._vars)

    def __len__(self):
        return len(self._vars)

    def keys(self):
        return self._vars.keys()


def _holds_device_data(x):
    return is_torch_tensor(_unwrap(x))


class _DataArrayFacade(type):
    def __instancecheck__(cls, obj):
        return isinstance(obj, ShimDataArray) or (HAVE_XARRAY and isinstance(obj, _xarray.DataArray))

    def __call__(cls, data=None, coords=None, dims=None, name=None, attrs=None):
        if HAVE_XARRAY and not _holds_device_data(data) and \
                not any(is_torch_tensor(_unwrap(v)) for v in (coords.values() if hasattr(coords, "values") else ())):
            kw = {}
            if coords is not None:
                kw["coords"] = coords
            if dims is not None:
                kw["dims"] = dims
            return _xarray.DataArray(data, name=name, attrs=attrs, **kw)
        return ShimDataArray(data, coords=coords, dims=dims, name=name, attrs=attrs)


class DataArray(metaclass=_DataArrayFacade):
    """`DataArray(...)` builds an xarray.DataArray when xarray is importable and the payload is
    host data, a ShimDataArray otherwise; isinstance() accepts both."""


class _DatasetFacade(type):
    def __instancecheck__(cls, obj):
        return isinstance(obj, ShimDataset) or (HAVE_XARRAY and isinstance(obj, _xarray.Dataset))

    def __call__(cls, data_vars=None, coords=None, attrs=None):
        vals = list((data_vars or {}).values())
        if HAVE_XARRAY and not any(_holds_device_data(v) or isinstance(v, ShimDataArray) for v in vals):
            return _xarray.Dataset(data_vars, coords=coords, attrs=attrs)
        return ShimDataset(data_vars, coords=coords, attrs=attrs)


class Dataset(metaclass=_DatasetFacade):
    """`Dataset(...)` builds an xarray.Dataset or a ShimDataset (see DataArray)."""


def dataset_like(template, data_vars, attrs=None):
    """A Dataset able to hold `data_vars`: real xarray only if every variable is one."""
    return Dataset(data_vars, attrs=attrs)


def concat(objs, dim):
    """Stack 2-D DataArrays along a new leading dimension described by a pandas Index."""
    if HAVE_XARRAY and all(isinstance(o, _xarray.DataArray) for o in objs):
        return _xarray.concat(objs, dim)
    name = getattr(dim, "name", None) or "concat_dim"
    first = objs[0]
    datas = [o.data for o in objs]
    if isinstance(datas[0], np.ndarray):
        data = np.stack(datas, axis=0)
    else:
        import torch
        data = torch.stack(datas, dim=0)
    coords = OrderedDict()
    coords[name] = np.asarray(list(dim), dtype=object)
    coords.update(first.coords)
    return ShimDataArray(data, coords=coords, dims=(name,) + tuple(first.dims), attrs=first.attrs,
                         name=first.name)
