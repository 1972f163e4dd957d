"""xarray access point.

The package is written against the xarray API (DataArray / Dataset / concat).  When the real
`xarray` is importable it is used unchanged; this container and the GPU boxes do not ship it
(SURVEY.md section 0 fact 7), so a minimal stand-in with the handful of members the hot path
touches is provided.  It is NOT a general xarray replacement.
"""
from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - depends on the environment
    import xarray as _xarray
    DataArray = _xarray.DataArray
    Dataset = _xarray.Dataset
    concat = _xarray.concat
    HAVE_XARRAY = True
except ImportError:
    HAVE_XARRAY = False

    def _shape_of(data):
        return tuple(data.shape)

    class DataArray(object):
        """Minimal stand-in for xarray.DataArray (data + dims + coords + attrs + name)."""

        def __init__(self, data=None, coords=None, dims=None, name=None, attrs=None):
            if isinstance(data, DataArray):
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
            self.dims = tuple(dims)
            if len(self.dims) != nd:
                raise ValueError("different number of dimensions on data and dims: %d vs %d"
                                 % (nd, len(self.dims)))
            self.coords = OrderedDict()
            if coords is not None:
                items = coords.items() if hasattr(coords, "items") else zip(self.dims, coords)
                for k, v in items:
                    self.coords[k] = v.data if isinstance(v, DataArray) else np.asarray(v)
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
                    return DataArray(np.asarray(self.coords[key]), dims=(key,), name=key)
                raise KeyError(key)
            raise NotImplementedError("positional indexing is not part of the shim")

        def __setitem__(self, key, value):
            if not isinstance(key, str):
                raise NotImplementedError("only coordinate assignment is supported")
            self.coords[key] = value.data if isinstance(value, DataArray) else np.asarray(value)

        def min(self):
            return DataArray(np.asarray(np.nanmin(self.values)))

        def max(self):
            return DataArray(np.asarray(np.nanmax(self.values)))

        def item(self):
            return self.values.item()

        def copy(self, deep=True):
            d = self.data
            if deep:
                d = d.clone() if hasattr(d, "clone") else d.copy()
            return DataArray(d, coords=self.coords, dims=self.dims, name=self.name, attrs=self.attrs)

        def __repr__(self):
            return "<shim.DataArray %r %s %s>\n%r" % (self.name, dict(zip(self.dims, self.shape)),
                                                     self.dtype, self.data)

    class Dataset(object):
        """Minimal stand-in for xarray.Dataset: an ordered mapping name -> DataArray."""

        def __init__(self, data_vars=None, coords=None, attrs=None):
            self.data_vars = OrderedDict()
            for k, v in (data_vars or {}).items():
                self.data_vars[k] = v if isinstance(v, DataArray) else DataArray(v)
            self.coords = OrderedDict(coords or {})
            self.attrs = dict(attrs) if attrs is not None else {}

        def __getitem__(self, key):
            return self.data_vars[key]

        def __contains__(self, key):
            return key in self.data_vars

        def __iter__(self):
            return iter(self.data_vars)

        def __len__(self):
            return len(self.data_vars)

        def keys(self):
            return self.data_vars.keys()

    def concat(objs, dim):
        """Stack 2-D DataArrays along a new leading dimension described by a pandas Index."""
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
        return DataArray(data, coords=coords, dims=(name,) + tuple(first.dims), attrs=first.attrs,
                         name=first.name)
