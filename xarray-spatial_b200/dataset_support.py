"""Dataset fan-out for DataArray operators.

Behavioural contract of the reference's decorators (xrspatial/dataset_support.py:11-80):

* ``supports_dataset``: a one-raster operator called with a Dataset runs once per data variable;
  the variable's name is passed as ``name=`` when the operator takes such a parameter; the
  results come back as a Dataset carrying the input's attrs.
* ``supports_dataset_bands(alias=parameter, ...)``: a multi-band operator called as
  ``op(ds, alias='variable name', ..., other=kwargs)`` looks the bands up in the Dataset;
  a missing alias is a TypeError, an unknown variable a ValueError.
"""
import functools
import inspect

from ._xr import Dataset


class _PerVariable(object):
    """Callable wrapper applying a single-raster operator to every variable of a Dataset."""

    def __init__(self, op):
        self.op = op
        self.passes_name = 'name' in inspect.signature(op).parameters
        functools.update_wrapper(self, op)

    def _one(self, ds, var, args, kwargs):
        extra = dict(kwargs, name=var) if self.passes_name else kwargs
        return self.op(ds[var], *args, **extra)

    def __call__(self, agg, *args, **kwargs):
        if not isinstance(agg, Dataset):
            return self.op(agg, *args, **kwargs)
        return Dataset({var: self._one(agg, var, args, kwargs) for var in agg.data_vars}, attrs=agg.attrs)


def supports_dataset(op):
    wrapper = _PerVariable(op)

    @functools.wraps(op)
    def call(agg, *args, **kwargs):
        return wrapper(agg, *args, **kwargs)

    return call


def _resolve_bands(ds, aliases, kwargs):
    """Translate ``alias='variable'`` keywords into ``parameter=DataArray`` keywords."""
    resolved = {k: v for k, v in kwargs.items() if k not in aliases}
    for alias, parameter in aliases.items():
        try:
            variable = kwargs[alias]
        except KeyError:
            raise TypeError(f"'{alias}' keyword required when passing a Dataset") from None
        if variable not in ds.data_vars:
            raise ValueError(f"'{variable}' not in Dataset. Available: {list(ds.data_vars)}")
        resolved[parameter] = ds[variable]
    return resolved


def supports_dataset_bands(**aliases):
    def decorate(op):
        @functools.wraps(op)
        def call(*args, **kwargs):
            if args and isinstance(args[0], Dataset):
                return op(**_resolve_bands(args[0], aliases, kwargs))
            return op(*args, **kwargs)

        return call

    return decorate
