"""Dataset fan-out decorators with the reference's contract (dataset_support.py:11-80)."""
import functools
import inspect

from ._xr import Dataset


def supports_dataset(func):
    """Single-input DataArray function -> also accepts a Dataset (applied per data variable;
    `name=<var>` is injected when the function has a `name` parameter)."""
    has_name_param = 'name' in inspect.signature(func).parameters

    @functools.wraps(func)
    def wrapper(agg, *args, **kwargs):
        if isinstance(agg, Dataset):
            results = {}
            for var_name in agg.data_vars:
                kw = dict(kwargs)
                if has_name_param:
                    kw['name'] = var_name
                results[var_name] = func(agg[var_name], *args, **kw)
            return Dataset(results, attrs=agg.attrs)
        return func(agg, *args, **kwargs)

    return wrapper


def supports_dataset_bands(**band_param_map):
    """Multi-band function -> also accepts `f(ds, nir='B8', red='B4', ...)`."""

    def decorator(func):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            if args and isinstance(args[0], Dataset):
                ds = args[0]
                func_kwargs = {}
                used = set()
                for alias, param in band_param_map.items():
                    if alias not in kwargs:
                        raise TypeError(f"'{alias}' keyword required when passing a Dataset")
                    var_name = kwargs[alias]
                    if var_name not in ds.data_vars:
                        raise ValueError(f"'{var_name}' not in Dataset. Available: {list(ds.data_vars)}")
                    func_kwargs[param] = ds[var_name]
                    used.add(alias)
                for k, v in kwargs.items():
                    if k not in used:
                        func_kwargs[k] = v
                return func(**func_kwargs)
            return func(*args, **kwargs)

        return wrapper

    return decorator
