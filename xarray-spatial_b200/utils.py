"""Backend dispatch, resolution and validation helpers -- the L2/L3 layer of the reference
(xrspatial/utils.py) re-stated for the B200 backend.

Array kinds understood by this package
  * numpy.ndarray            -> "host" runners (xrs_host_* C-ABI, pipelined H2D/compute/D2H)
  * torch CUDA tensor, or any object exposing __cuda_array_interface__ (e.g. cupy.ndarray)
                             -> "device" runners (device-pointer C-ABI on the current stream)
Dask arrays are not supported (the north star excludes Dask): NotImplementedError.
"""
import ctypes
import os

import numpy as np

from . import _lib

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None


def ngjit(f):  # name kept for signature parity; nothing is JIT-compiled in this package
    return f


# ----------------------------------------------------------------------------- array kinds
def is_torch_cuda(a):
    return torch is not None and isinstance(a, torch.Tensor) and a.is_cuda


def is_device_array(a):
    return is_torch_cuda(a) or (hasattr(a, "__cuda_array_interface__") and not isinstance(a, np.ndarray))


def is_dask_array(a):
    return type(a).__module__.split(".")[0] == "dask"


def as_device_tensor(a):
    """View a device array as a torch CUDA tensor (zero copy)."""
    if is_torch_cuda(a):
        return a
    if hasattr(a, "__cuda_array_interface__"):
        return torch.as_tensor(a, device="cuda")
    raise TypeError("not a device array: %r" % type(a))


def like_container(result, template):
    """Return `result` (torch CUDA tensor) in the container type of `template`."""
    if is_torch_cuda(template) or template is None:
        return result
    mod = type(template).__module__.split(".")[0]
    if mod == "cupy":  # pragma: no cover - cupy is optional
        import cupy
        return cupy.asarray(result)
    return result


def stream_ptr(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def host_device_index():
    return int(os.environ.get("XRS_B200_DEVICE", "0"))


def host_devices():
    """CUDA ordinals the host-buffer (numpy) path stripes a raster over.

    XRS_B200_DEVICES = "all" | comma-separated ordinals.  Default: every visible GPU -- one PCIe link
    per stripe -- unless this process is one rank of a one-process-per-GPU job (LOCAL_RANK set, e.g.
    under torchrun) or XRS_B200_DEVICE pins a device, in which case only that GPU is used."""
    spec = os.environ.get("XRS_B200_DEVICES")
    if spec is None:
        if "XRS_B200_DEVICE" in os.environ:
            return [host_device_index()]
        if "LOCAL_RANK" in os.environ:
            return [torch.cuda.current_device() if torch is not None and torch.cuda.is_available()
                    else int(os.environ["LOCAL_RANK"])]
        spec = "all"
    if spec.strip().lower() == "all":
        n = ctypes.c_int(0)
        _lib.call("xrs_device_count", ctypes.byref(n))
        return list(range(max(1, n.value)))
    devs = [int(x) for x in spec.split(",") if x.strip() != ""]
    if not devs or len(set(devs)) != len(devs):
        raise ValueError("XRS_B200_DEVICES must list distinct CUDA ordinals, got %r" % spec)
    return devs


def _dev_array(devs):
    return (ctypes.c_int * len(devs))(*devs), len(devs)


def not_implemented_func(agg, *args, messages='Not yet implemented.'):
    raise NotImplementedError(messages)


class ArrayTypeFunctionMapping(object):
    """utils.py:117-143 with the same constructor; the `cupy_func` slot is the B200 device
    runner and `numpy_func` the B200 host-buffer runner."""

    def __init__(self, numpy_func, cupy_func, dask_func=None, dask_cupy_func=None):
        self.numpy_func = numpy_func
        self.cupy_func = cupy_func
        self.dask_func = dask_func
        self.dask_cupy_func = dask_cupy_func

    def __call__(self, arr):
        data = arr.data
        if isinstance(data, np.ndarray):
            return self.numpy_func
        if is_device_array(data):
            return self.cupy_func
        if is_dask_array(data):
            if self.dask_func is not None:
                return self.dask_func
            raise NotImplementedError("dask-backed DataArrays are not supported by the B200 backend; "
                                      "row-stripe large rasters with xrspatial_b200.stripes instead")
        raise TypeError("Unsupported Array Type: {}".format(type(arr)))


def validate_arrays(*arrays):
    """utils.py:146-165."""
    if len(arrays) < 2:
        raise ValueError("validate_arrays() input must contain 2 or more arrays")
    first = arrays[0]
    for other in arrays[1:]:
        if not tuple(first.data.shape) == tuple(other.data.shape):
            raise ValueError("input arrays must have equal shapes")
        if not isinstance(first.data, type(other.data)):
            raise ValueError("input arrays must have same type")


# ----------------------------------------------------------------------------- resolution
def get_xy_range(raster, xdim=None, ydim=None):
    if ydim is None:
        ydim = raster.dims[-2]
    if xdim is None:
        xdim = raster.dims[-1]
    xmin = raster[xdim].min().item()
    xmax = raster[xdim].max().item()
    ymin = raster[ydim].min().item()
    ymax = raster[ydim].max().item()
    return (xmin, xmax), (ymin, ymax)


def calc_res(raster, xdim=None, ydim=None):
    """utils.py:204-230."""
    h, w = raster.shape[-2:]
    xrange, yrange = get_xy_range(raster, xdim, ydim)
    xres = (xrange[-1] - xrange[0]) / (w - 1)
    yres = (yrange[-1] - yrange[0]) / (h - 1)
    return xres, yres


def get_dataarray_resolution(agg, xdim=None, ydim=None):
    """utils.py:233-277: attrs['res'] (2-sequence or scalar) else coordinates."""
    try:
        cellsize = agg.attrs.get("res")
        if (isinstance(cellsize, (tuple, np.ndarray, list)) and len(cellsize) == 2
                and isinstance(cellsize[0], (int, float)) and isinstance(cellsize[1], (int, float))):
            cellsize_x, cellsize_y = cellsize
        elif isinstance(cellsize, (int, float)):
            cellsize_x = cellsize
            cellsize_y = cellsize
        else:
            cellsize_x, cellsize_y = calc_res(agg, xdim, ydim)
    except Exception:
        cellsize_x, cellsize_y = calc_res(agg, xdim, ydim)
    return cellsize_x, cellsize_y


# ----------------------------------------------------------------------------- runner helpers
def _dbl_array(vals):
    arr = (ctypes.c_double * max(1, len(vals)))(*[float(v) for v in vals])
    return arr


def device_f32_2d(data):
    """(tensor, template): 2-D float32 CUDA tensor with unit inner stride (cast/copy only if
    needed, like the reference's `data.astype(cupy.float32)`, slope.py:150)."""
    t = as_device_tensor(data)
    if t.dim() != 2:
        raise ValueError("expected a 2-D raster, got %d-D" % t.dim())
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if t.numel() and t.stride(1) != 1:
        t = t.contiguous()
    return t


def run_stencil_device(fn_name, data, *scalars, aux=None, naux=None, extra_ints=(), dtype=None):
    """Call a `(in, pitch, out, pitch, H, W, ...)` device entry point on torch's current stream."""
    if dtype is None:
        t = device_f32_2d(data)
    else:
        t = as_device_tensor(data)
        if t.dtype != dtype:
            t = t.to(dtype)
        if t.numel() and t.stride(1) != 1:
            t = t.contiguous()
    H, W = t.shape
    out = torch.empty((H, W), dtype=t.dtype, device=t.device)
    if H == 0 or W == 0:
        return like_container(out, data)
    esz = t.element_size()
    args = [ctypes.c_void_p(t.data_ptr()), t.stride(0) * esz, ctypes.c_void_p(out.data_ptr()),
            out.stride(0) * esz, H, W]
    args += [float(s) for s in scalars]
    if aux is not None:
        args += [aux]
        if naux is not None:
            args += [int(naux)]
    args += [int(i) for i in extra_ints]
    with torch.cuda.device(t.device):
        args.append(stream_ptr(t))
        _lib.call(fn_name, *args)
    return like_container(out, data)


_INGEST_NP = {"int16": 4, "uint16": 5, "int32": 2, "float64": 1}


def run_surface_device(op, fn_name, data, *scalars):
    """slope / aspect / curvature / hillshade on a device raster.  int16 / uint16 / int32 / float64
    rasters go through the direct-ingest kernels (xrs_surface_typed: no separate cast pass); other
    dtypes are cast to float32 first, like the reference does (slope.py:150)."""
    t = as_device_tensor(data)
    code = _INGEST_NP.get(str(t.dtype).replace("torch.", ""))
    if code is not None and t.dim() == 2 and t.numel() and t.stride(1) == 1:
        H, W = t.shape
        out = torch.empty((H, W), dtype=torch.float32, device=t.device)
        try:
            with torch.cuda.device(t.device):
                _lib.call("xrs_surface_typed", _lib.OPS[op], ctypes.c_void_p(t.data_ptr()), code,
                          t.stride(0) * t.element_size(), ctypes.c_void_p(out.data_ptr()), out.stride(0) * 4, H, W,
                          _dbl_array(scalars), stream_ptr(t))
            return like_container(out, data)
        except NotImplementedError:
            pass  # layout the ingest kernels do not take (odd width, misaligned rows): cast instead
    return run_stencil_device(fn_name, data, *scalars)


def run_surface_host(op, data, p=()):
    """Same for numpy rasters: raw 16-bit / int32 / float64 cells cross PCIe and are converted on the
    device when the layout allows it, else the raster is cast on the host (slope.py:58)."""
    from . import _hostmem
    code = _INGEST_NP.get(data.dtype.name)
    if code is not None and data.ndim == 2 and data.size and data.shape[1] % 4 == 0:
        d = np.ascontiguousarray(data)
        H, W = d.shape
        out = _hostmem.empty((H, W), np.float32)
        try:
            devs, nd = _dev_array(host_devices())
            _lib.call("xrs_host_surface_typed_multi", _lib.OPS[op], ctypes.c_void_p(d.ctypes.data), code,
                      ctypes.c_void_p(out.ctypes.data), H, W, _dbl_array(p), devs, nd)
            return out
        except NotImplementedError:
            pass
    return run_stencil_host(op, data, p)


def run_stencil_host(op, data, p=(), aux=(), out_dtype=np.float32, in_dtype=np.float32):
    """Call xrs_host_stencil on a numpy raster; the result is a numpy array in pinned memory."""
    from . import _hostmem
    if data.ndim != 2:
        raise ValueError("expected a 2-D raster, got %d-D" % data.ndim)
    d = np.ascontiguousarray(data, dtype=in_dtype)
    H, W = d.shape
    if H == 0 or W == 0:
        return np.empty((H, W), out_dtype)
    out = _hostmem.empty((H, W), out_dtype)
    devs, nd = _dev_array(host_devices())
    _lib.call("xrs_host_stencil_multi", _lib.OPS[op], ctypes.c_void_p(d.ctypes.data),
              ctypes.c_void_p(out.ctypes.data), H, W, _dbl_array(p), _dbl_array(aux), len(aux), devs, nd)
    return out


# ----------------------------------------------------------------------------- geodesic helpers
# z-unit factors and lat/lon extraction for method='geodesic' (utils.py:593-713)
Z_UNITS = {
    'meter': 1.0, 'meters': 1.0, 'm': 1.0,
    'foot': 0.3048, 'feet': 0.3048, 'ft': 0.3048,
    'kilometer': 1000.0, 'kilometers': 1000.0, 'km': 1000.0,
    'mile': 1609.344, 'miles': 1609.344, 'mi': 1609.344,
}
_LAT_NAMES = {'lat', 'latitude', 'y'}
_LON_NAMES = {'lon', 'longitude', 'x'}


def _coord_values(agg, name):
    c = agg.coords[name]
    return np.asarray(getattr(c, "values", c))


def _find_coord(agg, dim_name, known_names, label):
    """A numeric coordinate named like the dimension, else any coordinate with a known name."""
    if dim_name in agg.coords and np.issubdtype(_coord_values(agg, dim_name).dtype, np.number):
        return _coord_values(agg, dim_name)
    for name in agg.coords:
        if str(name).lower() in known_names and np.issubdtype(_coord_values(agg, name).dtype, np.number):
            return _coord_values(agg, name)
    raise ValueError(
        f"geodesic method requires {label} coordinates on the DataArray. "
        f"No numeric coordinate found for dim '{dim_name}' or any of {sorted(known_names)}.")


def _validate_geographic_range(lat, lon):
    lat_min, lat_max = np.nanmin(lat), np.nanmax(lat)
    lon_min, lon_max = np.nanmin(lon), np.nanmax(lon)
    if lat_min < -90 or lat_max > 90:
        raise ValueError(f"Latitude values must be in [-90, 90], got [{lat_min}, {lat_max}]. "
                         f"Are your coordinates in a projected CRS?")
    if lon_min < -180 or lon_max > 360:
        raise ValueError(f"Longitude values must be in [-180, 360], got [{lon_min}, {lon_max}]. "
                         f"Are your coordinates in a projected CRS?")
    if lat_max - lat_min > 180 or lon_max - lon_min > 360:
        raise ValueError(f"Coordinate span too large for geographic coordinates "
                         f"(lat span={lat_max - lat_min}, lon span={lon_max - lon_min}). "
                         f"Are your coordinates in a projected CRS?")


def extract_latlon(agg):
    """(lat, lon, is_2d): float64 latitude / longitude of the raster's cells -- 1-D per-row /
    per-column vectors for a regular grid, (H, W) arrays for a curvilinear one (utils.py:608-662,
    without materialising the broadcast for regular grids)."""
    if agg.ndim < 2:
        raise ValueError(f"geodesic method requires a 2-D DataArray, got {agg.ndim}-D")
    dim_y, dim_x = agg.dims[-2], agg.dims[-1]
    lat = np.asarray(_find_coord(agg, dim_y, _LAT_NAMES, 'latitude'), dtype=np.float64)
    lon = np.asarray(_find_coord(agg, dim_x, _LON_NAMES, 'longitude'), dtype=np.float64)
    if lat.ndim == 1 and lon.ndim == 1:
        is_2d = False
    elif lat.ndim == 2 and lon.ndim == 2:
        is_2d = True
    else:
        raise ValueError(f"lat/lon coordinates must be both 1-D or both 2-D, got lat={lat.ndim}-D and lon={lon.ndim}-D")
    _validate_geographic_range(lat, lon)
    return np.ascontiguousarray(lat), np.ascontiguousarray(lon), is_2d


def run_geodesic(data, lat, lon, is_2d, z_factor, want_aspect):
    """xrs_geodesic on a device or host raster; result in the container type of `data`."""
    host_in = isinstance(data, np.ndarray)
    if host_in:
        t = torch.from_numpy(np.ascontiguousarray(data)).cuda()
    else:
        t = as_device_tensor(data)
    if t.dim() != 2:
        raise ValueError("expected a 2-D raster, got %d-D" % t.dim())
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)           # the reference computes on data.astype(float64), slope.py:169
    if t.numel() and t.stride(1) != 1:
        t = t.contiguous()
    H, W = t.shape
    out = torch.empty((H, W), dtype=torch.float32, device=t.device)
    if H and W:
        lat_t = torch.as_tensor(lat, dtype=torch.float64, device=t.device).contiguous()
        lon_t = torch.as_tensor(lon, dtype=torch.float64, device=t.device).contiguous()
        with torch.cuda.device(t.device):
            _lib.call("xrs_geodesic", ctypes.c_void_p(t.data_ptr()), 0 if t.dtype == torch.float32 else 1,
                      t.stride(0) * t.element_size(), ctypes.c_void_p(lat_t.data_ptr()),
                      ctypes.c_void_p(lon_t.data_ptr()), 1 if is_2d else 0, ctypes.c_void_p(out.data_ptr()),
                      out.stride(0) * 4, H, W, float(z_factor), 1 if want_aspect else 0, stream_ptr(t))
    if host_in:
        return out.cpu().numpy()
    return like_container(out, data)
