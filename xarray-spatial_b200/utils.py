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
    _lib.call("xrs_host_stencil", _lib.OPS[op], ctypes.c_void_p(d.ctypes.data),
              ctypes.c_void_p(out.ctypes.data), H, W, _dbl_array(p), _dbl_array(aux), len(aux),
              host_device_index())
    return out
