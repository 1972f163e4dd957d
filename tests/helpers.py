"""Shared test helpers: tolerances of the parity gate and input makers."""
import numpy as np


def assert_same_nan(a, b, what=""):
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b), err_msg="NaN mask differs " + what)


def assert_close_f32(gpu, ref, rtol=1e-5, atol=1e-6, what=""):
    """|gpu - ref| <= rtol*|ref| + atol, identical NaN masks (float32 ops, bar 1e-5 relative)."""
    gpu = np.asarray(gpu, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert gpu.shape == ref.shape, (gpu.shape, ref.shape)
    assert_same_nan(gpu, ref, what)
    m = ~np.isnan(ref)
    inf = m & np.isinf(ref)
    np.testing.assert_array_equal(gpu[inf], ref[inf])
    m &= ~np.isinf(ref)
    err = np.abs(gpu[m] - ref[m])
    tol = rtol * np.abs(ref[m]) + atol
    bad = err > tol
    assert not bad.any(), "%s: %d cells out of tolerance, worst err %g (ref %g)" % (
        what, bad.sum(), err[bad].max(), ref[m][bad][np.argmax(err[bad])])


def assert_aspect_close(gpu, ref, what=""):
    """aspect: -1 (flat) mask identical, NaN mask identical, circular distance within
    1e-5 relative + 1e-4 degrees."""
    gpu = np.asarray(gpu, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert_same_nan(gpu, ref, what)
    np.testing.assert_array_equal(gpu == -1, ref == -1, err_msg="flat mask differs " + what)
    m = ~np.isnan(ref) & (ref != -1)
    d = np.abs(gpu[m] - ref[m])
    d = np.minimum(d, 360.0 - d)
    tol = 1e-5 * np.abs(ref[m]) + 1e-4
    assert (d <= tol).all(), "%s: aspect worst circular err %g" % (what, d.max())


def terrain(rng, h, w, water=False, nans=0.0, integer=False, zmax=4000.0):
    z = rng.standard_normal((h, w)).cumsum(0).cumsum(1)
    z += np.linspace(0, 30, w)[None, :] + np.linspace(0, 10, h)[:, None]
    z = (z - z.min()) / (z.max() - z.min() + 1e-9) * zmax
    if water:
        z[z < 0.3 * z.max()] = 0.0
    if integer:
        z = np.round(z)
    z = z.astype(np.float32)
    if nans:
        z[rng.random((h, w)) < nans] = np.nan
    return z
