"""Counter-based synthetic streams: the NumPy restatement (s2geometry_b200/synth.py) equals the host
compilation of the device header (csrc/s2b_synth.cuh) bit for bit, slices compose, and the distributions are what the
benchmark configurations say (uniform on the sphere / in a cap)."""
import ctypes

import numpy as np

from s2geometry_b200 import synth
from tests.util import vp


def _hs(hs, name, seed, first, n, cols, *extra):
    out = np.empty((n, cols))
    getattr(hs, name)(ctypes.c_uint64(seed), ctypes.c_uint64(first), ctypes.c_size_t(n), *extra, vp(out))
    return out


def test_numpy_restatement_equals_device_header(hostsim):
    for seed, first, n in ((1, 0, 200_000), (2**63 + 5, 10**12, 50_000), (7, 2**40 - 3, 1000)):
        assert np.array_equal(synth.stream_sphere_points(seed, first, n), _hs(hostsim, "hs_synth_sphere_points", seed, first, n, 3))
        assert np.array_equal(synth.stream_sphere_latlng(seed, first, n), _hs(hostsim, "hs_synth_sphere_latlng", seed, first, n, 2))
        c = np.array([0.3, -0.5, 0.8])
        f = synth.cap_frame(c)
        assert np.array_equal(synth.stream_cap_points(seed, first, n, c, 0.2125),
                              _hs(hostsim, "hs_synth_cap_points", seed, first, n, 3, vp(f), ctypes.c_double(synth.cap_height(0.2125))))


def test_slices_compose_and_seeds_differ():
    whole = synth.stream_sphere_points(4, 1000, 30_000)
    parts = np.concatenate([synth.stream_sphere_points(4, 1000 + lo, 10_000) for lo in (0, 10_000, 20_000)])
    assert np.array_equal(whole, parts)
    assert not np.array_equal(whole, synth.stream_sphere_points(5, 1000, 30_000))
    ll = synth.stream_sphere_latlng(4, 0, 30_000)
    assert np.array_equal(ll[7000:9000], synth.stream_sphere_latlng(4, 7000, 2000))


def test_distributions():
    p = synth.stream_sphere_points(11, 0, 1_000_000)
    assert np.abs(np.sqrt((p * p).sum(1)) - 1).max() < 4e-16
    assert np.abs(p.mean(0)).max() < 3e-3 and np.abs((p * p).mean(0) - 1 / 3).max() < 2e-3
    ll = synth.stream_sphere_latlng(11, 0, 1_000_000)
    assert np.abs(ll[:, 0]).max() < np.pi / 2 and np.abs(ll[:, 1]).max() < np.pi
    assert abs(np.sin(ll[:, 0]).mean()) < 3e-3 and abs((np.sin(ll[:, 0]) ** 2).mean() - 1 / 3) < 2e-3   # sin(lat) uniform
    z = np.linspace(-1, 1, 200_001)
    assert np.abs(synth._asin(z) - np.arcsin(z)).max() < 4e-16
    c = np.array([0.3, -0.5, 0.8]); c /= np.linalg.norm(c)
    q = synth.stream_cap_points(11, 0, 500_000, c, 0.2125)
    d = q @ c
    assert d.min() >= np.cos(0.2125) - 1e-15 and np.abs(np.sqrt((q * q).sum(1)) - 1).max() < 4e-16
    assert abs(d.mean() - (1 + np.cos(0.2125)) / 2) < 1e-4                      # uniform in height
