"""The OpenMP C restatement (CPU baseline) agrees with the pinned NumPy oracle."""
from __future__ import annotations

import itertools

import numpy as np
import pytest

from oracle import kx_oracle as ko
from oracle import kx_oracle_c as kc


@pytest.mark.parametrize("seed", range(16))
def test_affine_matches_numpy_oracle(seed):
    rng = np.random.default_rng(seed)
    nd = 1 + seed % 3
    shape = tuple(int(v) for v in rng.integers(5, 12, size=nd))
    ksize = tuple(int(rng.integers(1, 4)) for _ in shape)
    strides = tuple(int(rng.integers(1, 3)) for _ in shape)
    border = tuple((int(rng.integers(-1, 3)), int(rng.integers(-1, 3))) for _ in shape)
    if any(o <= 0 for o in ko.output_shape(shape, ksize, strides, border)):
        pytest.skip("empty")
    taps = list(itertools.product(*[range(k) for k in ksize]))
    if seed % 2:
        x = rng.standard_normal(shape).astype(np.float32)
        w = rng.standard_normal(len(taps)).astype(np.float32)
        got = kc.affine_map(x, ksize, strides, border, taps, w, bias=0.5)
        want = ko.affine_map(x, ksize, strides, border, taps, w, bias=0.5)
        np.testing.assert_array_equal(got, want)  # same tap order, no FMA contraction -> bit equal
    else:
        x = rng.integers(-99, 99, shape).astype(np.int32)
        w = rng.integers(-5, 5, len(taps)).astype(np.int32)
        np.testing.assert_array_equal(kc.affine_map(x, ksize, strides, border, taps, w, bias=3),
                                      ko.affine_map(x, ksize, strides, border, taps, w, bias=3))


def test_patch_matches_numpy_oracle():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((9, 11)).astype(np.float32)
    got = kc.patch_map(x, (3, 3), (1, 1), ((1, 1), (1, 1)))
    np.testing.assert_array_equal(got, ko.patch_map(x, (3, 3), (1, 1), ((1, 1), (1, 1))))


def test_rowmarch_matches_literal_scan():
    import catalog
    nt, nx = 14, 19

    # literal oracle result of the README convection mesh (tests/golden/catalog.py)
    from oracle_iface import OracleIface as _OracleIface
    F, x = catalog.MESH_SCENARIOS["mesh_convection"](lambda **kw: _OracleIface(True, **kw), np.random.default_rng(0))
    want = F(x)
    kappa = 0.5 * (0.5 / (nt - 1)) / (2.0 / (nx - 1))
    q1, q2 = int((nx - 1) / 4) + 1, int((nx - 1) / 2)
    regions = [  # insertion order of catalog._mesh_convection: bc, ic1, ic2, convection (later wins)
        (0, nt, 0, 1, (0, 0, 0), 1.0), (0, nt, nx - 1, nx, (0, 0, 0), 1.0),
        (0, nt, 0, q1, (0, 0, 0), 1.0), (0, nt, q2, nx, (0, 0, 0), 1.0),
        (0, 1, q1, q2, (0, 0, 0), 2.0),
        (1, nt, 1, nx - 1, (kappa, 1.0 - kappa, 0.0), 0.0),
    ]
    got = kc.rowmarch(nt, nx, regions)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("seed", range(12))
def test_views_gather_restatement_matches_literal_oracle(seed):
    """kxo_map_views_f32 (pad at the end + materialised views + gather + roll + function, the
    reference's own algorithm, used by bench.py's reference arm) against the literal NumPy oracle."""
    rng = np.random.default_rng(300 + seed)
    shape = (int(rng.integers(5, 14)), int(rng.integers(5, 14)))
    ksize = (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
    strides = (int(rng.integers(1, 3)), int(rng.integers(1, 3)))
    border = tuple((int(rng.integers(0, 3)), int(rng.integers(0, 3))) for _ in shape)
    relative = bool(seed % 2)
    if any(o <= 0 for o in ko.output_shape(shape, ksize, strides, border)):
        pytest.skip("empty")
    taps = list(itertools.product(range(ksize[0]), range(ksize[1])))   # patch coordinates
    w = rng.standard_normal(len(taps)).astype(np.float32)
    x = rng.standard_normal(shape).astype(np.float32)

    def fn(p):
        acc = np.float32(0.5)
        for t, c in zip(taps, w):
            acc = acc + c * p[t]
        return acc

    want = ko.kernel_map({fn: ()}, shape, ksize, strides, border, relative)(x)
    got = kc.map_views(x, ksize, strides, border, taps, w, bias=0.5, relative=relative)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)


def test_views_gather_restatement_readme_laplacian():
    # README.md:122-143: Laplacian of ones (valid) is zero; relative indices -1 -> patch index k-1
    x = np.ones((10, 10), np.float32)
    rel = {-1: 2, 0: 0, 1: 1}
    terms = [((1, -1), 0), ((1, 0), 1), ((1, 1), 0), ((0, -1), 1), ((0, 0), -4), ((0, 1), 1), ((-1, -1), 0),
             ((-1, 0), 1), ((-1, 1), 0)]
    taps = [(rel[a], rel[b]) for (a, b), _ in terms]
    got = kc.map_views(x, (3, 3), (1, 1), ((0, 0), (0, 0)), taps, [c for _, c in terms], relative=True)
    np.testing.assert_array_equal(got, np.zeros((8, 8), np.float32))
