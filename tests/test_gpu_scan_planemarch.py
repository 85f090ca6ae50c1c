"""kscan / sscan over a 3-D array whose function reads only the plane above the cell it writes (time marching with two
space axes, the reference's diffusion test tests/test_interface.py:28-97): one 2-D kmap launch per plane
(kx_abi.cu try_scan_planemarch) instead of the wavefront sweep.  Checked against the literal oracle, against the
wavefront kernel (KX_NO_PLANEMARCH=1: int32 bit for bit) and against a NumPy FDM loop."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import kx_oracle as ko

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def kex():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kernex_b200
    return kernex_b200


def _last_kernel():
    from kernex_b200 import _lib
    return _lib.last_kernel()


def _step(integer):
    if integer:
        return lambda v: v[-1, 0, 0] + 2 * v[-1, 1, 0] - v[-1, -1, 0] + 3 * v[-1, 0, 1] - v[-1, 0, -1] + 1
    return lambda v: (v[-1, 0, 0] + 0.1 * (v[-1, 1, 0] - 2 * v[-1, 0, 0] + v[-1, -1, 0])
                      + 0.05 * (v[-1, 0, 1] - 2 * v[-1, 0, 0] + v[-1, 0, -1]) + 0.01)


BORDERS = {
    "same (full cover)": ((1, 1), (1, 1), (1, 1)),
    "sscan-like (interior only)": ((0, 1), (0, 0), (0, 0)),
    "mixed": ((1, 0), (1, 0), (0, 1)),
    "wide pads": ((2, 1), (1, 2), (2, 2)),
}


@pytest.mark.parametrize("border_name", list(BORDERS))
@pytest.mark.parametrize("integer", [True, False])
def test_planemarch_kscan_matches_oracle_and_wavefront(kex, border_name, integer, monkeypatch):
    border = BORDERS[border_name]
    shape, ksize = (5, 130, 140), (3, 3, 3)
    rng = np.random.default_rng(len(border_name) + integer)
    x = rng.integers(-3, 4, shape).astype(np.int32) if integer else rng.standard_normal(shape).astype(np.float32)
    fmap = {_step(integer): ()}
    got = kex.kernel_scan(fmap, shape, ksize, (1, 1, 1), border, True)(x)
    assert _last_kernel() in ("stencil2d_kernel", "generic_map_kernel"), _last_kernel()     # per-plane 2-D launches
    want = ko.kernel_scan(fmap, shape, ksize, (1, 1, 1), border, True)(x)
    monkeypatch.setenv("KX_NO_PLANEMARCH", "1")
    wave = kex.kernel_scan(fmap, shape, ksize, (1, 1, 1), border, True)(x)
    assert _last_kernel().startswith("scan_")
    assert got.shape == want.shape and got.dtype == want.dtype
    if integer:
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(got, wave)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(got, wave, rtol=1e-5, atol=1e-5)


def test_planemarch_sscan_diffusion_matches_numpy_fdm(kex):
    """The reference's diffusion test at a size where the planes are worth a launch each (256 x 512), 12 steps."""
    ny, nx, nt = 256, 512, 12
    a, b = 0.2, 0.15

    @kex.sscan(kernel_size=(3, 3, 3), offset=((1, 0), (1, 1), (1, 1)), named_axis={0: "n", 1: "i", 2: "j"}, relative=True)
    def diffusion(T, a, b):
        return (T["n-1", "i", "j"] + a * (T["n-1", "i+1", "j"] - 2 * T["n-1", "i", "j"] + T["n-1", "i-1", "j"])
                + b * (T["n-1", "i", "j+1"] - 2 * T["n-1", "i", "j"] + T["n-1", "i", "j-1"]))

    rng = np.random.default_rng(1)
    U = rng.standard_normal((nt, ny, nx)).astype(np.float32)      # every plane has its own border values
    got = diffusion(torch.from_numpy(U).cuda(), a, b).cpu().numpy()
    ref = U.astype(np.float64).copy()
    for n in range(1, nt):
        p = ref[n - 1]
        ref[n, 1:-1, 1:-1] = (p[1:-1, 1:-1] + a * (p[2:, 1:-1] - 2 * p[1:-1, 1:-1] + p[:-2, 1:-1])
                              + b * (p[1:-1, 2:] - 2 * p[1:-1, 1:-1] + p[1:-1, :-2]))
    np.testing.assert_array_equal(got[0], U[0])
    np.testing.assert_array_equal(got[:, 0, :], U[:, 0, :])       # border cells keep their input values
    np.testing.assert_array_equal(got[:, :, -1], U[:, :, -1])
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)


def test_planemarch_mesh_2d_convection_with_boundary_conditions(kex, monkeypatch):
    """The README's linear-convection mesh (README.md:231-266) with TWO space axes: `F[slice] = fn` boundary / initial
    conditions + an upwind update that reads the plane above, as a 3-D kscan with 'same'-style borders.  Every plane is a
    2-D multi-function kmap (mesh kernel); int32 coefficients make the comparison with the wavefront kernel exact."""
    nt, ny, nx = 6, 140, 260

    def mesh(factory, integer):
        F = factory(kernel_size=(3, 3, 3), padding=((1, 1), (1, 1), (1, 1)), relative=True)
        F[:, 0, :] = F[:, -1, :] = lambda u: 1                    # boundary conditions
        F[:, :, 0] = F[:, :, -1] = lambda u: 1
        F[0, 1:-1, 1:-1] = lambda u: 1                             # initial conditions
        F[0, 40:80, 60:120] = lambda u: 2
        if integer:
            F[1:, 1:-1, 1:-1] = lambda u: 3 * u[-1, 0, 0] - u[-1, -1, 0] - u[-1, 0, -1]
        else:
            F[1:, 1:-1, 1:-1] = lambda u: u[-1, 0, 0] - 0.3 * (u[-1, 0, 0] - u[-1, -1, 0]) - 0.2 * (u[-1, 0, 0] - u[-1, 0, -1])
        return F

    for integer in (True, False):
        x = (np.ones((nt, ny, nx), dtype=np.int32) if integer else np.ones((nt, ny, nx), dtype=np.float32))
        got = mesh(kex.kscan, integer)(torch.from_numpy(x).cuda()).cpu().numpy()
        assert _last_kernel() in ("stencil2d_mesh_kernel", "generic_map_kernel", "stencil2d_kernel"), _last_kernel()
        monkeypatch.setenv("KX_NO_PLANEMARCH", "1")
        wave = mesh(kex.kscan, integer)(torch.from_numpy(x).cuda()).cpu().numpy()
        assert _last_kernel().startswith("scan_")
        monkeypatch.delenv("KX_NO_PLANEMARCH")
        # NumPy reference of the same scheme
        ref = np.ones((nt, ny, nx), dtype=np.float64)
        ref[0, 40:80, 60:120] = 2
        for n in range(1, nt):
            p = ref[n - 1]
            if integer:
                ref[n, 1:-1, 1:-1] = 3 * p[1:-1, 1:-1] - p[:-2, 1:-1] - p[1:-1, :-2]
            else:
                ref[n, 1:-1, 1:-1] = p[1:-1, 1:-1] - 0.3 * (p[1:-1, 1:-1] - p[:-2, 1:-1]) - 0.2 * (p[1:-1, 1:-1] - p[1:-1, :-2])
        if integer:
            np.testing.assert_array_equal(got, wave)
            np.testing.assert_array_equal(got, ref.astype(np.int32))
        else:
            np.testing.assert_allclose(got, wave, rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)
