"""GPU parity of kscan: the row-marching fast path (time-marching stencils that read only
the row above) and the general wavefront kernel, against the oracle."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import kx_oracle as ko
from oracle import kx_oracle_c as kc
from oracle_iface import OracleIface

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


def _convection_mesh(cls, nt, nx, kappa):
    F = cls(kernel_size=(3, 3), padding=((1, 1), (1, 1)), named_axis={0: "n", 1: "i"}, relative=True)
    q1, q2 = int((nx - 1) / 4) + 1, int((nx - 1) / 2)
    F[:, 0] = F[:, -1] = lambda u: 1
    F[:, :q1] = F[:, q2:] = lambda u: 1
    F[0:1, q1:q2] = lambda u: 2
    F[1:, 1:-1] = lambda u: u["i", "n-1"] - kappa * (u["i", "n-1"] - u["i-1", "n-1"])
    regions = [(0, nt, 0, 1, (0, 0, 0), 1.0), (0, nt, nx - 1, nx, (0, 0, 0), 1.0),
               (0, nt, 0, q1, (0, 0, 0), 1.0), (0, nt, q2, nx, (0, 0, 0), 1.0),
               (0, 1, q1, q2, (0, 0, 0), 2.0), (1, nt, 1, nx - 1, (kappa, 1.0 - kappa, 0.0), 0.0)]
    return F, regions


@pytest.mark.parametrize("nt,nx", [(70, 301), (130, 4096), (300, 5003), (64, 2048), (65, 9000)])
def test_rowmarch_convection_matches_oracle(kex, nt, nx):
    kappa = 0.5
    F, regions = _convection_mesh(kex.kscan, nt, nx, kappa)
    got = F(np.ones((nt, nx), np.float32))
    assert _last_kernel() == "scan_rowmarch_kernel"
    want = kc.rowmarch(nt, nx, regions)
    # convex update (|coefs| sum to 1): rounding does not amplify; fp32 FMA vs mul+add
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
    if nt * nx < 40000:  # literal sequential oracle (the reference's sweep order) on the small case
        Fo, _ = _convection_mesh(lambda **kw: OracleIface(True, **kw), nt, nx, kappa)
        np.testing.assert_allclose(got, Fo(np.ones((nt, nx), np.float32)), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("seed", range(6))
def test_rowmarch_random_meshes_match_literal_oracle(kex, seed):
    rng = np.random.default_rng(500 + seed)
    nt, nx = int(rng.integers(66, 80)), int(rng.integers(40, 400))
    k1 = 3 if seed % 2 else 5
    r = (k1 - 1) // 2
    coefs = rng.uniform(-0.3, 0.3, size=(3, k1))
    bias = rng.uniform(-1, 1, size=3)

    def make(cls):
        F = cls(kernel_size=(3, k1), padding="same", relative=True)
        fs = []
        for q in range(3):
            def f(u, q=q):
                acc = float(bias[q])
                for d in range(-r, r + 1):
                    acc = acc + float(coefs[q, d + r]) * u[-1, d]
                return acc
            fs.append(f)
        F[:] = fs[0]
        F[3:nt - 5, 2:] = fs[1]
        F[nt // 2, :] = F[::1, nx // 3: nx // 2: 2] = fs[2]     # strided COLUMN key (centre % 2 == 0)
        F[0] = lambda u: 0.5
        return F

    x = rng.standard_normal((nt, nx)).astype(np.float32)
    got = make(kex.kscan)(x)
    assert _last_kernel() == "scan_rowmarch_kernel"
    want = make(lambda **kw: OracleIface(True, **kw))(x)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)


def test_scan_reading_same_row_uses_the_wavefront_kernel(kex):
    # README kscan sum (README.md:63-69): reads the just-written left neighbour -> serial
    x = np.array([1, 2, 3, 4, 5], dtype=np.int32)
    out = kex.kscan(kernel_size=(3,))(lambda v: np.sum(v))(x)
    assert _last_kernel() in ("scan_wavefront_kernel", "scan_wavefront_rows_kernel")
    np.testing.assert_array_equal(out, [6, 13, 22])
    rng = np.random.default_rng(2)
    x2 = rng.integers(-3, 4, (6, 7)).astype(np.int32)
    f = lambda v: v[0, -1] + v[-1, 0] - v[0, 0]
    got = kex.kscan(kernel_size=(3, 3), padding="same", relative=True)(f)(x2)
    want = ko.kernel_scan({f: ()}, (6, 7), (3, 3), (1, 1), ((1, 1), (1, 1)), True)(x2)
    np.testing.assert_array_equal(got, want)


def test_diffusion_sscan_matches_numpy_fdm(kex):
    """tests/test_interface.py:28-97: 3-D sscan with named axes and 4 runtime scalars vs the
    NumPy FDM loop (wavefront kernel, one plane per step, + offset write-back)."""
    nx = ny = 10
    nt, nu = 5, 1.0
    dx, dy = 2 / (nx - 1), 2 / (ny - 1)
    dt = 0.25 * dx * dy / nu

    @kex.sscan(kernel_size=(3, 3, 3), offset=((1, 0), (1, 1), (1, 1)), named_axis={0: "n", 1: "i", 2: "j"},
               relative=True)
    def diffusion(T, nu, dt, dx, dy):
        return (T["n-1", "i", "j"]
                + (nu * dt) / (dy ** 2) * (T["n-1", "i+1", "j"] - 2 * T["n-1", "i", "j"] + T["n-1", "i-1", "j"])
                + (nu * dt) / (dx ** 2) * (T["n-1", "i", "j+1"] - 2 * T["n-1", "i", "j"] + T["n-1", "i", "j-1"]))

    u = np.ones((ny, nx))
    u[int(0.5 / dy):int(1 / dy + 1), int(0.5 / dx):int(1 / dx + 1)] = 2
    U = np.ones((nt + 3, ny, nx), dtype=np.float32)
    U[0] = u
    for n in range(nt + 1):
        un = u.copy()
        u[1:-1, 1:-1] = (un[1:-1, 1:-1] + nu * dt / dx ** 2 * (un[1:-1, 2:] - 2 * un[1:-1, 1:-1] + un[1:-1, 0:-2])
                         + nu * dt / dy ** 2 * (un[2:, 1:-1] - 2 * un[1:-1, 1:-1] + un[0:-2, 1:-1]))
        u[0, :] = u[-1, :] = u[:, 0] = u[:, -1] = 1
    got = diffusion(U, nu, dt, dx, dy)
    assert _last_kernel() == "scan_wavefront_kernel"
    np.testing.assert_allclose(got[-2], u, rtol=1e-5, atol=1e-5)


def test_rowmarch_config4_prefix_strip_full_depth(kex):
    """BASELINE config 4 semantics at nt=4096: the update reads only (n-1, i) and (n-1, i-1), so
    the first W columns of the full problem equal the same mesh solved on those W columns plus
    nothing else -- the oracle is run on that strip and compared for all 4096 rows."""
    nt, nx, W = 4096, 1 << 18, 3000
    kappa = 0.5
    F, _ = _convection_mesh(kex.kscan, nt, nx, kappa)
    u = torch.ones((nt, nx), device="cuda")
    out = F(u)
    assert _last_kernel() == "scan_rowmarch_kernel"
    assert out.shape == (nt, nx)
    q1, q2 = int((nx - 1) / 4) + 1, int((nx - 1) / 2)
    assert W < q1  # the strip lies inside the first ic1 region: row 0 is all ones there
    regions = [(0, nt, 0, 1, (0, 0, 0), 1.0), (0, nt, 0, W, (0, 0, 0), 1.0),
               (1, nt, 1, W, (kappa, 1.0 - kappa, 0.0), 0.0)]
    want = kc.rowmarch(nt, W, regions)
    np.testing.assert_allclose(out[:, :W].cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    # ... and ON THE WAVE FRONT, where the recurrence is non-trivial for all 4096 rows: column q1-1 reads only
    # ones (it stays exactly 1), so the strip [q1-1, q1-1+W) is the same mesh with a constant-1 left boundary,
    # row 0 = 2 and the update below; row n of it is 1 + P(Binomial(n, 1/2) <= j), the smoothed front that
    # sits near column q1 + n/2 (inside the strip down to the last row).
    assert q1 - 1 + W <= q2
    front = [(0, nt, 0, 1, (0, 0, 0), 1.0), (0, 1, 1, W, (0, 0, 0), 2.0), (1, nt, 1, W, (kappa, 1.0 - kappa, 0.0), 0.0)]
    want_f = kc.rowmarch(nt, W, front)
    got_f = out[:, q1 - 1:q1 - 1 + W].cpu().numpy()
    for n in (1, 63, 64, 65, 1000, 2048, 4095):                 # the strip really varies at every depth
        assert np.ptp(want_f[n]) > 0.9, n
        mid = np.count_nonzero((want_f[n] > 1.01) & (want_f[n] < 1.99))
        assert mid >= min(n, 15), (n, mid)
    np.testing.assert_allclose(got_f, want_f, rtol=1e-5, atol=1e-5)
    # boundary / initial conditions hold exactly
    assert torch.all(out[:, 0] == 1) and torch.all(out[:, -1] == 1)
    assert torch.all(out[0, q1:q2] == 2) and torch.all(out[0, :q1] == 1) and torch.all(out[0, q2:] == 1)
    # the square wave is transported, never amplified (convex combination of the row above)
    assert float(out.max()) <= 2.0 and float(out.min()) >= 1.0


@pytest.mark.parametrize("world", [2, 3, 8])
def test_column_sharded_scan_is_bit_identical_to_unsharded(kex, world):
    """kscan shards on axis 1 (axis 0 is time): every emulated rank recomputes its nt*R-column
    dependency cone instead of exchanging it; the stitched result must equal the single-launch
    result bit for bit (no communication is involved, so the ranks can run one after another
    on one GPU)."""
    from kernex_b200.dist import ColumnShardedScan
    nt, nx, kappa = 96, 6000, 0.5
    F, _ = _convection_mesh(kex.kscan, nt, nx, kappa)
    whole = F(torch.ones((nt, nx), device="cuda"))
    parts = []
    for rank in range(world):
        Fr, _ = _convection_mesh(kex.kscan, nt, nx, kappa)
        sh = ColumnShardedScan(Fr, nt, nx, rank=rank, world=world)
        assert sh.rl == 1 and sh.rr == 0
        parts.append(sh.run().clone())
        assert _last_kernel() == "scan_rowmarch_kernel"
    stitched = torch.cat(parts, dim=1)
    assert stitched.shape == whole.shape
    assert torch.equal(stitched, whole)
