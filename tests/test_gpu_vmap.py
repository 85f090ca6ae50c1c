"""GPU parity of ``kernex_b200.vmap`` (the reference composes kmap with ``jax.vmap``:
README.md:310-316 depthwise convolution, :331-336 average pooling).  Expected values are the
oracle applied item by item, which is what ``jax.vmap`` means."""
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


def _oracle_items(fn, x, ksize, strides, border, relative=False, extra=None):
    outs = []
    for b in range(x.shape[0]):
        a = () if extra is None else tuple(e[b] if m else e for e, m in extra)
        outs.append(ko.kernel_map({fn: ()}, x.shape[1:], ksize, strides, border, relative)(x[b], *a))
    return np.stack(outs)


def test_readme_depthwise_conv_int(kex):
    # README.md:301-320, bit-exact integer arithmetic
    h, w, c, k = 5, 5, 2, 3
    x = np.arange(1, h * w * c + 1, dtype=np.int32).reshape(c, h, w)
    wt = np.arange(1, k * k * c + 1, dtype=np.int32).reshape(c, k, k)
    conv = kex.vmap(kex.kmap(kernel_size=(3, 3), padding=("same", "same"))(lambda x, w: np.sum(x * w)))
    got = conv(x, torch.from_numpy(wt).cuda())
    want = _oracle_items(lambda x, w: np.sum(x * w), x, (3, 3), (1, 1), ((1, 1), (1, 1)), extra=[(wt, True)])
    assert got.dtype == want.dtype
    np.testing.assert_array_equal(got, want)
    # independent check against a direct correlation
    pad = np.pad(x, ((0, 0), (1, 1), (1, 1)))
    direct = np.zeros_like(x)
    for i in range(3):
        for j in range(3):
            direct += pad[:, i:i + h, j:j + w] * wt[:, i:i + 1, j:j + 1]
    np.testing.assert_array_equal(got, direct)


@pytest.mark.parametrize("shape,ksize,padding", [((4, 33, 70), (3, 3), "same"), ((3, 64, 128), (3, 3), "valid"),
                                                 ((5, 17, 19), (5, 3), "same"), ((2, 40), (3,), "same")])
def test_depthwise_conv_float_device_weights(kex, shape, ksize, padding):
    rng = np.random.default_rng(3)
    x = rng.standard_normal(shape, dtype=np.float32)
    wt = rng.standard_normal((shape[0],) + ksize, dtype=np.float32)
    fn = lambda x, w: np.sum(x * w)
    conv = kex.vmap(kex.kmap(kernel_size=ksize, padding=padding)(fn))
    got = conv(torch.from_numpy(x).cuda(), torch.from_numpy(wt).cuda())
    assert got.is_cuda
    border = ko.resolve_padding(padding, ksize)
    want = _oracle_items(fn, x, ksize, (1,) * len(ksize), border, extra=[(wt, True)])
    tol = 1e-5 * float(np.abs(x).max() * np.abs(wt).max()) * int(np.prod(ksize))
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=tol)


def test_in_axes_broadcast_weight(kex):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((3, 20, 36), dtype=np.float32)
    wt = rng.standard_normal((3, 3), dtype=np.float32)
    fn = lambda x, w: np.sum(x * w)
    conv = kex.vmap(kex.kmap(kernel_size=(3, 3), padding="same")(fn), in_axes=(0, None))
    got = conv(x, torch.from_numpy(wt).cuda())
    want = _oracle_items(fn, x, (3, 3), (1, 1), ((1, 1), (1, 1)), extra=[(wt, False)])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-4)


@pytest.mark.parametrize("shape", [(3, 9, 9), (2, 64, 130), (4, 31, 47)])
def test_readme_avgpool(kex, shape):
    # README.md:327-337: mean over 3x3 windows, stride 2, no padding
    rng = np.random.default_rng(5)
    x = rng.standard_normal(shape, dtype=np.float32)
    pool = kex.vmap(kex.kmap(kernel_size=(3, 3), strides=(2, 2))(lambda x: np.mean(x)))
    got = pool(x)
    want = _oracle_items(lambda x: np.mean(x), x, (3, 3), (2, 2), ((0, 0), (0, 0)))
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-5 * 9 * float(np.abs(x).max()))


@pytest.mark.parametrize("op", ["max", "min"])
def test_maxpool_batched(kex, op):
    x = np.random.default_rng(6).integers(-50, 50, (3, 33, 66)).astype(np.int32)
    fn = (lambda x: np.max(x)) if op == "max" else (lambda x: np.min(x))
    pool = kex.vmap(kex.kmap(kernel_size=(2, 2), strides=(2, 2))(fn))
    got = pool(x)
    want = _oracle_items(fn, x, (2, 2), (2, 2), ((0, 0), (0, 0)))
    np.testing.assert_array_equal(got, want)


def test_vmap_of_mesh_and_patches(kex):
    x = np.arange(1, 2 * 6 * 7 + 1, dtype=np.int32).reshape(2, 6, 7)
    F = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)
    f0, f1 = (lambda x: x[0, 0] * 10), (lambda x: x[0, 1] - x[0, -1])
    F[:, :] = f0
    F[1:-1, 1:-1] = f1
    got = kex.vmap(F)(x)
    keys = {f0: [ko.normalize_index((slice(None), slice(None)), (6, 7))],
            f1: [ko.normalize_index((slice(1, -1), slice(1, -1)), (6, 7))]}
    want = np.stack([ko.kernel_map(keys, (6, 7), (3, 3), (1, 1), ((1, 1), (1, 1)), True)(x[b]) for b in range(2)])
    np.testing.assert_array_equal(got, want)
    # patches: vector-valued function under vmap
    patches = kex.vmap(kex.kmap(kernel_size=(3, 3), padding="same")(lambda x: x))(x)
    wantp = _oracle_items(lambda x: x, x, (3, 3), (1, 1), ((1, 1), (1, 1)))
    np.testing.assert_array_equal(patches, wantp)


def test_vmap_errors(kex):
    with pytest.raises(TypeError):
        kex.vmap(lambda x: x)
    with pytest.raises(kex.UnsupportedStencilError):
        kex.vmap(kex.kscan(kernel_size=(3,))(lambda x: x[0]))
    conv = kex.vmap(kex.kmap(kernel_size=(3,))(lambda x, w: np.sum(x * w)))
    with pytest.raises(ValueError):
        conv(np.ones((2, 8), np.float32), torch.ones(3, 3).cuda())  # leading axes disagree
