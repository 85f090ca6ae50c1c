"""GPU tests of the host pipeline and of the slab-decomposed multi-GPU path."""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import kx_oracle as ko

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kex():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kernex_b200
    return kernex_b200


def lap(x):
    return x[1, 0] + x[0, -1] - 4 * x[0, 0] + x[0, 1] + x[-1, 0]


@pytest.mark.parametrize("padding", ["valid", "same"])
def test_pinned_host_pipeline_matches_plain_path(kex, padding):
    """>= 64 MB pinned NumPy input goes through hostpipe.py (3-stream slab pipeline)."""
    from kernex_b200 import _lib
    rows, cols = 4200, 4096
    pinned = torch.empty((rows, cols), dtype=torch.float32, pin_memory=True)
    x = pinned.numpy()
    x[...] = np.random.default_rng(0).standard_normal((rows, cols), dtype=np.float32)
    f = kex.kmap(kernel_size=(3, 3), padding=padding, relative=True)(lap)
    before = _lib.launch_count()
    got = f(x)
    assert _lib.launch_count() - before >= 2, "expected several slab launches"
    want = f(torch.from_numpy(x).cuda()).cpu().numpy()   # device-resident path, single launch
    np.testing.assert_array_equal(got, want)
    b = ((0, 0), (0, 0)) if padding == "valid" else ((1, 1), (1, 1))
    ref = ko.affine_map(x[:300], (3, 3), (1, 1), (b[0] if padding == "valid" else (1, 0), b[1]),
                        [(2, 1), (1, 0), (1, 1), (1, 2), (0, 1)], [1, 1, -4, 1, 1])
    n = ref.shape[0] - 2
    np.testing.assert_allclose(got[:n], ref[:n], rtol=0, atol=2e-5)


def test_slab_stencil_world1_equals_kmap(kex):
    from kernex_b200.dist import SlabStencil
    slab = SlabStencil(lap, kernel_size=(3, 3), relative=True, rows_per_rank=96, width=200, padding="same")
    slab.fill_random(torch.Generator(device="cuda").manual_seed(1))
    out = slab.step()
    want = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)(lap)(slab.ext)
    assert torch.equal(out, want)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_slab_stencil_multi_gpu_bit_identical():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "MISMATCH" not in res.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_tensor_on_a_non_current_device(kex):
    """Several GPUs in ONE process: the library switches to the device that owns the buffers for the call."""
    assert torch.cuda.current_device() == 0
    rng = np.random.default_rng(3)
    xh = rng.standard_normal((300, 516)).astype(np.float32)
    x1 = torch.from_numpy(xh).to("cuda:1")
    f = kex.kmap(kernel_size=(3, 3), padding="same", relative=True)(lap)
    y1 = f(x1)
    assert y1.device == x1.device and torch.cuda.current_device() == 0
    y0 = f(torch.from_numpy(xh).to("cuda:0"))
    assert torch.equal(y1.cpu(), y0.cpu())
    s1 = kex.kscan(kernel_size=(3, 3), padding="same", relative=True)(lambda u: 0.25 * (u[0, -1] + u[-1, 0] + u[0, 1] + u[1, 0]))
    assert torch.equal(s1(x1).cpu(), s1(torch.from_numpy(xh).to("cuda:0")).cpu())
