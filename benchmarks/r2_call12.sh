set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_convc.py -q -x 2>&1 | tail -3
python - <<'PY' | tee gpurun_out/r2_convc_bench.txt
import sys, json, os
sys.path.insert(0, "benchmarks"); sys.path.insert(0, ".")
import numpy as np, torch
import kernex_b200 as kex
from kernex_b200 import _lib
from run_configs import timeit, peak
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
for C, H in [(4, 1024), (8, 1024), (16, 1024), (32, 1024), (64, 1024), (16, 4096), (64, 2048), (8, 8192)]:
    x = torch.randn((C, H, H), device=dev, generator=g)
    w = torch.randn((C, 3, 3), device=dev, generator=g)
    conv = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"), relative=False)(lambda v, w: np.sum(v * w))
    y = conv(x, w)
    med, best = timeit(lambda: conv(x, w), 10)
    nbytes = (C + 1) * H * H * 4
    print(json.dumps({"shape": [C, H, H], "us": med * 1e3, "glups": H * H / (med * 1e-3) / 1e9, "alg_gbs": nbytes / (med * 1e-3) / 1e9,
                      "frac_of_measured_peak": nbytes / (med * 1e-3) / 1e9 / peak(), "kernel": _lib.last_kernel()}), flush=True)
    if (C, H) in ((16, 1024), (64, 1024)):
        os.environ["KX_NO_CONVC"] = "1"
        conv2 = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"), relative=False)(lambda v, w: np.sum(v * w))
        conv2(x, w)
        med2, _ = timeit(lambda: conv2(x, w), 3)
        del os.environ["KX_NO_CONVC"]
        print(json.dumps({"shape": [C, H, H], "generic_kernel_us": med2 * 1e3, "speedup": med2 / med}), flush=True)
    del x, y
    torch.cuda.empty_cache()
PY
