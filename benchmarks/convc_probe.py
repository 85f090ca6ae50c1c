#!/usr/bin/env python
"""convc_kernel (conv-as-kmap, many channels) over the reference's benchmark shapes and a few HBM-sized ones, once per
ring depth (KX_CONVC_STAGES = 3 / 6 / unset = the heuristic).  One JSON line per (shape, setting); device time from CUDA
events over back-to-back launches, fraction = (C*4+4) bytes per output element against the measured copy bandwidth."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex
from kernex_b200 import _lib

dev = torch.device("cuda")
peak = 6547.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbps"]
except Exception:
    pass


def device_time(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


g = torch.Generator(device=dev).manual_seed(0)
shapes = [(8, 256), (8, 512), (8, 1024), (16, 512), (16, 1024), (32, 512), (32, 1024), (64, 512), (64, 1024), (64, 2048), (16, 4096), (8, 8192)]
for C, H in shapes:
    conv = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"), relative=False)(lambda x, w: np.sum(x * w))
    x = torch.randn((C, H, H), device=dev, generator=g)
    w = torch.randn((C, 3, 3), device=dev, generator=g)
    ref = None
    for setting in ("", "3", "6"):
        if setting:
            os.environ["KX_CONVC_STAGES"] = setting
        else:
            os.environ.pop("KX_CONVC_STAGES", None)
        y = conv(x, w)
        if ref is None:
            ref = y.clone()
        same = bool(torch.equal(y, ref))
        iters = 200 if H <= 1024 else 20
        us = device_time(lambda: conv(x, w), iters)
        gbs = (C * 4 + 4) * H * H / (us * 1e-6) / 1e9
        print(json.dumps({"dim": [C, H, H], "stages": setting or "auto", "device_us": round(us, 2), "GBps": round(gbs, 1),
                          "frac": round(gbs / peak, 3), "bit_identical": same, "kernel": _lib.last_kernel()}), flush=True)
    del x, y, ref
    torch.cuda.empty_cache()
