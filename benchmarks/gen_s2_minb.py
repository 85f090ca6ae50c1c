#!/usr/bin/env python
"""Regenerate kernex_b200/csrc/kx_stencil2d_minb.inc: resident blocks per SM of every stencil2d_kernel instantiation, from
the ptxas logs (build/csrc/kx_stencil2d_k*.ptxas.log) of a build WITHOUT the cap (-DKX_S2_MINB=1 and the bounds-tested
path only: git show 0c57be6:kernex_b200/csrc/kx_stencil2d_impl.cuh)."""
import glob, os, re
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ent = {}
for f in glob.glob(os.path.join(root, "build/csrc/kx_stencil2d_k*.ptxas.log")):
    for m in re.finditer(r"Compiling entry function '_ZN2kx2s216stencil2d_kernelI([fi])Li(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)E.*?Used (\d+) registers",
                         open(f).read(), re.S):
        t = 1 if m.group(1) == "f" else 0
        ky, kx, sh, vw, u, op, regs = [int(g) for g in m.groups()[1:]]
        key = ((((((t * 16 + ky) * 16 + kx) * 4 + sh) * 8 + vw) * 8 + u) * 8 + op)
        ent[key] = min(65536 // (((regs + 7) // 8 * 8) * 128), 16)
items = sorted(ent.items())
print("static constexpr unsigned kS2MinB[][2] = {")
for i in range(0, len(items), 8):
    print("    " + " ".join("{%d, %d}," % kv for kv in items[i:i + 8]))
print("};")
