"""Generate the golden fixtures under ``tests/golden/``.

Run HERE (the build container), where ``/root/reference`` exists:

    python tests/golden/make_golden.py

Two fixture files are written and committed:

``engine_vectors.json``
    Every ``assert_array_equal(result, pred)`` of the reference's
    ``tests/test_scan_and_map.py`` (106 vectors, direct calls of the four L1
    engine factories) extracted with ``ast``: the call arguments are kept as
    source text (the window functions are lambdas), the expected array is
    evaluated to numbers.  Nothing of the reference is copied as code; the
    fixture is data: arguments + expected output + the ``file:line`` it came from.

``shim_cases.npz`` + ``shim_cases.json``
    Outputs of the reference's *own* ``kernel_map / kernel_scan /
    offset_kernel_map / offset_kernel_scan`` and of the ``kmap/kscan/smap/sscan``
    interface classes, executed on ``oracle/jax_shim`` (NumPy stand-in for the
    JAX primitives; real JAX is not installable here) over seeded random
    geometries and the window functions of ``tests/golden/catalog.py``.

The GPU box has no ``/root/reference``; tests there read only these fixtures.
"""

from __future__ import annotations

import ast
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

sys.path.insert(0, os.path.join(REPO, "oracle", "jax_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)


# --------------------------------------------------------------------------- #
# part 1: the reference's literal golden vectors                               #
# --------------------------------------------------------------------------- #

ENGINE_OPS = ("kernel_map", "offset_kernel_map", "kernel_scan", "offset_kernel_scan")


def _mat(*args):
    return np.arange(1, int(np.prod(args)) + 1, dtype=np.int32).reshape(*args)


def extract_engine_vectors():
    path = os.path.join(REF, "tests", "test_scan_and_map.py")
    src = open(path).read()
    tree = ast.parse(src)
    seg = lambda node: ast.get_source_segment(src, node)
    import jax.numpy as jnp  # the shim

    cases = []
    for fn in tree.body:
        if not (isinstance(fn, ast.FunctionDef) and fn.name.startswith("test_")):
            continue
        env: dict[str, str] = {}
        pending = None
        for st in fn.body:
            if isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name):
                name = st.targets[0].id
                if name == "result":
                    call = st.value  # OP(f0, in_dim, k, s, border, relative)(array)
                    inner = call.func
                    op = inner.func.id
                    assert op in ENGINE_OPS, op
                    names = [seg(a) for a in inner.args]
                    pending = {
                        "op": op,
                        "func_map": env[names[0]],
                        "shape": env[names[1]],
                        "kernel_size": env[names[2]],
                        "strides": env[names[3]],
                        "border": env[names[4]],
                        "relative": env[names[5]],
                        "array": env[seg(call.args[0])],
                        "line": st.lineno,
                    }
                else:
                    env[name] = seg(st.value)
            elif isinstance(st, ast.Expr) and isinstance(st.value, ast.Call) and \
                    getattr(st.value.func, "id", "") == "assert_array_equal":
                assert pending is not None
                pred = eval(env["pred"], {"jnp": jnp, "mat": _mat, "np": np})
                pred = np.asarray(pred)
                case = dict(pending)
                for key in ("shape", "kernel_size", "strides", "relative"):
                    case[key] = eval(case[key])
                case["border"] = eval(case["border"])
                case["expect"] = {"shape": list(pred.shape), "data": pred.astype(np.float64).ravel().tolist()}
                case["ref"] = f"tests/test_scan_and_map.py:{case.pop('line')}"
                case["id"] = f"{fn.name}:{len(cases)}"
                cases.append(case)
                pending = None
    return cases


# --------------------------------------------------------------------------- #
# part 2: the reference's own code executed on the shim                        #
# --------------------------------------------------------------------------- #


def shim_cases():
    import catalog
    import kernex as kex  # the real reference package, on the shim
    import jax.numpy as jnp

    rng = np.random.default_rng(20261019)
    meta, arrays = [], {}

    def add(kind, spec, inputs, out):
        i = len(meta)
        spec = dict(spec, kind=kind, id=i, out_dtype=str(np.asarray(out).dtype))
        meta.append(spec)
        arrays[f"out_{i}"] = np.asarray(out)
        for name, val in inputs.items():
            arrays[f"{name}_{i}"] = np.asarray(val)

    def rand_array(shape, dtype):
        if dtype == "int32":
            return rng.integers(-9, 10, size=shape).astype(np.int32)
        return rng.standard_normal(shape).astype(np.float32)

    # -- engine factories over random geometry ------------------------------ #
    for ndim in (1, 2, 3):
        for trial in range(14 if ndim < 3 else 8):
            shape = tuple(int(v) for v in rng.integers(4, 9 if ndim < 3 else 6, size=ndim))
            ksize = tuple(int(rng.integers(1, min(4, n) + 1)) for n in shape)
            strides = tuple(int(rng.integers(1, 3)) for _ in shape)
            border = tuple((int(rng.integers(-1, 3)), int(rng.integers(-1, 3))) for _ in shape)
            # keep at least one window per axis
            if any((n + lo + hi - k) // s + 1 <= 0 for n, k, s, (lo, hi) in zip(shape, ksize, strides, border)):
                continue
            relative = bool(rng.integers(0, 2))
            dtype = "int32" if trial % 2 == 0 else "float32"
            x = rand_array(shape, dtype)
            for fname in catalog.for_kernel(ksize, relative):
                fn, extra = catalog.build(fname, ksize, relative, rng)
                spec = dict(shape=shape, kernel_size=ksize, strides=strides, border=border,
                            relative=relative, fn=fname, dtype=dtype)
                args = [jnp.array(v) for v in extra]
                out = kex.kernel_map({fn: ()}, shape, ksize, strides, border, relative)(jnp.array(x), *args)
                add("kernel_map", spec, dict(x=x, **{f"arg{j}": v for j, v in enumerate(extra)}), out)
                if catalog.is_scalar(fname):
                    out = kex.kernel_scan({fn: ()}, shape, ksize, strides, border, relative)(jnp.array(x), *args)
                    add("kernel_scan", spec, dict(x=x, **{f"arg{j}": v for j, v in enumerate(extra)}), out)

    # -- offset variants ------------------------------------------------------ #
    for ndim in (1, 2, 3):
        for trial in range(8):
            shape = tuple(int(v) for v in rng.integers(5, 9 if ndim < 3 else 7, size=ndim))
            ksize = tuple(int(rng.integers(1, 4)) for _ in shape)
            strides = (1,) * ndim
            offset = tuple((int(rng.integers(0, 3)), int(rng.integers(0, 3))) for _ in shape)
            relative = bool(rng.integers(0, 2))
            dtype = "int32" if trial % 2 == 0 else "float32"
            x = rand_array(shape, dtype)
            for fname in catalog.for_kernel(ksize, relative):
                if not catalog.is_scalar(fname):
                    continue
                fn, extra = catalog.build(fname, ksize, relative, rng)
                args = [jnp.array(v) for v in extra]
                spec = dict(shape=shape, kernel_size=ksize, strides=strides, offset=offset,
                            relative=relative, fn=fname, dtype=dtype)
                ins = dict(x=x, **{f"arg{j}": v for j, v in enumerate(extra)})
                out = kex.offset_kernel_map({fn: ()}, shape, ksize, strides, offset, relative)(jnp.array(x), *args)
                add("offset_kernel_map", spec, ins, out)
                out = kex.offset_kernel_scan({fn: ()}, shape, ksize, strides, offset, relative)(jnp.array(x), *args)
                add("offset_kernel_scan", spec, ins, out)

    # -- multi-function region maps / scans through the interface classes ---- #
    for name in catalog.MESH_SCENARIOS:
        for inplace in (False, True):
            build = catalog.MESH_SCENARIOS[name]
            cls = kex.kscan if inplace else kex.kmap
            F, x = build(cls, rng)
            out = F(jnp.array(x))
            add("mesh", dict(scenario=name, inplace=inplace), dict(x=x), out)

    return meta, arrays


def main():
    vectors = extract_engine_vectors()
    with open(os.path.join(HERE, "engine_vectors.json"), "w") as f:
        json.dump(vectors, f, indent=1)
    print(f"engine_vectors.json: {len(vectors)} vectors")
    meta, arrays = shim_cases()
    np.savez_compressed(os.path.join(HERE, "shim_cases.npz"), **arrays)
    with open(os.path.join(HERE, "shim_cases.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(f"shim_cases: {len(meta)} cases")


if __name__ == "__main__":
    main()
