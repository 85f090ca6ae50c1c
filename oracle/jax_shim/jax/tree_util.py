from __future__ import annotations

import functools as _ft

Partial = _ft.partial


def tree_leaves(tree):
    if isinstance(tree, (tuple, list)):
        out = []
        for t in tree:
            out.extend(tree_leaves(t))
        return out
    if isinstance(tree, dict):
        out = []
        for t in tree.values():
            out.extend(tree_leaves(t))
        return out
    return [] if tree is None else [tree]
