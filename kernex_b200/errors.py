"""Exception types of the B200 stencil engine."""


class UnsupportedStencilError(NotImplementedError):
    """The window function is outside the classes the CUDA templates cover
    (affine tap combination, window sum/mean/max/min, patch extraction).  The
    engine has no interpreter, XLA or CPU fallback by design, so this is raised
    instead (BASELINE.json north_star)."""


class EngineUnavailableError(RuntimeError):
    """libkx_b200.so is missing / not loadable, or no CUDA device is present."""
