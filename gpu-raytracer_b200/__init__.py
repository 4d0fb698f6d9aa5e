"""B200-native wavefront path tracer: hand-written sm_100a kernels behind a thin C ABI (include/ptb.h),
driven from a host façade shaped like the reference's Pathtracer/Integrator entry points."""
from . import build  # noqa: F401

__all__ = ["build", "scene", "pathtracer"]
