"""Host-side mirror of the reference's ``raypier.core`` package for the hot path:
the same class and function names, so ``from raypier.core import ...`` call sites can
be pointed here unchanged.  ``tracer.trace_rays`` is the drop-in entry point."""
from . import cbezier, cdistortions, cfaces, cimplicit_surfs, cmaterials, cshapes, ctracer, obbtree  # noqa: F401
