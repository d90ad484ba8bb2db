"""drake_b200: B200-native back end for Drake's hydroelastic contact-surface path.

The product is the CUDA shared library built from drake_b200/csrc (C-ABI in
include/drake_b200_hydro.h); `engine` is its thin ctypes host mirror.
"""
from . import build  # noqa: F401
from .engine import (COMPLIANT, POLYGON, RIGID, STRICT, TRIANGLE, WITH_FALLBACK, ContactSurface, Context,  # noqa: F401
                     HydroError, Mesh, PackedResults, SceneBatch)
