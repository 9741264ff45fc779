"""pyhalma_b200: B200-native (sm_100a) replacement for pyHALMA's direct-sum potential and
unbinding hot path.  Importing the package does not touch the GPU; the first call into
libhalma_unbind does, and raises if the CUDA library or a device is missing."""
from . import _lib  # noqa: F401
from .particle import particle  # noqa: F401
from .unbind import (G_CONST, CatalogueResult, UnbindPlan, UnbindResult, unbind_catalogue,  # noqa: F401
                     unbind_halo)

__all__ = ["particle", "UnbindPlan", "unbind_halo", "unbind_catalogue", "G_CONST", "CatalogueResult",
           "UnbindResult"]
