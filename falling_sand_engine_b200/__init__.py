"""B200-native falling-sand world tick (fse).

Host-side mirror of the reference's `world` tick API (source/engine/world.hpp:148-192) over the
C ABI in include/fse.h.  The CUDA library is loaded on first use of `World`/`Context`; there is no
CPU fallback — using the API without the built extension or without a GPU raises.
"""
from . import types, worldgen  # noqa: F401

__all__ = ["types", "worldgen", "Context", "World", "FseError", "load_library", "default_materials"]

# fse_pixels_read / fse_pixels_device plane ids (include/fse.h)
PIXELS_MAIN, PIXELS_FIRE, PIXELS_EMISSION, PIXELS_FLOW, PIXELS_LAYER2, PIXELS_BACKGROUND = range(6)


def __getattr__(name):
    if name in ("Context", "World", "FseError", "load_library", "default_materials"):
        from . import api

        return getattr(api, name)
    raise AttributeError(name)
