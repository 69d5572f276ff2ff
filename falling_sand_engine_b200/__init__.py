"""B200-native falling-sand world tick (fse).

Host-side mirror of the reference's `world` tick API (source/engine/world.hpp:148-192) over the
C ABI in include/fse.h.  The CUDA library is loaded on first use of `World`/`Context`; there is no
CPU fallback — using the API without the built extension or without a GPU raises.
"""
from . import types, worldgen  # noqa: F401

__all__ = ["types", "worldgen", "Context", "World", "load_library", "default_materials"]


def __getattr__(name):
    if name in ("Context", "World", "load_library", "default_materials"):
        from . import api

        return getattr(api, name)
    raise AttributeError(name)
