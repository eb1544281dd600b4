"""lumenrenderer_b200 — B200-native wavefront path-tracing core behind the LumenPT renderer interface.

The product is `liblumen_b200.so` (C ABI in include/lumen_b200.h, CUDA sm_100a, built from csrc/). This package is the
thin host-side mirror of the reference's `LumenRenderer` interface over that ABI (api.py) plus procedural scene
descriptions for the benchmark configurations (scenes.py). There is no CPU fallback: constructing a `Renderer` without
the CUDA library raises, and the library itself refuses to start without an sm_100 device.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

from . import api
from .api import (MaterialData, Settings, SceneDescription, LumenError, Bindings, C_ABI_SYMBOLS, HIT_DTYPE,
                  EMISSION_ENABLED, EMISSION_DISABLED, EMISSION_OVERRIDE,
                  CHANNEL_DIRECT, CHANNEL_INDIRECT, CHANNEL_SPECULAR, CHANNEL_VOLUMETRIC,
                  VOLUME_COMPAT, VOLUME_DELTA, SURF_EMISSIVE, SURF_ALPHA, SURF_MISS, pack_material24)

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# LUMEN_B200_LIB selects another build of the same library (A/B experiments of compile flags); never the oracle.
LIB_PATH = os.environ.get("LUMEN_B200_LIB") or os.path.join(_PKG_DIR, "liblumen_b200.so")
_bindings = None


def build_native(verbose: bool = False) -> str:
    """Compile csrc/ into liblumen_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-j8", "-C", os.path.join(_PKG_DIR, "csrc")], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building liblumen_b200.so failed")
    return LIB_PATH


def bindings() -> Bindings:
    """Typed entry points of the CUDA library. Raises when the library has not been built: there is no fallback."""
    global _bindings
    if _bindings is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the renderer has no CPU or PyTorch fallback path)")
        _bindings = Bindings(ctypes.CDLL(LIB_PATH), "lb_")
    return _bindings


class Renderer(api.Renderer):
    """The B200 renderer: `api.Renderer` bound to liblumen_b200.so."""

    def __init__(self, settings: Settings | None = None, **kwargs):
        super().__init__(bindings(), settings if settings is not None else Settings(**kwargs))


def version() -> str:
    return bindings().version().decode()
