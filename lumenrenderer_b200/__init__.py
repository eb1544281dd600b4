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


# ---- multi-GPU inside the library (csrc/lb_multigpu.cpp, include/lumen_b200.h "multi-GPU inside the library")
def _prefer_bundled_nccl():
    """The library loads NCCL at run time by soname. When this interpreter also has torch (which bundles a newer libnccl.so.2), the first copy
    mapped serves both — so name the bundled file before the library's first NCCL call (LB_NCCL_LIB, csrc/lb_multigpu.cpp). torch is not imported."""
    if os.environ.get("LB_NCCL_LIB"):
        return
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        path = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(path):
            os.environ["LB_NCCL_LIB"] = path
            return


def _mg(code: int):
    if code != 0:
        raise LumenError(code, (bindings().multigpu_last_error() or b"").decode())


def comm_unique_id() -> bytes:
    """128-byte NCCL id: create on one rank, hand to the others (any launcher), pass to Renderer.comm_init on every rank."""
    _prefer_bundled_nccl()
    buf = (ctypes.c_uint8 * 128)()
    _mg(bindings().comm_unique_id(buf))
    return bytes(buf)


def band_settings(settings: Settings, rank: int, ranks: int):
    """lb_band_settings: settings of the renderer producing band `rank` (+ halo) of the frame `settings` describes, and its owned rows (y0, y1)."""
    full, out, y0, y1 = settings.to_c(), api.LbSettings(), ctypes.c_uint32(), ctypes.c_uint32()
    _mg(bindings().band_settings(ctypes.byref(full), rank, ranks, ctypes.byref(out), ctypes.byref(y0), ctypes.byref(y1)))
    return Settings.from_c(out), (y0.value, y1.value)


def shard_settings(settings: Settings, rank: int, ranks: int) -> Settings:
    base, out = settings.to_c(), api.LbSettings()
    _mg(bindings().shard_settings(ctypes.byref(base), rank, ranks, ctypes.byref(out)))
    return Settings.from_c(out)


class Group:
    """One process driving n GPUs (lb_group_*): n member renderers + one NCCL communicator (ncclCommInitAll). mode = "samples" (disjoint
    frameCount streams, one reduce of the accumulation buffers) or "bands" (row bands of one frame with a ReSTIR halo, one gather per frame)."""

    def __init__(self, devices, settings: Settings, mode: str = "samples"):
        _prefer_bundled_nccl()
        self.b = bindings()
        self._g = ctypes.c_void_p()
        dev = (ctypes.c_int * len(devices))(*devices)
        cs = settings.to_c()
        _mg(self.b.group_create(dev, len(devices), ctypes.byref(cs), {"samples": 0, "bands": 1}[mode], ctypes.byref(self._g)))
        self.width, self.height, self.mode = settings.width, settings.height, mode
        self.members = []
        for i in range(len(devices)):
            h = ctypes.c_void_p()
            _mg(self.b.group_member(self._g, i, ctypes.byref(h)))
            m = api.Renderer.__new__(api.Renderer)              # a view of the member: the group owns and destroys it
            m.b, m._h, m._borrowed = self.b, h, True
            m.settings = m.get_settings(); m.width, m.height = m.settings.width, m.settings.height
            self.members.append(m)

    def load_scene(self, scene):
        for m in self.members:
            m.load_scene(scene)

    def set_camera(self, position, rotation):
        for m in self.members:
            m.set_camera(position, rotation)

    def render(self, frames: int = 1):
        _mg(self.b.group_render(self._g, frames))

    def reduce(self):
        _mg(self.b.group_reduce(self._g))

    def reset(self):
        _mg(self.b.group_reset(self._g))

    def synchronize(self):
        _mg(self.b.group_synchronize(self._g))

    def read_hdr(self):
        import numpy as np
        out = np.empty((self.height, self.width, 4), np.float32)
        _mg(self.b.group_read_hdr(self._g, out.ctypes.data, out.nbytes))
        return out

    def close(self):
        if self._g:
            for m in self.members:
                m._h = ctypes.c_void_p()
            self.b.group_destroy(self._g)
            self._g = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def version() -> str:
    return bindings().version().decode()
