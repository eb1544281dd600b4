"""NanoVDB ingest: thin ctypes view of the host-only `lb_nanovdb_*` entry points of liblumen_b200.so.

`NanoVdbGrid(path)` reads a .vndb / .nvdb file with the library (no GPU needed) the way the reference's
`PTVolume::Load` does through `nanovdb::io::readGrid` (LumenPT/src/Framework/PTVolume.cpp:93-98) and exposes the grid's
meta data, single-voxel lookups and the dense box of values; `create_volume(renderer)` is `LumenRenderer::CreateVolume`."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import api

CLASS_UNKNOWN, CLASS_LEVEL_SET, CLASS_FOG_VOLUME = 0, 1, 2


class NanoVdbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code


class NanoVdbGrid:
    def __init__(self, source, grid_index: int = 0, bindings: Optional[api.Bindings] = None):
        """source: a file path, or the bytes of a file."""
        if bindings is None:
            from . import bindings as _b
            bindings = _b()
        self.b = bindings
        self._h = C.c_void_p()
        if isinstance(source, (bytes, bytearray, memoryview)):
            raw = bytes(source)
            rc = self.b.nanovdb_open_memory(raw, len(raw), grid_index, C.byref(self._h))
        else:
            rc = self.b.nanovdb_open(os.fsencode(source), grid_index, C.byref(self._h))
        if rc != 0:
            raise NanoVdbError(rc, (self.b.nanovdb_last_error() or b"").decode())
        i = api.LbNanoVdbInfo()
        self.b.nanovdb_info(self._h, C.byref(i))
        self.info = {"grid_type": i.grid_type, "grid_class": i.grid_class, "version": tuple(i.version), "codec": i.codec, "grid_count": i.grid_count,
                     "node_count": tuple(i.node_count), "index_min": np.array(i.index_min, np.int32), "index_max": np.array(i.index_max, np.int32),
                     "world_min": np.array(i.world_min), "world_max": np.array(i.world_max), "voxel_size": np.array(i.voxel_size),
                     "map_matrix": np.array(i.map_matrix).reshape(3, 3), "map_translation": np.array(i.map_translation),
                     "active_voxels": i.active_voxels, "grid_bytes": i.grid_bytes, "background": i.background,
                     "value_min": i.value_min, "value_max": i.value_max, "name": i.name.decode(errors="replace")}

    def close(self):
        if self._h:
            self.b.nanovdb_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def dims(self):
        """(nz, ny, nx) of the dense box."""
        d = self.info["index_max"].astype(np.int64) - self.info["index_min"] + 1
        return (int(max(d[2], 0)), int(max(d[1], 0)), int(max(d[0], 0)))

    def values(self, ijk):
        """ReadAccessor::getValue / isActive for an [n, 3] array of index coordinates."""
        c = np.ascontiguousarray(ijk, np.int32).reshape(-1, 3)
        v, a = np.zeros(len(c), np.float32), np.zeros(len(c), np.uint8)
        rc = self.b.nanovdb_values(self._h, c.ctypes.data, len(c), v.ctypes.data, a.ctypes.data)
        if rc != 0:
            raise NanoVdbError(rc, (self.b.nanovdb_last_error() or b"").decode())
        return v, a.astype(bool)

    def dense(self, as_density: bool = False) -> np.ndarray:
        out = np.zeros(self.dims, np.float32)
        rc = self.b.nanovdb_dense(self._h, 1 if as_density else 0, out.ctypes.data, out.size)
        if rc != 0:
            raise NanoVdbError(rc, (self.b.nanovdb_last_error() or b"").decode())
        return out

    def create_volume(self, renderer: api.Renderer) -> int:
        out = C.c_int32()
        rc = self.b.volume_create_nanovdb(renderer._h, self._h, C.byref(out))
        if rc != 0:
            raise NanoVdbError(rc, (self.b.nanovdb_last_error() or b"").decode())
        return out.value


def create_volume_from_file(renderer: api.Renderer, path: str) -> int:
    """LumenRenderer::CreateVolume(path) (LM/Renderer/LumenRenderer.h:168)."""
    out = C.c_int32()
    rc = renderer.b.volume_create_file(renderer._h, os.fsencode(path), C.byref(out))
    if rc != 0:
        raise NanoVdbError(rc, (renderer.b.nanovdb_last_error() or b"").decode())
    return out.value
