"""glTF ingest: thin ctypes view of the host-only `lb_gltf_*` entry points of liblumen_b200.so.

`GltfDocument(path)` parses a .gltf / .glb with the library (no GPU needed) and exposes what the reference's converter
(LumenPT/src/Tools/LumenPTModelConverter.cpp) would hand to CreateTexture / CreateMaterial / CreatePrimitive / AddMesh.
`to_scene_description()` repackages it as a `SceneDescription`, which uploads through any `api.Renderer`."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import api

_MATERIAL_FIELDS = ("transmission_factor", "clear_coat_factor", "clear_coat_roughness_factor", "index_of_refraction", "specular_factor",
                    "specular_tint_factor", "subsurface_factor", "luminance", "anisotropic", "sheen_factor", "sheen_tint_factor",
                    "metallic_factor", "roughness_factor")
_TEXTURE_FIELDS = ("diffuse_texture", "normal_texture", "metallic_roughness_texture", "emissive_texture", "transmission_texture",
                   "clear_coat_texture", "clear_coat_roughness_texture", "tint_texture")


class GltfError(RuntimeError):
    pass


class GltfDocument:
    def __init__(self, path: str, bindings: Optional[api.Bindings] = None, image_decoder=None, cached: bool = False):
        if bindings is None:
            from . import bindings as _b
            bindings = _b()
        self.b = bindings
        self._h = C.c_void_p()
        self._decoder = api.IMAGE_DECODE_FN(image_decoder) if image_decoder is not None else None       # keep the thunk alive
        rc = (self.b.gltf_open_cached if cached else self.b.gltf_open)(os.fsencode(path), C.cast(self._decoder, C.c_void_p) if self._decoder else None, None, C.byref(self._h))
        if rc != 0:
            raise GltfError(f"[{rc}] {(self.b.gltf_last_error() or b'').decode()}")
        info = api.LbGltfInfo()
        self.b.gltf_info(self._h, C.byref(info))
        self.info = {n: getattr(info, n) for n, _ in api.LbGltfInfo._fields_}

    def close(self):
        if self._h:
            self.b.gltf_close(self._h)
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

    def image(self, i: int) -> dict:
        p, w, h, srgb, dec = C.POINTER(C.c_uint8)(), C.c_uint32(), C.c_uint32(), C.c_int(), C.c_int()
        if self.b.gltf_image(self._h, i, C.byref(p), C.byref(w), C.byref(h), C.byref(srgb), C.byref(dec)) != 0:
            raise GltfError("image index")
        px = np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy()
        return {"pixels": px, "srgb": bool(srgb.value), "decoded": bool(dec.value)}

    def material(self, i: int) -> dict:
        d = api.LbMaterialDesc()
        if self.b.gltf_material(self._h, i, C.byref(d)) != 0:
            raise GltfError("material index")
        out = {"diffuse_color": tuple(d.diffuse_color), "emission": tuple(d.emission), "tint_factor": tuple(d.tint_factor), "transmittance": tuple(d.transmittance)}
        out.update({f: getattr(d, f) for f in _MATERIAL_FIELDS})
        out.update({f: getattr(d, f) for f in _TEXTURE_FIELDS})
        return out

    def primitives(self, mesh: int) -> list:
        n = C.c_uint32()
        if self.b.gltf_mesh_primitive_count(self._h, mesh, C.byref(n)) != 0:
            raise GltfError("mesh index")
        out = []
        for p in range(n.value):
            d = api.LbPrimitiveDesc()
            self.b.gltf_primitive(self._h, mesh, p, C.byref(d))
            nv = d.vertex_count

            def arr(ptr, width):
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(nv, width)).copy() if nv else np.zeros((0, width), np.float32)
            idx = np.ctypeslib.as_array(C.cast(d.indices, C.POINTER(C.c_uint32)), shape=(d.index_count,)).copy() if d.index_count else np.zeros(0, np.uint32)
            out.append({"positions": arr(d.positions, 3), "uvs": arr(d.uvs, 2), "normals": arr(d.normals, 3), "tangents": arr(d.tangents, 4),
                        "indices": idx, "material": d.material})
        return out

    def instance(self, i: int) -> dict:
        mesh, m = C.c_uint32(), np.zeros(16, np.float32)
        if self.b.gltf_instance(self._h, i, C.byref(mesh), m.ctypes.data) != 0:
            raise GltfError("instance index")
        return {"mesh": mesh.value, "transform": m.reshape(4, 4)}

    def save_ollad(self, path: str) -> None:
        """lb_gltf_save_ollad: the reference converter's `.ollad` cache file of this document (GltfDocument(path.ollad) reads it back)."""
        if self.b.gltf_save_ollad(self._h, os.fsencode(path)) != 0:
            raise GltfError((self.b.gltf_last_error() or b"").decode())

    def to_scene_description(self) -> api.SceneDescription:
        s = api.SceneDescription(name="gltf")
        images = [self.image(i) for i in range(self.info["images"])]
        s.textures = [{"pixels": im["pixels"], "srgb": im["srgb"]} for im in images]
        for i in range(self.info["materials"]):
            m = self.material(i)
            for f in _TEXTURE_FIELDS:
                if m[f] >= 0 and not images[m[f]]["decoded"]:
                    m[f] = -1
            s.materials.append(m)
        default = None
        for mi in range(self.info["meshes"]):
            prims = self.primitives(mi)
            for p in prims:
                if p["material"] < 0:
                    if default is None:
                        default = len(s.materials)
                        s.materials.append(dict(diffuse_color=(1.0, 1.0, 1.0, 1.0), metallic_factor=1.0, roughness_factor=1.0, luminance=1.0, index_of_refraction=1.0))
                    p["material"] = default
            s.meshes.append(prims)
        s.instances = [self.instance(i) for i in range(self.info["instances"])]
        return s

    def upload(self, renderer: api.Renderer, root_transform=None):
        """lb_gltf_upload: the whole document through the renderer's own create/add calls (product library only)."""
        first, count = C.c_int32(), C.c_uint32()
        root = np.ascontiguousarray(root_transform, np.float32).reshape(16) if root_transform is not None else None
        rc = self.b.gltf_upload(renderer._h, self._h, root.ctypes.data if root is not None else None, C.byref(first), C.byref(count))
        if rc != 0:
            raise GltfError(f"[{rc}] {(self.b.gltf_last_error() or b'').decode()}")
        return first.value, count.value
