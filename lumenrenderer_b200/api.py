"""ctypes binding of the C ABI in include/lumen_b200.h and the host-side mirror of the reference interface.

`Renderer` follows `LumenRenderer` / `WaveFront::WaveFrontRenderer`
(/root/reference/Lumen_Engine/Lumen/src/Lumen/Renderer/LumenRenderer.h:37-218,
/root/reference/Lumen_Engine/LumenPT/src/Framework/WaveFrontRenderer.h:50-269): the same verbs
(CreateTexture, CreateMaterial, CreatePrimitive, CreateMesh, AddMesh, SetRenderResolution, SetBlendMode,
GetOutputTexturePixels ...) are available under their reference names next to snake_case ones.

The class is generic over (shared library, symbol prefix) so that the test-suite can drive the CPU oracle
through the very same code; the product only ever instantiates it with its own CUDA library
(`lumenrenderer_b200.Renderer`), and importing this module never touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field, fields
from typing import Optional, Sequence

import numpy as np

LB_OK = 0
EMISSION_ENABLED, EMISSION_DISABLED, EMISSION_OVERRIDE = 0, 1, 2
CHANNEL_DIRECT, CHANNEL_INDIRECT, CHANNEL_SPECULAR, CHANNEL_VOLUMETRIC = 0, 1, 2, 3
VOLUME_COMPAT, VOLUME_DELTA = 0, 1
SURF_EMISSIVE, SURF_ALPHA, SURF_MISS = 1, 2, 4


class LbSettings(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("depth", C.c_uint32), ("blend_output", C.c_uint32),
                ("restir", C.c_uint32), ("restir_temporal", C.c_uint32), ("restir_spatial", C.c_uint32),
                ("device", C.c_int32), ("volume_mode", C.c_uint32), ("first_frame_count", C.c_uint32),
                ("frame_count_stride", C.c_uint32), ("band_row0", C.c_uint32), ("band_full_height", C.c_uint32), ("restir_unbiased", C.c_uint32), ("band_own_row0", C.c_uint32), ("band_own_rows", C.c_uint32)]


class LbMaterialDesc(C.Structure):
    _fields_ = [("diffuse_color", C.c_float * 4), ("emission", C.c_float * 3), ("transmission_factor", C.c_float),
                ("clear_coat_factor", C.c_float), ("clear_coat_roughness_factor", C.c_float),
                ("index_of_refraction", C.c_float), ("specular_factor", C.c_float), ("specular_tint_factor", C.c_float),
                ("subsurface_factor", C.c_float), ("luminance", C.c_float), ("anisotropic", C.c_float),
                ("sheen_factor", C.c_float), ("sheen_tint_factor", C.c_float), ("metallic_factor", C.c_float),
                ("roughness_factor", C.c_float), ("tint_factor", C.c_float * 3), ("transmittance", C.c_float * 3),
                ("diffuse_texture", C.c_int32), ("normal_texture", C.c_int32), ("metallic_roughness_texture", C.c_int32),
                ("emissive_texture", C.c_int32), ("transmission_texture", C.c_int32), ("clear_coat_texture", C.c_int32),
                ("clear_coat_roughness_texture", C.c_int32), ("tint_texture", C.c_int32)]


class LbPrimitiveDesc(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("position_stride", C.c_uint32),
                ("uvs", C.c_void_p), ("uv_stride", C.c_uint32),
                ("normals", C.c_void_p), ("normal_stride", C.c_uint32),
                ("tangents", C.c_void_p), ("tangent_stride", C.c_uint32),
                ("vertex_count", C.c_uint32), ("indices", C.c_void_p), ("index_size", C.c_uint32),
                ("index_count", C.c_uint32), ("material", C.c_int32)]


class LbEmissiveness(C.Structure):
    _fields_ = [("mode", C.c_int32), ("override_radiance", C.c_float * 3), ("scale", C.c_float)]


class LbVolumeDesc(C.Structure):
    _fields_ = [("density", C.c_void_p), ("nx", C.c_uint32), ("ny", C.c_uint32), ("nz", C.c_uint32),
                ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3)]


@dataclass
class MaterialData:
    """LumenRenderer::MaterialData with the reference defaults (LumenRenderer.h:66-82)."""
    diffuse_color: Sequence[float] = (1.0, 1.0, 1.0, 1.0)
    emission: Sequence[float] = (0.0, 0.0, 0.0)
    transmission_factor: float = 0.0
    clear_coat_factor: float = 0.0
    clear_coat_roughness_factor: float = 0.0
    index_of_refraction: float = 1.0
    specular_factor: float = 0.0
    specular_tint_factor: float = 0.0
    subsurface_factor: float = 0.0
    luminance: float = 1.0
    anisotropic: float = 0.0
    sheen_factor: float = 0.0
    sheen_tint_factor: float = 0.0
    metallic_factor: float = 1.0
    roughness_factor: float = 1.0
    tint_factor: Sequence[float] = (1.0, 1.0, 1.0)
    transmittance: Sequence[float] = (1.0, 1.0, 1.0)
    diffuse_texture: int = -1
    normal_texture: int = -1
    metallic_roughness_texture: int = -1
    emissive_texture: int = -1
    transmission_texture: int = -1
    clear_coat_texture: int = -1
    clear_coat_roughness_texture: int = -1
    tint_texture: int = -1

    def to_c(self) -> LbMaterialDesc:
        d = LbMaterialDesc()
        for name, _ in LbMaterialDesc._fields_:
            v = getattr(self, name)
            if isinstance(v, (tuple, list, np.ndarray)):
                arr = getattr(d, name)
                for i, x in enumerate(v):
                    arr[i] = float(x)
            else:
                setattr(d, name, v)
        return d


@dataclass
class Settings:
    """WaveFrontSettings (WaveFrontRenderer.h:31-48) + ReSTIR toggles (ReSTIRData.h:58-65)."""
    width: int = 1280
    height: int = 720
    depth: int = 5
    blend_output: bool = False
    restir: bool = True
    restir_temporal: bool = True
    restir_spatial: bool = True
    device: int = 0
    volume_mode: int = VOLUME_COMPAT
    first_frame_count: int = 0
    frame_count_stride: int = 0
    band_row0: int = 0
    band_full_height: int = 0
    restir_unbiased: bool = False
    band_own_row0: int = 0
    band_own_rows: int = 0

    @classmethod
    def from_c(cls, c: "LbSettings") -> "Settings":
        return cls(**{f.name: type(getattr(cls, f.name))(getattr(c, f.name)) for f in fields(cls)})

    def to_c(self) -> LbSettings:
        s = LbSettings()
        for name, _ in LbSettings._fields_:
            if name != "reserved":
                setattr(s, name, int(getattr(self, name)))
        return s


class LumenError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code


_SIGS = {
    "create": [C.POINTER(LbSettings), C.POINTER(C.c_void_p)],
    "destroy": [C.c_void_p],
    "texture_create": [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_int32)],
    "material_create": [C.c_void_p, C.POINTER(LbMaterialDesc), C.POINTER(C.c_int32)],
    "material_update": [C.c_void_p, C.c_int32, C.POINTER(LbMaterialDesc)],
    "primitive_create": [C.c_void_p, C.POINTER(LbPrimitiveDesc), C.POINTER(C.c_int32)],
    "mesh_create": [C.c_void_p, C.POINTER(C.c_int32), C.c_uint32, C.POINTER(C.c_int32)],
    "volume_create": [C.c_void_p, C.POINTER(LbVolumeDesc), C.POINTER(C.c_int32)],
    "scene_add_mesh_instance": [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(LbEmissiveness), C.c_int32, C.POINTER(C.c_int32)],
    "instance_set_transform": [C.c_void_p, C.c_int32, C.c_void_p],
    "instance_set_emissiveness": [C.c_void_p, C.c_int32, C.POINTER(LbEmissiveness)],
    "instance_set_override_material": [C.c_void_p, C.c_int32, C.c_int32],
    "scene_add_volume_instance": [C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.POINTER(C.c_int32)],
    "scene_clear": [C.c_void_p],
    "camera_set_pose": [C.c_void_p, C.c_void_p, C.c_void_p],
    "camera_set_matrix": [C.c_void_p, C.c_void_p],
    "hdr_buffer": [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)],
    "camera_set_fov_y": [C.c_void_p, C.c_float],
    "set_render_resolution": [C.c_void_p, C.c_uint32, C.c_uint32],
    "get_render_resolution": [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
    "get_settings": [C.c_void_p, C.POINTER(LbSettings)],
    "set_depth": [C.c_void_p, C.c_uint32],
    "set_blend_mode": [C.c_void_p, C.c_int],
    "get_blend_mode": [C.c_void_p, C.POINTER(C.c_int)],
    "reset_history": [C.c_void_p],
    "render_frames": [C.c_void_p, C.c_uint32],
    "synchronize": [C.c_void_p],
    "start_rendering": [C.c_void_p],
    "stop_rendering": [C.c_void_p],
    "read_hdr": [C.c_void_p, C.c_void_p, C.c_size_t],
    "read_hdr_async": [C.c_void_p, C.c_void_p, C.c_size_t],
    "read_gbuffer": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t],
    "camera_set_min_max_distance": [C.c_void_p, C.c_float, C.c_float],
    "readback_wait": [C.c_void_p],
    "read_ldr": [C.c_void_p, C.c_void_p, C.c_size_t],
    "read_channel": [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t],
    "read_motion_vectors": [C.c_void_p, C.c_void_p, C.c_size_t],
    "frame_stats": [C.c_void_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)],
    "frame_counters": [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)],
    "save_png": [C.c_void_p, C.c_char_p],
    "frame_stats_json": [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)],
    "accum_buffer": [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)],
    "resolve_accum": [C.c_void_p, C.c_uint32],
    "set_overlap": [C.c_void_p, C.c_int],
    "set_stream": [C.c_void_p, C.c_void_p],
    "debug_trace_closest": [C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_void_p],
    "debug_trace_any": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_void_p],
    "debug_read_lights": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)],
    "debug_read_primary_hits": [C.c_void_p, C.c_void_p, C.c_size_t],
    "debug_read_surface": [C.c_void_p, C.c_void_p, C.c_size_t],
    "debug_read_reservoirs": [C.c_void_p, C.c_void_p, C.c_size_t],
    "debug_eval_bsdf": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p],
    "debug_sample_bsdf": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p],
}
# Host-only entry points (asset ingest): no renderer, no device work, nothing for the oracle to mirror.
class LbGltfInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("images", "undecoded_images", "materials", "meshes", "primitives", "instances", "triangles", "vertices")]


class LbNanoVdbInfo(C.Structure):
    _fields_ = [("grid_type", C.c_uint32), ("grid_class", C.c_uint32), ("version", C.c_uint32 * 3), ("codec", C.c_uint32), ("grid_count", C.c_uint32),
                ("node_count", C.c_uint32 * 4), ("index_min", C.c_int32 * 3), ("index_max", C.c_int32 * 3),
                ("world_min", C.c_double * 3), ("world_max", C.c_double * 3), ("voxel_size", C.c_double * 3),
                ("map_matrix", C.c_double * 9), ("map_translation", C.c_double * 3),
                ("active_voxels", C.c_uint64), ("grid_bytes", C.c_uint64),
                ("background", C.c_float), ("value_min", C.c_float), ("value_max", C.c_float), ("name", C.c_char * 256)]


IMAGE_DECODE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p)
_HOST_SIGS = {
    "gltf_open": [C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)],
    "gltf_open_cached": [C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)],
    "gltf_close": [C.c_void_p],
    "gltf_save_ollad": [C.c_void_p, C.c_char_p],
    "gltf_info": [C.c_void_p, C.POINTER(LbGltfInfo)],
    "gltf_image": [C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "gltf_material": [C.c_void_p, C.c_uint32, C.POINTER(LbMaterialDesc)],
    "gltf_mesh_primitive_count": [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)],
    "gltf_primitive": [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(LbPrimitiveDesc)],
    "gltf_instance": [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p],
    "gltf_upload": [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)],
    "nanovdb_open": [C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)],
    "nanovdb_open_memory": [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_void_p)],
    "nanovdb_close": [C.c_void_p],
    "nanovdb_info": [C.c_void_p, C.POINTER(LbNanoVdbInfo)],
    "nanovdb_values": [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p],
    "nanovdb_dense": [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t],
    "volume_create_nanovdb": [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)],
    "volume_create_file": [C.c_void_p, C.c_char_p, C.POINTER(C.c_int32)],
    # multi-GPU inside the library (csrc/lb_multigpu.cpp): NCCL over NVLink, no oracle counterpart (the CPU tests use gloo for the same exchange)
    "get_stream": [C.c_void_p, C.POINTER(C.c_void_p)],
    "reduce_begin": [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)],
    "reduce_end": [C.c_void_p, C.c_int, C.c_uint32],
    "reduce_wait": [C.c_void_p],
    "band_settings": [C.POINTER(LbSettings), C.c_uint32, C.c_uint32, C.POINTER(LbSettings), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
    "shard_settings": [C.POINTER(LbSettings), C.c_uint32, C.c_uint32, C.POINTER(LbSettings)],
    "comm_unique_id": [C.c_void_p],
    "comm_init": [C.c_void_p, C.c_void_p, C.c_int, C.c_int],
    "comm_reduce_accum": [C.c_void_p, C.c_int, C.c_uint32],
    "comm_gather_bands": [C.c_void_p, C.c_int, C.c_void_p],
    "comm_destroy": [C.c_void_p],
    "group_create": [C.POINTER(C.c_int), C.c_uint32, C.POINTER(LbSettings), C.c_int, C.POINTER(C.c_void_p)],
    "group_size": [C.c_void_p, C.POINTER(C.c_uint32)],
    "group_member": [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)],
    "group_render": [C.c_void_p, C.c_uint32],
    "group_reduce": [C.c_void_p],
    "group_reset": [C.c_void_p],
    "group_read_hdr": [C.c_void_p, C.c_void_p, C.c_size_t],
    "group_synchronize": [C.c_void_p],
    "group_destroy": [C.c_void_p],
}
HOST_ONLY_SYMBOLS = tuple(_HOST_SIGS) + ("gltf_last_error", "nanovdb_last_error", "multigpu_last_error")
C_ABI_SYMBOLS = tuple(_SIGS) + ("last_error", "version")

HIT_DTYPE = np.dtype([("instance", np.uint32), ("primitive", np.uint32), ("u", np.float32), ("v", np.float32), ("t", np.float32)])


class Bindings:
    """Typed access to one shared library exporting the lumen_b200.h entry points under `prefix`."""

    def __init__(self, lib: C.CDLL, prefix: str = "lb_"):
        self.lib, self.prefix = lib, prefix
        for name, args in _SIGS.items():
            fn = getattr(lib, prefix + name)
            fn.argtypes, fn.restype = args, C.c_int
            setattr(self, name, fn)
        self.last_error = getattr(lib, prefix + "last_error")
        self.last_error.restype = C.c_char_p
        self.version = getattr(lib, prefix + "version")
        self.version.restype = C.c_char_p
        if prefix == "lb_":                                   # the product library also carries the host-only asset ingest
            for name, args in _HOST_SIGS.items():
                fn = getattr(lib, prefix + name)
                fn.argtypes, fn.restype = args, C.c_int
                setattr(self, name, fn)
            self.gltf_last_error = lib.lb_gltf_last_error
            self.gltf_last_error.restype = C.c_char_p
            self.nanovdb_last_error = lib.lb_nanovdb_last_error
            self.nanovdb_last_error.restype = C.c_char_p
            self.multigpu_last_error = lib.lb_multigpu_last_error
            self.multigpu_last_error.restype = C.c_char_p

    def check(self, code: int):
        if code != LB_OK:
            raise LumenError(code, (self.last_error() or b"").decode())


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def pack_material24(m: MaterialData, ior_is_eta: Optional[float] = None) -> np.ndarray:
    """Packed material for the debug BSDF taps: color4, transmittance3+ior, tint3+luminance, 11 byte-quantised params."""
    ior = m.index_of_refraction if ior_is_eta is None else ior_is_eta
    return np.array([*m.diffuse_color, *m.transmittance, ior, *m.tint_factor, m.luminance,
                     m.metallic_factor, m.subsurface_factor, m.specular_factor, m.roughness_factor,
                     m.specular_tint_factor, m.anisotropic, m.sheen_factor, m.sheen_tint_factor,
                     m.clear_coat_factor, 1.0 - m.clear_coat_roughness_factor, m.transmission_factor, 0.0], dtype=np.float32)


class Renderer:
    """Host-side mirror of LumenRenderer/WaveFrontRenderer over the C ABI."""

    def __init__(self, bindings: Bindings, settings: Settings):
        self.b = bindings
        self.settings = settings
        self._h = C.c_void_p()
        cs = settings.to_c()
        self.b.check(self.b.create(C.byref(cs), C.byref(self._h)))
        self.width, self.height = settings.width, settings.height

    # ---- lifetime
    def close(self):
        if self._h and not getattr(self, "_borrowed", False):          # a member of a lumenrenderer_b200.Group belongs to the group
            self.b.destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- resources (reference names kept as aliases)
    def create_texture(self, rgba8: np.ndarray, srgb: bool = False) -> int:
        px = np.ascontiguousarray(rgba8, dtype=np.uint8)
        if px.ndim != 3 or px.shape[2] != 4:
            raise ValueError("texture must be HxWx4 uint8")
        out = C.c_int32()
        self.b.check(self.b.texture_create(self._h, px.ctypes.data, px.shape[1], px.shape[0], int(srgb), C.byref(out)))
        return out.value

    def create_material(self, m: MaterialData) -> int:
        out = C.c_int32()
        d = m.to_c()
        self.b.check(self.b.material_create(self._h, C.byref(d), C.byref(out)))
        return out.value

    def update_material(self, handle: int, m: MaterialData):
        d = m.to_c()
        self.b.check(self.b.material_update(self._h, handle, C.byref(d)))

    def create_primitive(self, positions, indices, material: int, uvs=None, normals=None, tangents=None) -> int:
        pos = _f32(positions, (-1, 3))
        idx = np.ascontiguousarray(indices)
        if idx.dtype not in (np.uint16, np.uint32):
            idx = idx.astype(np.uint32)
        idx = idx.reshape(-1)
        d = LbPrimitiveDesc()
        keep = [pos, idx]
        d.positions, d.position_stride = pos.ctypes.data, 12
        if uvs is not None:
            uv = _f32(uvs, (-1, 2)); keep.append(uv); d.uvs, d.uv_stride = uv.ctypes.data, 8
        if normals is not None:
            nr = _f32(normals, (-1, 3)); keep.append(nr); d.normals, d.normal_stride = nr.ctypes.data, 12
        if tangents is not None:
            tg = _f32(tangents, (-1, 4)); keep.append(tg); d.tangents, d.tangent_stride = tg.ctypes.data, 16
        d.vertex_count = pos.shape[0]
        d.indices, d.index_size, d.index_count = idx.ctypes.data, idx.dtype.itemsize, idx.size
        d.material = material
        out = C.c_int32()
        self.b.check(self.b.primitive_create(self._h, C.byref(d), C.byref(out)))
        return out.value

    def create_primitive_interleaved(self, vertices48: np.ndarray, indices, material: int) -> int:
        """Interleaved reference `Vertex` (pos3, uv2, normal3, tangent4 = 48 bytes), ModelStructs.h:21-28."""
        v = _f32(vertices48, (-1, 12))
        idx = np.ascontiguousarray(indices).reshape(-1)
        if idx.dtype not in (np.uint16, np.uint32):
            idx = idx.astype(np.uint32)
        d = LbPrimitiveDesc()
        base = v.ctypes.data
        d.positions, d.position_stride = base, 48
        d.uvs, d.uv_stride = base + 12, 48
        d.normals, d.normal_stride = base + 20, 48
        d.tangents, d.tangent_stride = base + 32, 48
        d.vertex_count = v.shape[0]
        d.indices, d.index_size, d.index_count = idx.ctypes.data, idx.dtype.itemsize, idx.size
        d.material = material
        out = C.c_int32()
        self.b.check(self.b.primitive_create(self._h, C.byref(d), C.byref(out)))
        return out.value

    def create_mesh(self, primitives: Sequence[int]) -> int:
        arr = (C.c_int32 * len(primitives))(*primitives)
        out = C.c_int32()
        self.b.check(self.b.mesh_create(self._h, arr, len(primitives), C.byref(out)))
        return out.value

    def create_volume(self, density: Optional[np.ndarray], bbox_min, bbox_max) -> int:
        d = LbVolumeDesc()
        keep = None
        if density is not None:
            keep = _f32(density)
            if keep.ndim != 3:
                raise ValueError("density must be [nz, ny, nx]")
            d.density, d.nz, d.ny, d.nx = keep.ctypes.data, keep.shape[0], keep.shape[1], keep.shape[2]
        for i in range(3):
            d.bbox_min[i], d.bbox_max[i] = float(bbox_min[i]), float(bbox_max[i])
        out = C.c_int32()
        self.b.check(self.b.volume_create(self._h, C.byref(d), C.byref(out)))
        return out.value

    @staticmethod
    def _emissiveness(mode=EMISSION_ENABLED, radiance=(0.0, 0.0, 0.0), scale=1.0) -> LbEmissiveness:
        e = LbEmissiveness()
        e.mode = mode
        for i in range(3):
            e.override_radiance[i] = float(radiance[i])
        e.scale = float(scale)
        return e

    def add_mesh_instance(self, mesh: int, transform=None, emission_mode=EMISSION_ENABLED, override_radiance=(0.0, 0.0, 0.0),
                          emission_scale=1.0, override_material: int = -1) -> int:
        m = _f32(np.eye(4) if transform is None else transform, (16,))
        e = self._emissiveness(emission_mode, override_radiance, emission_scale)
        out = C.c_int32()
        self.b.check(self.b.scene_add_mesh_instance(self._h, mesh, m.ctypes.data, C.byref(e), override_material, C.byref(out)))
        return out.value

    def set_instance_transform(self, instance: int, transform):
        m = _f32(transform, (16,))
        self.b.check(self.b.instance_set_transform(self._h, instance, m.ctypes.data))

    def set_instance_emissiveness(self, instance: int, mode, radiance=(0.0, 0.0, 0.0), scale=1.0):
        e = self._emissiveness(mode, radiance, scale)
        self.b.check(self.b.instance_set_emissiveness(self._h, instance, C.byref(e)))

    def set_instance_override_material(self, instance: int, material: int):
        self.b.check(self.b.instance_set_override_material(self._h, instance, material))

    def add_volume_instance(self, volume: int, transform=None, density: float = 0.001) -> int:
        m = _f32(np.eye(4) if transform is None else transform, (16,))
        out = C.c_int32()
        self.b.check(self.b.scene_add_volume_instance(self._h, volume, m.ctypes.data, density, C.byref(out)))
        return out.value

    def clear_scene(self):
        self.b.check(self.b.scene_clear(self._h))

    # ---- camera / settings
    def set_camera(self, position, rotation_wxyz=(1.0, 0.0, 0.0, 0.0), fov_y: Optional[float] = None):
        p, q = _f32(position, (3,)), _f32(rotation_wxyz, (4,))
        self.b.check(self.b.camera_set_pose(self._h, p.ctypes.data, q.ctypes.data))
        if fov_y is not None:
            self.b.check(self.b.camera_set_fov_y(self._h, fov_y))

    def set_camera_matrix(self, world_row_major):
        """Camera world matrix (columns right / up / forward / position), row-major: Camera::GetMatrixData, Camera.cpp:95-104."""
        m = _f32(world_row_major, (16,))
        self.b.check(self.b.camera_set_matrix(self._h, m.ctypes.data))

    def set_render_resolution(self, width: int, height: int):
        self.b.check(self.b.set_render_resolution(self._h, width, height))
        self.width, self.height = width, height

    def get_render_resolution(self):
        w, h = C.c_uint32(), C.c_uint32()
        self.b.check(self.b.get_render_resolution(self._h, C.byref(w), C.byref(h)))
        return w.value, h.value

    def set_depth(self, depth: int):
        self.b.check(self.b.set_depth(self._h, depth))

    def set_blend_mode(self, blend: bool):
        self.b.check(self.b.set_blend_mode(self._h, int(blend)))

    def get_blend_mode(self) -> bool:
        v = C.c_int()
        self.b.check(self.b.get_blend_mode(self._h, C.byref(v)))
        return bool(v.value)

    def reset_history(self):
        self.b.check(self.b.reset_history(self._h))

    # ---- hot path
    def render_frames(self, frames: int = 1):
        self.b.check(self.b.render_frames(self._h, frames))

    def synchronize(self):
        self.b.check(self.b.synchronize(self._h))

    def start_rendering(self):
        self.b.check(self.b.start_rendering(self._h))

    def stop_rendering(self):
        self.b.check(self.b.stop_rendering(self._h))

    # ---- outputs
    def read_hdr(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        self.b.check(self.b.read_hdr(self._h, out.ctypes.data, out.nbytes))
        return out

    def read_hdr_into(self, ptr: int, nbytes: int):
        """Read-back into caller-owned (e.g. pinned) host memory."""
        self.b.check(self.b.read_hdr(self._h, ptr, nbytes))

    def read_hdr_async(self, ptr: int, nbytes: int):
        """Enqueue the read-back of the frame just rendered into (pinned) host memory; it overlaps the next frame. Pair with readback_wait()."""
        self.b.check(self.b.read_hdr_async(self._h, ptr, nbytes))

    def readback_wait(self):
        self.b.check(self.b.readback_wait(self._h))

    def read_gbuffer(self):
        """(depth [H, W], normal_roughness [H, W, 4], albedo [H, W, 4]) of the frame just rendered."""
        n = self.width * self.height
        depth = np.empty((self.height, self.width), np.float32)
        nr = np.empty((self.height, self.width, 4), np.float32); al = np.empty_like(nr)
        self.b.check(self.b.read_gbuffer(self._h, depth.ctypes.data, nr.ctypes.data, al.ctypes.data, n))
        return depth, nr, al

    def set_camera_min_max_distance(self, min_distance: float, max_distance: float):
        self.b.check(self.b.camera_set_min_max_distance(self._h, min_distance, max_distance))

    def read_ldr(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), np.uint8)
        self.b.check(self.b.read_ldr(self._h, out.ctypes.data, out.nbytes))
        return out

    def read_channel(self, channel: int) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), np.float32)
        self.b.check(self.b.read_channel(self._h, channel, out.ctypes.data, out.nbytes))
        return out

    def read_motion_vectors(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 2), np.float32)
        self.b.check(self.b.read_motion_vectors(self._h, out.ctypes.data, out.nbytes))
        return out

    def frame_stats(self) -> dict:
        names, n = C.c_char_p(), C.c_uint32()
        vals = np.zeros(64, np.float32)
        self.b.check(self.b.frame_stats(self._h, C.byref(names), vals.ctypes.data, 64, C.byref(n)))
        keys = (names.value or b"").decode().split(";") if n.value else []
        out: dict = {}
        for k, v in zip(keys, vals[: n.value]):
            out[k] = out.get(k, 0.0) + float(v)
        return out

    def frame_counters(self) -> dict:
        vals, n = np.zeros(16, np.uint64), C.c_uint32()
        self.b.check(self.b.frame_counters(self._h, vals.ctypes.data, 16, C.byref(n)))
        keys = ["extend_rays", "shadow_rays", "visibility_rays", "kernel_launches", "lights", "triangles", "bvh_nodes", "bvh_bytes",
                "bvh_build_us", "bvh_levels", "bvh_build_rounds", "stack_overflows", "bvh_refit_us", "bvh_refits", "bvh_alloc_us"]
        return {k: int(v) for k, v in zip(keys, vals[: n.value])}

    def save_png(self, path: str):
        """Screenshot of the 8-bit output (OutputLayer::MakeScreenshot, Sandbox/src/OutputLayer.cpp:882-896)."""
        self.b.check(self.b.save_png(self._h, os.fsencode(path)))

    def frame_stats_json(self) -> dict:
        """FrameStats of the last frame as parsed JSON (times in microseconds per stage + counters)."""
        import json
        need = C.c_size_t(0)
        self.b.frame_stats_json(self._h, None, 0, C.byref(need))
        buf = C.create_string_buffer(need.value)
        self.b.check(self.b.frame_stats_json(self._h, buf, need.value, None))
        return json.loads(buf.value.decode())

    def hdr_buffer(self):
        """(pointer, bytes) of the merged fp32 RGBA frame — a device pointer for the CUDA library, host memory for the oracle."""
        p, n = C.c_void_p(), C.c_size_t()
        self.b.check(self.b.hdr_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def accum_buffer(self):
        p, nbytes, frames = C.c_void_p(), C.c_size_t(), C.c_uint32()
        self.b.check(self.b.accum_buffer(self._h, C.byref(p), C.byref(nbytes), C.byref(frames)))
        return p.value, nbytes.value, frames.value

    def resolve_accum(self, total_frames: int):
        self.b.check(self.b.resolve_accum(self._h, total_frames))

    def set_overlap(self, mode: int):
        """Overlap mask: bit 0 = bounce-wave shadow rays under the next extend launch, bit 1 = ReSTIR chain beside all bounce waves,
        bit 2 = late bounce waves beside the ReSTIR chain (default 5); 0 = every launch serialised (exclusive stage times)."""
        self.b.check(self.b.set_overlap(self._h, int(mode)))

    def set_stream(self, cuda_stream: int):
        self.b.check(self.b.set_stream(self._h, C.c_void_p(cuda_stream)))

    def get_settings(self) -> Settings:
        c = LbSettings()
        self.b.check(self.b.get_settings(self._h, C.byref(c)))
        return Settings.from_c(c)

    # ---- multi-GPU, one rank per process (CUDA library only; csrc/lb_multigpu.cpp)
    def _mg(self, code: int):
        if code != LB_OK:
            raise LumenError(code, (self.b.multigpu_last_error() or b"").decode())

    def comm_init(self, unique_id: bytes, rank: int, ranks: int):
        """Joins the NCCL communicator described by the 128-byte id of `comm_unique_id()` (created on one rank, handed round by the launcher)."""
        import lumenrenderer_b200 as _pkg
        _pkg._prefer_bundled_nccl()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._mg(self.b.comm_init(self._h, buf, rank, ranks))

    def comm_reduce_accum(self, root: int, total_frames: int):
        """Sample sharding: the one collective — sum-reduce of the fp32 accumulation buffers onto `root`, which resolves sum / total_frames."""
        self._mg(self.b.comm_reduce_accum(self._h, root, total_frames))

    def reduce_wait(self):
        """Makes the renderer's stream wait for the reduce started by comm_reduce_accum (which otherwise overlaps the following frames)."""
        self.b.check(self.b.reduce_wait(self._h))

    def comm_gather_bands(self, root: int, full_frame_device_ptr: int = 0):
        """Row bands: every rank's owned rows to `root` (device to device); `full_frame_device_ptr` = H x W float4 on the root."""
        self._mg(self.b.comm_gather_bands(self._h, root, C.c_void_p(full_frame_device_ptr)))

    def comm_destroy(self):
        self._mg(self.b.comm_destroy(self._h))

    # ---- debug taps
    def trace_closest(self, origins, directions, tmin=0.01, tmax=5000.0) -> np.ndarray:
        rays = np.ascontiguousarray(np.concatenate([_f32(origins, (-1, 3)), _f32(directions, (-1, 3))], axis=1))
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        self.b.check(self.b.debug_trace_closest(self._h, rays.ctypes.data, rays.shape[0], tmin, tmax, hits.ctypes.data))
        return hits

    def trace_any(self, origins, directions, tmax, tmin=0.01) -> np.ndarray:
        rays = np.ascontiguousarray(np.concatenate([_f32(origins, (-1, 3)), _f32(directions, (-1, 3))], axis=1))
        tm = _f32(tmax, (-1,))
        occ = np.empty(rays.shape[0], np.uint8)
        self.b.check(self.b.debug_trace_any(self._h, rays.ctypes.data, tm.ctypes.data, rays.shape[0], tmin, occ.ctypes.data))
        return occ

    def read_lights(self):
        n = C.c_uint32()
        self.b.debug_read_lights(self._h, None, None, 0, C.byref(n))
        lights, cdf = np.empty((n.value, 16), np.float32), np.empty(n.value, np.float32)
        if n.value:
            self.b.check(self.b.debug_read_lights(self._h, lights.ctypes.data, cdf.ctypes.data, n.value, C.byref(n)))
        return lights, cdf

    def read_primary_hits(self) -> np.ndarray:
        out = np.empty(self.width * self.height, HIT_DTYPE)
        self.b.check(self.b.debug_read_primary_hits(self._h, out.ctypes.data, out.nbytes))
        return out.reshape(self.height, self.width)

    def read_surface(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 24), np.float32)
        self.b.check(self.b.debug_read_surface(self._h, out.ctypes.data, out.nbytes))
        return out

    def read_reservoirs(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 20), np.float32)
        self.b.check(self.b.debug_read_reservoirs(self._h, out.ctypes.data, out.nbytes))
        return out

    def eval_bsdf(self, mat24, normal, tangent, wo, wi) -> np.ndarray:
        v = np.ascontiguousarray(np.concatenate([_f32(normal, (-1, 3)), _f32(tangent, (-1, 3)), _f32(wo, (-1, 3)), _f32(wi, (-1, 3))], axis=1))
        m = _f32(mat24, (24,))
        out = np.empty((v.shape[0], 4), np.float32)
        self.b.check(self.b.debug_eval_bsdf(self._h, m.ctypes.data, v.ctypes.data, v.shape[0], out.ctypes.data))
        return out

    def sample_bsdf(self, mat24, normal, tangent, wo, r012) -> np.ndarray:
        v = np.ascontiguousarray(np.concatenate([_f32(normal, (-1, 3)), _f32(tangent, (-1, 3)), _f32(wo, (-1, 3)), _f32(r012, (-1, 3))], axis=1))
        m = _f32(mat24, (24,))
        out = np.empty((v.shape[0], 8), np.float32)
        self.b.check(self.b.debug_sample_bsdf(self._h, m.ctypes.data, v.ctypes.data, v.shape[0], out.ctypes.data))
        return out

    # ---- scene description upload (see scenes.py)
    def load_scene(self, scene: "SceneDescription"):
        tex = [self.create_texture(t["pixels"], t.get("srgb", False)) for t in scene.textures]
        mats = []
        for m in scene.materials:
            mm = MaterialData(**{k: v for k, v in m.items() if not k.endswith("_texture")})
            for k, v in m.items():
                if k.endswith("_texture") and v is not None and v >= 0:
                    setattr(mm, k, tex[v])
            mats.append(self.create_material(mm))
        meshes = []
        for mesh in scene.meshes:
            prims = [self.create_primitive(p["positions"], p["indices"], mats[p["material"]], p.get("uvs"), p.get("normals"), p.get("tangents")) for p in mesh]
            meshes.append(self.create_mesh(prims))
        for inst in scene.instances:
            self.add_mesh_instance(meshes[inst["mesh"]], inst.get("transform"), inst.get("emission_mode", EMISSION_ENABLED),
                                   inst.get("override_radiance", (0, 0, 0)), inst.get("emission_scale", 1.0),
                                   mats[inst["override_material"]] if inst.get("override_material", -1) >= 0 else -1)
        for v in scene.volumes:
            h = self.create_volume(v.get("density"), v["bbox_min"], v["bbox_max"])
            self.add_volume_instance(h, v.get("transform"), v.get("instance_density", 0.001))
        if scene.camera is not None:
            self.set_camera(scene.camera["position"], scene.camera.get("rotation", (1, 0, 0, 0)), scene.camera.get("fov_y"))

    # reference-named aliases (LumenRenderer.h:151-200)
    CreateTexture = create_texture
    CreateMaterial = create_material
    CreatePrimitive = create_primitive
    CreateMesh = create_mesh
    CreateVolume = create_volume
    AddMesh = add_mesh_instance
    AddVolume = add_volume_instance
    SetRenderResolution = set_render_resolution
    GetRenderResolution = get_render_resolution
    SetBlendMode = set_blend_mode
    GetBlendMode = get_blend_mode
    StartRendering = start_rendering
    TraceFrame = render_frames
    GetLastFrameStats = frame_stats

    def GetOutputTexturePixels(self):
        px = self.read_ldr()
        return px, self.width, self.height


@dataclass
class SceneDescription:
    textures: list = field(default_factory=list)       # {"pixels": HxWx4 u8, "srgb": bool}
    materials: list = field(default_factory=list)      # MaterialData kwargs, *_texture = index into textures
    meshes: list = field(default_factory=list)         # list of primitives: {"positions","indices","uvs","normals","tangents","material"}
    instances: list = field(default_factory=list)      # {"mesh","transform","emission_mode","override_radiance","emission_scale"}
    volumes: list = field(default_factory=list)
    camera: Optional[dict] = None
    name: str = "scene"

    def triangle_count(self) -> int:
        per_mesh = [sum(len(np.asarray(p["indices"]).reshape(-1)) // 3 for p in m) for m in self.meshes]
        return sum(per_mesh[i["mesh"]] for i in self.instances)
