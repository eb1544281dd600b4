"""Procedural scene descriptions for the benchmark configurations of BASELINE.json.

Everything is generated from fixed seeds with numpy, so the CUDA renderer and the CPU oracle receive byte-identical
inputs without shipping assets (the reference's Sandbox assets are not redistributable and do not exist on the GPU box):

  cornell_box()      C1: 32 triangles, 1 emissive quad, the geometry/material/camera constants of SURVEY §8d C1
  atrium()           C2: "Sponza-class" ~260 K triangles, colonnades + arches + cloth, 16 procedural 1024^2 textures,
                         instanced columns, >= 1 k override-emissive lamp triangles
  instanced_field()  C4: 64 instances of a ~156 K-triangle mesh (~10 M triangles)
  fog_room()         C3: room + homogeneous box + heterogeneous procedural density grid
  material_gallery() small scene touching every Disney lobe (parity tests)
"""
from __future__ import annotations

import numpy as np

from .api import SceneDescription, EMISSION_ENABLED, EMISSION_DISABLED, EMISSION_OVERRIDE


# ------------------------------------------------------------------ geometry helpers
def _normalize(v):
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    return v / np.maximum(n, 1e-20)


def _prim(positions, normals, uvs, tangents, indices, material):
    positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    n = positions.shape[0]
    tang = np.ascontiguousarray(tangents, np.float32).reshape(-1, 3)
    tang4 = np.concatenate([tang, np.ones((n, 1), np.float32)], axis=1)
    return {"positions": positions, "normals": np.ascontiguousarray(normals, np.float32).reshape(-1, 3),
            "uvs": np.ascontiguousarray(uvs, np.float32).reshape(-1, 2), "tangents": tang4,
            "indices": np.ascontiguousarray(indices, np.uint32).reshape(-1), "material": material}


def quad(p0, eu, ev, material, uv_scale=1.0):
    """Parallelogram p0 + s*eu + t*ev, normal = normalize(eu x ev), two triangles."""
    p0, eu, ev = (np.asarray(a, np.float64) for a in (p0, eu, ev))
    pos = np.stack([p0, p0 + eu, p0 + eu + ev, p0 + ev])
    nrm = _normalize(np.cross(eu, ev))
    return _prim(pos, np.tile(nrm, (4, 1)), np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32) * uv_scale,
                 np.tile(_normalize(eu), (4, 1)), [0, 1, 2, 0, 2, 3], material)


def merge_prims(prims, material=None):
    pos, nrm, uv, tan, idx, base = [], [], [], [], [], 0
    for p in prims:
        pos.append(p["positions"]); nrm.append(p["normals"]); uv.append(p["uvs"]); tan.append(p["tangents"])
        idx.append(p["indices"].astype(np.uint32) + base); base += p["positions"].shape[0]
    out = {"positions": np.concatenate(pos), "normals": np.concatenate(nrm), "uvs": np.concatenate(uv),
           "tangents": np.concatenate(tan), "indices": np.concatenate(idx), "material": prims[0]["material"] if material is None else material}
    return out


def parametric(fn, nu, nv, material, uv_scale=(1.0, 1.0), flip=False):
    """Tessellate fn(u, v) -> xyz over [0,1]^2 with nu x nv quads; normals/tangents from central differences."""
    u = np.linspace(0.0, 1.0, nu + 1); v = np.linspace(0.0, 1.0, nv + 1)
    U, V = np.meshgrid(u, v, indexing="xy")
    P = fn(U, V)
    h = 1e-4
    dPu = (fn(U + h, V) - fn(U - h, V)) / (2 * h)
    dPv = (fn(U, V + h) - fn(U, V - h)) / (2 * h)
    N = np.cross(dPu, dPv)
    bad = np.linalg.norm(N, axis=-1) < 1e-12
    N = _normalize(N)
    N[bad] = np.array([0.0, 1.0, 0.0])
    T = _normalize(dPu - N * np.sum(dPu * N, axis=-1, keepdims=True))
    T[np.linalg.norm(T, axis=-1) < 0.5] = np.array([1.0, 0.0, 0.0])
    if flip:
        N = -N
    i = np.arange(nu)[None, :] + (nu + 1) * np.arange(nv)[:, None]
    a, b, c, d = i, i + 1, i + nu + 2, i + nu + 1
    tri = np.stack([a, b, c, a, c, d], axis=-1) if not flip else np.stack([a, c, b, a, d, c], axis=-1)
    uv = np.stack([U * uv_scale[0], V * uv_scale[1]], axis=-1)
    return _prim(P.reshape(-1, 3), N.reshape(-1, 3), uv.reshape(-1, 2), T.reshape(-1, 3), tri.reshape(-1), material)


def cylinder(radius, height, nu, nv, material, uv_scale=(2.0, 4.0), bulge=0.0):
    def fn(U, V):
        ang = 2 * np.pi * U
        r = radius * (1.0 + bulge * np.sin(np.pi * V) + 0.04 * np.cos(12 * ang) * (bulge > 0))
        return np.stack([r * np.cos(ang), height * V, -r * np.sin(ang)], axis=-1)
    return parametric(fn, nu, nv, material, uv_scale)


def sphere(radius, nu, nv, material, center=(0, 0, 0), uv_scale=(2.0, 1.0)):
    c = np.asarray(center, np.float64)

    def fn(U, V):
        th = np.pi * (0.002 + 0.996 * V); ph = 2 * np.pi * U
        return c + radius * np.stack([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)], axis=-1)
    return parametric(fn, nu, nv, material, uv_scale, flip=True)


def box5(center, half, angle_deg, material):
    """Axis box rotated about +y, five faces (no bottom) = 10 triangles, outward normals."""
    c = np.asarray(center, np.float64); hx, hy, hz = half
    a = np.deg2rad(angle_deg); ca, sa = np.cos(a), np.sin(a)
    R = np.array([[ca, 0, sa], [0, 1, 0], [-sa, 0, ca]])
    ex, ey, ez = R @ np.array([hx, 0, 0]), np.array([0, hy, 0]), R @ np.array([0, 0, hz])
    faces = [
        (c - ex + ey + ez, 2 * ex, -2 * ez),     # top (+y)
        (c - ex - ey + ez, 2 * ex, 2 * ey),      # +z side
        (c + ex - ey - ez, -2 * ex, 2 * ey),     # -z side
        (c + ex - ey + ez, -2 * ez, 2 * ey),     # +x side
        (c - ex - ey - ez, 2 * ez, 2 * ey),      # -x side
    ]
    return merge_prims([quad(p, u, v, material) for p, u, v in faces])


def translate(x, y, z, scale=1.0, angle_y_deg=0.0):
    a = np.deg2rad(angle_y_deg); ca, sa = np.cos(a), np.sin(a)
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = np.array([[ca, 0, sa], [0, 1, 0], [-sa, 0, ca]], np.float32) * scale
    m[:3, 3] = (x, y, z)
    return m


def _mat(color, **kw):
    d = dict(diffuse_color=(*color, 1.0), metallic_factor=0.0, roughness_factor=1.0, specular_factor=0.0, luminance=1.0,
             index_of_refraction=1.0, transmittance=(0.0, 0.0, 0.0))
    d.update(kw)
    return d


# ------------------------------------------------------------------ C1: Cornell box
def cornell_box() -> SceneDescription:
    """32 triangles: 5 walls, light quad, short + tall box. Constants from SURVEY §8d C1 (the reference's CornellBox glTF)."""
    white, red, green = (0.725, 0.71, 0.68), (0.63, 0.065, 0.05), (0.14, 0.45, 0.091)
    s = SceneDescription(name="cornell")
    s.materials = [_mat(white), _mat(red), _mat(green), _mat((0.0, 0.0, 0.0), emission=(1.0, 1.0, 1.0))]
    x0, x1, y0, y1, z0, z1 = -1.02, 1.0, 0.0, 1.99, -1.04, 0.99
    walls = [
        quad((x0, y0, z1), (x1 - x0, 0, 0), (0, 0, z0 - z1), 0),     # floor, normal +y
        quad((x0, y1, z0), (x1 - x0, 0, 0), (0, 0, z1 - z0), 0),     # ceiling, normal -y
        quad((x0, y0, z0), (x1 - x0, 0, 0), (0, y1 - y0, 0), 0),     # back wall, normal +z
        quad((x0, y0, z1), (0, 0, z0 - z1), (0, y1 - y0, 0), 1),     # left wall (red), normal +x
        quad((x1, y0, z0), (0, 0, z1 - z0), (0, y1 - y0, 0), 2),     # right wall (green), normal -x
    ]
    light = quad((-0.24, 1.98, -0.22), (0.47, 0, 0), (0, 0, 0.38), 3)            # normal -y
    short = box5((0.33, 0.3, 0.33), (0.3, 0.3, 0.3), -17.0, 0)
    tall = box5((-0.34, 0.6, -0.3), (0.3, 0.6, 0.3), 18.0, 0)
    s.meshes = [[walls[0], walls[1], walls[2]], [walls[3]], [walls[4]], [light], [short], [tall]]
    s.instances = [{"mesh": i} for i in range(len(s.meshes))]
    s.camera = {"position": (0.0, 1.0, 2.0), "rotation": (0.0, 0.0, 1.0, 0.0)}
    return s


# ------------------------------------------------------------------ procedural textures
def _value_noise(rng, size, cells):
    g = rng.random((cells + 1, cells + 1))
    x = np.linspace(0, cells, size, endpoint=False)
    xi = x.astype(int); xf = x - xi; w = xf * xf * (3 - 2 * xf)
    a = g[np.ix_(xi, xi)]; b = g[np.ix_(xi, xi + 1)]; c = g[np.ix_(xi + 1, xi)]; d = g[np.ix_(xi + 1, xi + 1)]
    wx = w[None, :]; wy = w[:, None]
    return (a * (1 - wx) + b * wx) * (1 - wy) + (c * (1 - wx) + d * wx) * wy


def _fbm(rng, size, octaves=4, base=8):
    out = np.zeros((size, size)); amp = 0.5; tot = 0.0
    for o in range(octaves):
        out += amp * _value_noise(rng, size, base << o); tot += amp; amp *= 0.5
    return out / tot


def make_textures(size=1024, seed=7):
    """16 RGBA8 textures: 8 albedo (sRGB), 4 metallic-roughness, 4 normal maps."""
    rng = np.random.default_rng(seed)
    tex = []
    yy, xx = np.mgrid[0:size, 0:size]
    palettes = [((0.62, 0.55, 0.45), (0.45, 0.40, 0.33)), ((0.55, 0.25, 0.20), (0.35, 0.30, 0.28)), ((0.70, 0.68, 0.62), (0.50, 0.50, 0.48)),
                ((0.25, 0.35, 0.55), (0.65, 0.60, 0.40)), ((0.20, 0.45, 0.25), (0.50, 0.42, 0.20)), ((0.60, 0.15, 0.15), (0.75, 0.65, 0.30)),
                ((0.40, 0.40, 0.42), (0.25, 0.25, 0.27)), ((0.80, 0.75, 0.65), (0.60, 0.45, 0.35))]
    for k, (c0, c1) in enumerate(palettes):
        n = _fbm(rng, size, 4, 8)
        if k % 3 == 0:       # bricks
            row = (yy // (size // 16)); bx = (xx + (row % 2) * (size // 16)) % (size // 8); by = yy % (size // 16)
            mortar = ((bx < 6) | (by < 6)).astype(float)
            t = np.clip(0.75 * n + 0.25 - 0.6 * mortar, 0, 1)
        elif k % 3 == 1:     # tiles
            t = (((xx // (size // 8)) + (yy // (size // 8))) % 2) * 0.6 + 0.4 * n
        else:
            t = n
        rgb = np.asarray(c0)[None, None, :] * t[..., None] + np.asarray(c1)[None, None, :] * (1 - t[..., None])
        alpha = np.ones((size, size, 1))
        tex.append({"pixels": np.clip(np.concatenate([rgb, alpha], -1) * 255 + 0.5, 0, 255).astype(np.uint8), "srgb": True})
    for k in range(4):       # metallic (b) / roughness (g)
        n = _fbm(rng, size, 3, 16)
        g = np.clip(0.35 + 0.6 * n, 0.05, 1.0); b = (n > 0.55 + 0.05 * k).astype(float) if k % 2 else np.full_like(n, 1.0)
        px = np.stack([np.ones_like(n), g, b, np.ones_like(n)], -1)
        tex.append({"pixels": np.clip(px * 255 + 0.5, 0, 255).astype(np.uint8), "srgb": False})
    for k in range(4):       # normal maps from a height field
        hgt = _fbm(rng, size, 4, 16 << (k % 2))
        dx = np.roll(hgt, -1, 1) - np.roll(hgt, 1, 1); dy = np.roll(hgt, -1, 0) - np.roll(hgt, 1, 0)
        nrm = _normalize(np.stack([-dx * 24, -dy * 24, np.ones_like(hgt)], -1))
        px = np.concatenate([nrm * 0.5 + 0.5, np.ones((size, size, 1))], -1)
        tex.append({"pixels": np.clip(px * 255 + 0.5, 0, 255).astype(np.uint8), "srgb": False})
    return tex


# ------------------------------------------------------------------ C2: Sponza-class atrium
def atrium(detail: float = 1.0, texture_size: int = 1024, lamp_radiance: float = 200.0) -> SceneDescription:
    """A two-storey colonnaded atrium, ~260 K triangles at detail=1 (scales ~ detail^2). Lamps are override-emissive spheres."""
    s = SceneDescription(name="atrium")
    s.textures = make_textures(texture_size)
    alb = list(range(0, 8)); mr = list(range(8, 12)); nm = list(range(12, 16))
    rng = np.random.default_rng(11)
    mats = []
    for k in range(24):
        m = _mat((1.0, 1.0, 1.0), diffuse_texture=alb[k % 8], normal_texture=nm[k % 4],
                 roughness_factor=float(np.clip(0.25 + 0.75 * rng.random(), 0.05, 1.0)), metallic_factor=0.0,
                 specular_factor=float(0.2 + 0.6 * rng.random()), luminance=1.0)
        if k % 5 == 1:
            m.update(metallic_factor=1.0, metallic_roughness_texture=mr[k % 4])
        if k % 7 == 3:
            m.update(clear_coat_factor=0.8, clear_coat_roughness_factor=0.1)
        if k % 6 == 4:
            m.update(sheen_factor=0.6, sheen_tint_factor=0.5, subsurface_factor=0.3)
        mats.append(m)
    mats.append(_mat((0.9, 0.9, 0.9)))        # 24: lamp body (emission comes from the instance override)
    s.materials = mats
    d = lambda n: max(2, int(round(n * detail)))
    L, Wd, H = 36.0, 16.0, 14.0              # length (x), width (z), height (y)

    def height_floor(U, V):
        x = (U - 0.5) * L; z = (V - 0.5) * Wd
        return np.stack([x, 0.03 * np.sin(3.1 * x) * np.cos(2.7 * z), -z], axis=-1)
    floor = parametric(height_floor, d(160), d(72), 1, (18, 8))
    upper_floor_l = parametric(lambda U, V: np.stack([(U - 0.5) * L, np.full_like(U, 6.0), -(-Wd / 2 + V * 3.5)], -1), d(120), d(12), 6, (18, 2))
    upper_floor_r = parametric(lambda U, V: np.stack([(U - 0.5) * L, np.full_like(U, 6.0), -(Wd / 2 - 3.5 + V * 3.5)], -1), d(120), d(12), 6, (18, 2))

    def wall(p0, eu, ev, nu, nv, mat, ripple=0.05):
        p0, eu, ev = (np.asarray(a, float) for a in (p0, eu, ev)); n = _normalize(np.cross(eu, ev))
        return parametric(lambda U, V: p0 + U[..., None] * eu + V[..., None] * ev + (ripple * np.sin(9 * U * np.linalg.norm(eu) / 4) * np.sin(7 * V * np.linalg.norm(ev) / 4))[..., None] * n,
                          nu, nv, mat, (np.linalg.norm(eu) / 4, np.linalg.norm(ev) / 4))
    walls = [wall((-L / 2, 0, -Wd / 2), (L, 0, 0), (0, H, 0), d(128), d(48), 0),
             wall((L / 2, 0, Wd / 2), (-L, 0, 0), (0, H, 0), d(128), d(48), 0),
             wall((-L / 2, 0, Wd / 2), (0, 0, -Wd), (0, H, 0), d(56), d(48), 3),
             wall((L / 2, 0, -Wd / 2), (0, 0, Wd), (0, H, 0), d(56), d(48), 3)]
    # ceiling ring with a central opening (rays can escape -> miss handling)
    ceil = [wall((-L / 2, H, Wd / 2), (L, 0, 0), (0, 0, -4.0), d(96), d(10), 2, 0.0), wall((-L / 2, H, -Wd / 2 + 4.0), (L, 0, 0), (0, 0, -4.0), d(96), d(10), 2, 0.0)]
    s.meshes.append([floor, upper_floor_l, upper_floor_r] + walls + ceil)               # mesh 0: shell
    s.instances.append({"mesh": 0})

    column = cylinder(0.45, 5.6, d(56), d(40), 7, bulge=0.06)
    capital = merge_prims([box5((0, 5.8, 0), (0.65, 0.2, 0.65), 0.0, 10)])
    base = merge_prims([box5((0, 0.15, 0), (0.6, 0.15, 0.6), 0.0, 10)])
    s.meshes.append([column, capital, base])                                             # mesh 1: column (instanced)
    n_cols = 11
    for storey, y in enumerate((0.0, 6.0)):
        for side, z in enumerate((-Wd / 2 + 3.4, Wd / 2 - 3.4)):
            for k in range(n_cols):
                x = -L / 2 + 2.0 + k * (L - 4.0) / (n_cols - 1)
                inst = {"mesh": 1, "transform": translate(x, y, z, 1.0 if storey == 0 else 0.8, 13.0 * k)}
                if (k + side + storey) % 4 == 0:
                    inst["override_material"] = 5 + ((k + storey) % 3) * 6     # exercise override materials
                s.instances.append(inst)

    def arch(U, V):      # half torus spanning two columns
        R, r = (L - 4.0) / (n_cols - 1) / 2, 0.28
        a = np.pi * U; b = 2 * np.pi * V
        return np.stack([-(R + r * np.cos(b)) * np.cos(a), (R * 0.8 + r * np.cos(b)) * np.sin(a), r * np.sin(b)], -1)
    s.meshes.append([parametric(arch, d(40), d(14), 4, (4, 1))])                          # mesh 2: arch (instanced)
    for y in (5.9, 10.7):
        for z in (-Wd / 2 + 3.4, Wd / 2 - 3.4):
            for k in range(n_cols - 1):
                x = -L / 2 + 2.0 + (k + 0.5) * (L - 4.0) / (n_cols - 1)
                s.instances.append({"mesh": 2, "transform": translate(x, y, z, 1.0 if y < 6 else 0.8)})

    def cloth(phase):
        def fn(U, V):
            x = (U - 0.5) * 5.0; y = -V * 6.5
            z = 0.35 * np.sin(5.0 * U * np.pi + phase) * (0.3 + V) + 0.12 * np.sin(11 * V + 3 * phase) * np.cos(7 * U)
            return np.stack([x, y, z], -1)
        return fn
    for k in range(6):
        s.meshes.append([parametric(cloth(0.9 * k), d(110), d(70), 16 + k % 6, (3, 4))])   # meshes 3..8: drapes
        s.instances.append({"mesh": len(s.meshes) - 1, "transform": translate(-13.0 + 5.2 * k, 12.6, (-1) ** k * 1.2, 1.0, 90.0 * (k % 2))})

    vase = parametric(lambda U, V: np.stack([(0.35 + 0.25 * np.sin(np.pi * V * 1.3)) * np.cos(2 * np.pi * U), 1.4 * V, -(0.35 + 0.25 * np.sin(np.pi * V * 1.3)) * np.sin(2 * np.pi * U)], -1),
                      d(48), d(32), 11, (2, 2))
    s.meshes.append([vase]); vase_mesh = len(s.meshes) - 1
    for k in range(8):
        s.instances.append({"mesh": vase_mesh, "transform": translate(-14.0 + 4.0 * k, 0.05, (-1) ** k * 1.6, 0.9 + 0.05 * k, 40.0 * k),
                            "override_material": [9, 13, 19, 21][k % 4]})

    lamp = sphere(0.22, d(16), d(8), 24)
    s.meshes.append([lamp]); lamp_mesh = len(s.meshes) - 1
    lamp_tris = len(lamp["indices"]) // 3
    n_lamps = max(4, int(np.ceil(1024 / lamp_tris)))
    for k in range(n_lamps):
        x = -L / 2 + 3.0 + (k % 8) * (L - 6.0) / 7; y = 4.6 if (k // 8) % 2 == 0 else 10.4; z = (-1) ** (k // 16) * (1.8 + 0.4 * ((k // 8) % 2))
        s.instances.append({"mesh": lamp_mesh, "transform": translate(x, y, z), "emission_mode": EMISSION_OVERRIDE,
                            "override_radiance": (lamp_radiance, lamp_radiance * 0.92, lamp_radiance * 0.8), "emission_scale": 1.0})
    s.camera = {"position": (-15.5, 2.2, 0.6), "rotation": _quat_y(-100.0)}
    return s


def _quat_y(deg):
    a = np.deg2rad(deg) / 2
    return (float(np.cos(a)), 0.0, float(np.sin(a)), 0.0)


# ------------------------------------------------------------------ C4: instanced stress scene
def instanced_field(instances_per_side: int = 8, mesh_res: int = 280) -> SceneDescription:
    """instances_per_side^2 copies of a displaced sphere with 2*mesh_res^2 triangles (8, 280 -> 64 x 156 800 = 10.0 M)."""
    s = SceneDescription(name="instanced_field")
    rng = np.random.default_rng(3)
    s.materials = [_mat((0.7, 0.7, 0.7), roughness_factor=0.6, specular_factor=0.5), _mat((0.75, 0.6, 0.3), metallic_factor=1.0, roughness_factor=0.35),
                   _mat((0.3, 0.5, 0.8), roughness_factor=0.8), _mat((0.5, 0.5, 0.5))]

    def blob(U, V):
        th = np.pi * (0.002 + 0.996 * V); ph = 2 * np.pi * U
        r = 1.0 + 0.12 * np.sin(9 * ph) * np.sin(7 * th) + 0.05 * np.sin(31 * ph + 2.0) * np.sin(23 * th)
        return r[..., None] * np.stack([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)], -1)
    s.meshes.append([parametric(blob, mesh_res, mesh_res, 0, (8, 4), flip=True)])
    n = instances_per_side; spacing = 3.2
    for i in range(n):
        for j in range(n):
            jit = rng.random(3)
            s.instances.append({"mesh": 0, "transform": translate((i - (n - 1) / 2) * spacing + 0.5 * jit[0], 1.2 + 0.6 * jit[1], (j - (n - 1) / 2) * spacing + 0.5 * jit[2],
                                                                 0.9 + 0.3 * jit[1], 360.0 * jit[2]), "override_material": (i * n + j) % 3})
    ext = n * spacing
    s.meshes.append([parametric(lambda U, V: np.stack([(U - 0.5) * ext * 1.5, np.zeros_like(U), -(V - 0.5) * ext * 1.5], -1), 64, 64, 3, (16, 16))])
    s.instances.append({"mesh": 1})
    s.meshes.append([sphere(0.5, 24, 12, 3)])
    for k in range(16):
        s.instances.append({"mesh": 2, "transform": translate(((k % 4) - 1.5) * ext / 4, 5.0, ((k // 4) - 1.5) * ext / 4), "emission_mode": EMISSION_OVERRIDE,
                            "override_radiance": (120.0, 110.0, 100.0), "emission_scale": 1.0})
    s.camera = {"position": (0.0, 6.5, ext * 0.62), "rotation": (0.0, 0.0, 1.0, 0.0)}
    return s


# ------------------------------------------------------------------ C3: room with participating media
def fog_room(grid: int = 64) -> SceneDescription:
    s = cornell_box()
    s.name = "fog_room"
    for m in s.meshes:                       # C1 geometry x10
        for p in m:
            p["positions"] = p["positions"] * 10.0
    s.materials[3]["emission"] = (40.0, 40.0, 40.0)
    s.camera = {"position": (0.0, 10.0, 20.0), "rotation": (0.0, 0.0, 1.0, 0.0)}
    z, y, x = np.mgrid[0:grid, 0:grid, 0:grid].astype(np.float32) / (grid - 1) - 0.5
    r = np.sqrt(x * x + y * y + z * z)
    dens = np.clip(1.0 - r / 0.5, 0, 1) * (0.6 + 0.4 * np.sin(18 * x) * np.sin(15 * y + 1.0) * np.sin(13 * z + 2.0))
    s.volumes = [
        {"density": None, "bbox_min": (-9.0, 0.5, -9.0), "bbox_max": (-2.0, 6.0, -2.0), "instance_density": 0.5},
        {"density": np.clip(dens, 0, 1).astype(np.float32), "bbox_min": (1.0, 1.0, -8.0), "bbox_max": (9.0, 9.0, 0.0), "instance_density": 2.0},
    ]
    return s


# ------------------------------------------------------------------ material gallery (parity of every lobe through the full pipeline)
def material_gallery() -> SceneDescription:
    s = SceneDescription(name="gallery")
    rng = np.random.default_rng(5)
    tex = np.clip(rng.random((16, 16, 4)) * 255, 0, 255).astype(np.uint8); tex[..., 3] = 255
    cut = tex.copy(); cut[::2, ::2, 3] = 0
    nrm = np.zeros((8, 8, 4), np.uint8); nrm[..., 0] = 128 + rng.integers(-30, 30, (8, 8)); nrm[..., 1] = 128 + rng.integers(-30, 30, (8, 8)); nrm[..., 2] = 240; nrm[..., 3] = 255
    s.textures = [{"pixels": tex, "srgb": True}, {"pixels": cut, "srgb": True}, {"pixels": nrm, "srgb": False}, {"pixels": tex, "srgb": False}]
    s.materials = [
        _mat((0.8, 0.8, 0.8)),
        _mat((0.9, 0.6, 0.3), metallic_factor=1.0, roughness_factor=0.3),
        _mat((0.3, 0.6, 0.9), specular_factor=0.8, roughness_factor=0.4, specular_tint_factor=0.5, tint_factor=(0.9, 0.7, 0.5)),
        _mat((0.7, 0.2, 0.2), clear_coat_factor=1.0, clear_coat_roughness_factor=0.2, roughness_factor=0.7, specular_factor=0.3),
        _mat((0.5, 0.7, 0.5), sheen_factor=0.8, sheen_tint_factor=0.4, subsurface_factor=0.5, roughness_factor=0.9),
        _mat((0.95, 0.95, 0.95), transmission_factor=0.9, index_of_refraction=1.5, roughness_factor=0.15, transmittance=(0.1, 0.2, 0.3), specular_factor=0.5),
        _mat((1.0, 1.0, 1.0), diffuse_texture=0, normal_texture=2, metallic_roughness_texture=3, roughness_factor=0.9, specular_factor=0.4),
        _mat((1.0, 1.0, 1.0), diffuse_texture=1, roughness_factor=0.8),                      # alpha cut-out
        _mat((0.0, 0.0, 0.0), emission=(6.0, 5.5, 5.0)),
        _mat((0.6, 0.6, 0.65), anisotropic=0.7, metallic_factor=0.8, roughness_factor=0.35, specular_factor=0.6),
    ]
    room = [quad((-4, 0, 3), (8, 0, 0), (0, 0, -6), 6, 4.0), quad((-4, 0, -3), (8, 0, 0), (0, 4, 0), 0), quad((-4, 0, 3), (0, 0, -6), (0, 4, 0), 2), quad((4, 0, -3), (0, 0, 6), (0, 4, 0), 4),
            quad((-4, 4, -3), (8, 0, 0), (0, 0, 6), 0)]
    s.meshes.append(room); s.instances.append({"mesh": 0})
    s.meshes.append([quad((-1.0, 3.95, -1.0), (2.0, 0, 0), (0, 0, 2.0), 8)]); s.instances.append({"mesh": 1})
    ball = sphere(0.55, 24, 12, 0)
    s.meshes.append([ball])
    for k, m in enumerate([1, 3, 5, 9, 6, 4]):
        s.instances.append({"mesh": 2, "transform": translate(-3.0 + 1.2 * k, 0.6, -0.8 + 0.3 * (k % 2)), "override_material": m})
    s.meshes.append([quad((-1.5, 0.2, 1.2), (3.0, 0, 0), (0, 1.6, 0), 7, 2.0)]); s.instances.append({"mesh": 3})
    s.instances.append({"mesh": 2, "transform": translate(2.8, 2.6, -1.5, 0.5), "emission_mode": EMISSION_OVERRIDE, "override_radiance": (20.0, 4.0, 2.0), "emission_scale": 0.5})
    s.camera = {"position": (0.0, 1.8, 5.2), "rotation": (0.0, 0.0, 1.0, 0.0)}
    return s


SCENES = {"cornell": cornell_box, "atrium": atrium, "instanced_field": instanced_field, "fog_room": fog_room, "gallery": material_gallery}


def real_sponza(path):
    """The reference's own Sponza asset (Sandbox/assets/models/Sponza/Sponza.gltf — not part of this repository) through the native ingest,
    with the lamps and the camera of profiles/sponza_real.py: Sponza has no emissive material, so — as SURVEY 8d C2 prescribes — eight
    override-emissive spheres light it; the reference renders the asset unscaled (~3 700 units long), lamps and camera are placed in those
    units. Returns (scene, camera position, camera rotation, glTF info, seconds spent loading and decoding)."""
    import time
    from . import api
    from .gltf import GltfDocument
    t0 = time.time()
    with GltfDocument(path) as doc:
        info = dict(doc.info); scene = doc.to_scene_description()
    t_load = time.time() - t0
    lamp_mat = len(scene.materials)
    scene.materials.append(dict(diffuse_color=(0.9, 0.9, 0.9, 1.0), metallic_factor=0.0, roughness_factor=1.0, luminance=1.0, index_of_refraction=1.0))
    scene.meshes.append([sphere(22.0, 16, 8, lamp_mat)]); lamp_mesh = len(scene.meshes) - 1
    for k in range(8):
        x = -1000.0 + 650.0 * (k % 4); y = 350.0 if k < 4 else 800.0; z = 120.0 if k % 2 else -120.0
        scene.instances.append({"mesh": lamp_mesh, "transform": translate(x, y, z), "emission_mode": api.EMISSION_OVERRIDE,
                                "override_radiance": (4000.0, 3700.0, 3200.0), "emission_scale": 1.0})
    return scene, (-1150.0, 250.0, 20.0), _quat_y(90.0), info, t_load           # in the nave, looking along +x
