"""Test infrastructure for the glTF ingest.

* `write_gltf` — serialises a small scene graph to .gltf (data URIs or external .bin/.png) or .glb.
* `load_reference_semantics` — an independent numpy/json restatement of what the reference's converter
  (LumenPT/src/Tools/LumenPTModelConverter.cpp: materials :347-531, LoadBinary :1027-1059, GenerateTangentBinary :734-900,
  nodes :953-1025 + :275-317) turns a document into; the C++ loader of the product is compared with it bit for bit.
Neither is used by the product."""
import base64
import json
import os
import struct
import zlib

import numpy as np

F = np.float32


# ----------------------------------------------------------------------------------------------------------------- PNG (writer)
def png_bytes(rgba: np.ndarray) -> bytes:
    h, w = rgba.shape[:2]
    raw = b"".join(b"\x00" + rgba[y].astype(np.uint8).tobytes() for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")


# ----------------------------------------------------------------------------------------------------------------- writer
def write_gltf(path, meshes, materials, nodes, roots, images=(), flavour="embedded"):
    """meshes: list of lists of primitives {"positions","normals"?, "uvs"?, "tangents"?, "indices", "index_type"? (np dtype), "material"};
    materials: glTF material dicts (already in glTF vocabulary; texture references are glTF texture indices = image indices here);
    nodes: glTF node dicts; roots: root node indices; images: list of HxWx4 uint8 arrays (written as PNG);
    flavour: "embedded" (base64 data URIs), "external" (.bin + .png files next to the .gltf), "glb"."""
    blob = bytearray(); views = []; accessors = []

    def add_view(data: bytes, stride=None):
        while len(blob) % 4:
            blob.append(0)
        v = {"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)}
        if stride:
            v["byteStride"] = stride
        blob.extend(data); views.append(v)
        return len(views) - 1

    def add_accessor(arr, ctype, kind):
        arr = np.ascontiguousarray(arr)
        accessors.append({"bufferView": add_view(arr.tobytes()), "componentType": ctype, "count": int(arr.shape[0]), "type": kind})
        return len(accessors) - 1

    gl_meshes = []
    for prims in meshes:
        gp = []
        for p in prims:
            at = {}
            pos = np.asarray(p["positions"], F)
            if p.get("interleave_pos_normal") and p.get("normals") is not None:         # one strided view holding both streams
                inter = np.concatenate([pos, np.asarray(p["normals"], F)], axis=1)
                view = add_view(inter.tobytes(), stride=24)
                accessors.append({"bufferView": view, "componentType": 5126, "count": len(pos), "type": "VEC3"}); at["POSITION"] = len(accessors) - 1
                accessors.append({"bufferView": view, "byteOffset": 12, "componentType": 5126, "count": len(pos), "type": "VEC3"}); at["NORMAL"] = len(accessors) - 1
            else:
                at["POSITION"] = add_accessor(pos, 5126, "VEC3")
                if p.get("normals") is not None:
                    at["NORMAL"] = add_accessor(np.asarray(p["normals"], F), 5126, "VEC3")
            if p.get("uvs") is not None:
                at["TEXCOORD_0"] = add_accessor(np.asarray(p["uvs"], F), 5126, "VEC2")
            if p.get("tangents") is not None:
                at["TANGENT"] = add_accessor(np.asarray(p["tangents"], F), 5126, "VEC4")
            it = np.dtype(p.get("index_type", np.uint16))
            prim = {"attributes": at, "indices": add_accessor(np.asarray(p["indices"]).reshape(-1).astype(it), {1: 5121, 2: 5123, 4: 5125}[it.itemsize], "SCALAR")}
            if p.get("material", -1) is not None and p.get("material", -1) >= 0:
                prim["material"] = int(p["material"])
            gp.append(prim)
        gl_meshes.append({"primitives": gp})

    base = os.path.splitext(path)[0]
    gl_images = []
    for i, im in enumerate(images):
        data = png_bytes(np.asarray(im, np.uint8))
        if flavour == "external":
            name = f"{os.path.basename(base)}_img{i}.png"
            open(os.path.join(os.path.dirname(path), name), "wb").write(data)
            gl_images.append({"uri": name})
        elif flavour == "glb":
            gl_images.append({"bufferView": add_view(data), "mimeType": "image/png"})
        else:
            gl_images.append({"uri": "data:image/png;base64," + base64.b64encode(data).decode()})

    doc = {"asset": {"version": "2.0", "generator": "lumenrenderer_b200 tests"}, "scene": 0, "scenes": [{"name": "Scene", "nodes": list(roots)}],
           "nodes": nodes, "materials": materials, "meshes": gl_meshes, "accessors": accessors, "bufferViews": views}
    if gl_images:
        doc["images"] = gl_images
        doc["textures"] = [{"source": i} for i in range(len(gl_images))]
    if flavour == "glb":
        doc["buffers"] = [{"byteLength": len(blob)}]
        js = json.dumps(doc).encode(); js += b" " * (-len(js) % 4)
        bin_ = bytes(blob) + b"\0" * (-len(blob) % 4)
        total = 12 + 8 + len(js) + 8 + len(bin_)
        open(path, "wb").write(b"glTF" + struct.pack("<II", 2, total) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(bin_), 0x004E4942) + bin_)
        return
    if flavour == "external":
        name = os.path.basename(base) + ".bin"
        open(os.path.join(os.path.dirname(path), name), "wb").write(bytes(blob))
        doc["buffers"] = [{"byteLength": len(blob), "uri": name}]
    else:
        doc["buffers"] = [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(bytes(blob)).decode()}]
    json.dump(doc, open(path, "w"))


# ----------------------------------------------------------------------------------------------------------------- restatement
def _read_document(path):
    data = open(path, "rb").read()
    glb_bin = None
    if data[:4] == b"glTF":
        pos = 12; js = None
        while pos + 8 <= len(data):
            n, kind = struct.unpack("<II", data[pos:pos + 8])
            body = data[pos + 8:pos + 8 + n]
            if kind == 0x4E4F534A and js is None:
                js = body
            elif kind == 0x004E4942 and glb_bin is None:
                glb_bin = body
            pos += 8 + n
        doc = json.loads(js.decode())
    else:
        doc = json.loads(data.decode())
    return doc, glb_bin


def _uri_bytes(uri, base_dir):
    if uri.startswith("data:"):
        return base64.b64decode(uri.split(",", 1)[1])
    return open(os.path.join(base_dir, uri), "rb").read()


_COMP = {5120: ("i1", 1), 5121: ("u1", 1), 5122: ("<i2", 2), 5123: ("<u2", 2), 5125: ("<u4", 4), 5126: ("<f4", 4)}
_COUNT = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT2": 4, "MAT3": 9, "MAT4": 16}


def _accessor(doc, buffers, index):
    acc = doc["accessors"][index]; view = doc["bufferViews"][acc["bufferView"]]
    dt, size = _COMP[acc["componentType"]]; n = _COUNT[acc["type"]]; elem = size * n
    stride = max(elem, view.get("byteStride", 0)); base = view.get("byteOffset", 0) + acc.get("byteOffset", 0)
    buf = buffers[view["buffer"]]
    rows = [np.frombuffer(buf, dt, n, base + i * stride) for i in range(acc["count"])]
    return np.stack(rows) if rows else np.zeros((0, n), dt)


def _norm3(v):
    d = F(F(v[0] * v[0]) + F(v[1] * v[1])) + F(v[2] * v[2])
    inv = F(1.0) / np.sqrt(F(d))
    return np.array([v[0] * inv, v[1] * inv, v[2] * inv], F)


def _dot3(a, b):
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def generate_tangents(pos, nrm, uv, idx):
    """GenerateTangentBinary (:734-900) in float32, triangle by triangle; the last triangle referencing a vertex wins."""
    tan = np.zeros((len(pos), 4), F)
    default = np.array([[1, 1], [0, 1], [1, 0]], F)
    eps = F(np.finfo(np.float32).eps)
    with np.errstate(all="ignore"):
        for t in range(0, len(idx) - 2, 3):
            ix = idx[t:t + 3]
            tuv = uv[ix].astype(F) if uv is not None else default.copy()
            length = lambda a, b: np.sqrt(F(F((a[0] - b[0]) * (a[0] - b[0])) + F((a[1] - b[1]) * (a[1] - b[1]))))
            if length(tuv[0], tuv[1]) < eps or length(tuv[0], tuv[2]) < eps or length(tuv[2], tuv[1]) < eps:
                tuv = default.copy()
            dp1, dp2 = pos[ix[1]] - pos[ix[0]], pos[ix[2]] - pos[ix[0]]
            du1, du2 = tuv[1] - tuv[0], tuv[2] - tuv[0]
            cross = F(du1[0] * du2[1]) - F(du1[1] * du2[0])
            if cross == 0:
                du1, du2 = default[1] - default[0], default[2] - default[0]
            den = F(du1[0] * du2[1]) - F(du2[0] * du1[1])
            tg = (F(du2[1]) * dp1 - F(du1[1]) * dp2) / den
            for k in range(3):
                ng = _norm3(nrm[ix[k]])
                d = _dot3(ng, tg)
                tan[ix[k], :3] = _norm3(tg - ng * d); tan[ix[k], 3] = 1.0
    return tan


def _flat_normals(pos, idx):
    nrm = np.zeros_like(pos)
    for t in range(0, len(idx) - 2, 3):
        a, b = pos[idx[t + 1]] - pos[idx[t]], pos[idx[t + 2]] - pos[idx[t]]
        n = np.array([F(a[1] * b[2]) - F(b[1] * a[2]), F(a[2] * b[0]) - F(b[2] * a[0]), F(a[0] * b[1]) - F(b[0] * a[1])], F)
        for k in range(3):
            nrm[idx[t + k]] = nrm[idx[t + k]] + n
    for v in range(len(nrm)):
        l = np.sqrt(_dot3(nrm[v], nrm[v]))
        nrm[v] = nrm[v] * (F(1.0) / l) if l > 0 else np.array([0, 0, 1], F)
    return nrm


def _mat_mul(a, b):
    """glm operator*(mat4, mat4) on column-major arrays a[col][row]: ((A0*b0 + A1*b1) + A2*b2) + A3*b3."""
    r = np.zeros((4, 4), F)
    for j in range(4):
        r[j] = F(F(F(a[0] * b[j][0]) + F(a[1] * b[j][1])) + F(a[2] * b[j][2])) + F(a[3] * b[j][3])
    return r


def _node_local(n):
    m = np.array(n.get("matrix", np.eye(4).reshape(-1)), F).reshape(4, 4)            # [col][row]
    if not np.array_equal(m, np.eye(4, dtype=F)):
        return m
    t = np.array(n.get("translation", (0, 0, 0)), F); s = np.array(n.get("scale", (1, 1, 1)), F)
    x, y, z, w = (F(v) for v in n.get("rotation", (0, 0, 0, 1)))
    qxx, qyy, qzz, qxz, qxy, qyz, qwx, qwy, qwz = x * x, y * y, z * z, x * z, x * y, y * z, w * x, w * y, w * z
    one, two = F(1), F(2)
    rot = np.zeros((4, 4), F)                                 # glm::mat4_cast
    rot[0, :3] = np.array([one - two * (qyy + qzz), two * (qxy + qwz), two * (qxz - qwy)], F)
    rot[1, :3] = np.array([two * (qxy - qwz), one - two * (qxx + qzz), two * (qyz + qwx)], F)
    rot[2, :3] = np.array([two * (qxz + qwy), two * (qyz - qwx), one - two * (qxx + qyy)], F)
    rot[3, 3] = one
    # Transform::UpdateLocalMatrix (Transform.cpp:264-280) in glm's operation order (it decides the sign of zeros):
    # glm::translate(I, t): col3 = ((I0*tx + I1*ty) + I2*tz) + I3; then operator*(mat4, mat4); then glm::scale: column j (w included) * s[j]
    eye = np.eye(4, dtype=F)
    tr = eye.copy()
    tr[3] = F(F(F(eye[0] * t[0]) + F(eye[1] * t[1])) + F(eye[2] * t[2])) + eye[3]
    r = _mat_mul(tr, rot)
    for j in range(3):
        r[j] = r[j] * s[j]
    return r


def load_reference_semantics(path):
    """-> dict(images, materials, meshes, instances) with the same field names GltfDocument exposes."""
    doc, glb_bin = _read_document(path)
    base_dir = os.path.dirname(path)
    buffers = [(_uri_bytes(b["uri"], base_dir) if "uri" in b else glb_bin) for b in doc.get("buffers", [])]

    def tex_image(info):
        if not isinstance(info, dict) or "index" not in info:
            return -1
        return doc["textures"][info["index"]].get("source", -1)

    n_images = len(doc.get("images", []))
    types = [0] * n_images                                    # TextureType (LumenPTModelConverter.h:66-77): the LAST role assigned wins (:364-521)
    materials = []
    for m in doc.get("materials", []):
        pbr = m.get("pbrMetallicRoughness", {}); ext = m.get("extensions", {})
        d = {"diffuse_color": tuple(F(v) for v in pbr.get("baseColorFactor", (1, 1, 1, 1))), "emission": tuple(F(v) for v in m.get("emissiveFactor", (0, 0, 0))),
             "diffuse_texture": tex_image(pbr.get("baseColorTexture")), "normal_texture": tex_image(m.get("normalTexture")),
             "metallic_roughness_texture": tex_image(pbr.get("metallicRoughnessTexture")), "emissive_texture": tex_image(m.get("emissiveTexture")),
             "metallic_factor": F(pbr.get("metallicFactor", 1.0)), "roughness_factor": max(F(0.01), F(pbr.get("roughnessFactor", 1.0))),
             "luminance": F(1), "subsurface_factor": F(0), "anisotropic": F(0), "tint_factor": (F(0), F(0), F(0)), "transmittance": (F(0), F(0), F(0))}
        tr = ext.get("KHR_materials_transmission")
        d["transmission_factor"] = F(tr.get("transmissionFactor", 0.0)) if tr is not None else F(0)
        d["transmission_texture"] = tex_image(tr.get("transmissionTexture")) if tr is not None else -1
        sh = ext.get("KHR_materials_sheen")
        d["sheen_factor"] = F(sh.get("sheenRoughnessFactor", 0.0)) if sh is not None else F(0)
        d["sheen_tint_factor"] = F(1 if sh is not None else 0)
        ior = ext.get("KHR_materials_ior")
        d["index_of_refraction"] = F(ior.get("ior", 1.0)) if ior is not None else F(1)
        cc = ext.get("KHR_materials_clearcoat")
        d["clear_coat_factor"] = F(cc.get("clearcoatFactor", 0.0)) if cc is not None else F(0)
        d["clear_coat_roughness_factor"] = F(cc.get("clearcoatRoughnessFactor", 0.0)) if cc is not None else F(0)
        d["clear_coat_texture"] = tex_image(cc.get("clearcoatTexture")) if cc is not None else -1
        d["clear_coat_roughness_texture"] = tex_image(cc.get("clearcoatRoughnessTexture")) if cc is not None else -1
        sp = ext.get("KHR_materials_specular")
        d["specular_factor"] = F(sp.get("specularFactor", 0.0)) if sp is not None else F(0)
        d["specular_tint_factor"] = F(1 if sp is not None else 0)
        d["tint_texture"] = tex_image(sp.get("specularColorTexture")) if sp is not None else -1
        for key, role in (("diffuse_texture", 1), ("normal_texture", 2), ("metallic_roughness_texture", 4), ("emissive_texture", 3), ("transmission_texture", 5),
                          ("clear_coat_roughness_texture", 7), ("clear_coat_texture", 6), ("tint_texture", 8)):          # in the converter's statement order
            if d[key] >= 0:
                types[d[key]] = role
        materials.append(d)

    meshes = []
    for mesh in doc.get("meshes", []):
        prims = []
        for fp in mesh["primitives"]:
            at = fp["attributes"]
            pos = _accessor(doc, buffers, at["POSITION"]).astype(F)
            uv = _accessor(doc, buffers, at["TEXCOORD_0"]).astype(F) if "TEXCOORD_0" in at else None
            nrm = _accessor(doc, buffers, at["NORMAL"]).astype(F) if "NORMAL" in at else None
            tan = _accessor(doc, buffers, at["TANGENT"]).astype(F) if "TANGENT" in at else None
            idx = _accessor(doc, buffers, fp["indices"]).reshape(-1).astype(np.uint32) if "indices" in fp else np.arange(len(pos), dtype=np.uint32)
            idx = idx[:len(idx) // 3 * 3]
            if nrm is None:
                nrm = _flat_normals(pos, idx)
            if tan is None:
                tan = generate_tangents(pos, nrm, uv, idx)
            prims.append({"positions": pos, "uvs": uv if uv is not None else np.zeros((len(pos), 2), F), "normals": nrm, "tangents": tan,
                          "indices": idx, "material": fp.get("material", -1)})
        meshes.append(prims)

    instances = []

    def visit(index, parent_world):
        n = doc["nodes"][index]
        local = _node_local(n)
        with_parent = _mat_mul(parent_world, local) if parent_world is not None else local
        if "mesh" in n:
            # a ROOT mesh node's instance keeps an identity world matrix: Transform::operator= (Transform.cpp:58-75) copies the local matrix
            # and a clean dirty flag but not the world matrix; only AddChild (parented nodes) raises the flag again
            shown = with_parent if parent_world is not None else np.eye(4, dtype=F)
            instances.append({"mesh": n["mesh"], "transform": shown.T.copy()})      # row-major
            own = local                                                                       # :296-306 quirk
        else:
            own = with_parent
        for c in n.get("children", []):
            visit(c, own)
    for scene in doc.get("scenes", []):
        for r in scene.get("nodes", []):
            visit(r, None)
    srgb = [t in (1, 3) for t in types]                       # LoadFile :131
    metal_rough = [t == 4 for t in types]                     # LoadFile :122
    return {"srgb": srgb, "metal_rough": metal_rough, "texture_types": types, "materials": materials, "meshes": meshes, "instances": instances, "doc": doc, "buffers": buffers}
