"""Golden data from the reference's OWN model converter (LumenPTModelConverter::GenerateContent, compiled for the host in place by
oracle/Makefile -> oracle/_ref/ref_gltf; needs /root/reference, so this runs in the build container only) for the two glTF assets the
reference ships: Sandbox/assets/models/CornellBox/scene.gltf (stored completely, 7 KB) and Sponza/Sponza.gltf (262 267 triangles: the
material table and node table completely, every vertex stream and index buffer as a SHA-256). Also the `.ollad` cache file the reference writes for each (Cornell: the file itself, tests/golden/cornell_reference.ollad, 7 KB;
Sponza: its SHA-256 and size), and the SHA-256 of the cache file of EVERY glTF / GLB asset under Sandbox/assets/models. Writes tests/golden/gltf_reference_converter.npz."""
import hashlib
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ASSETS = "/root/reference/Lumen_Engine/Sandbox/assets/models"
FILES = {"cornell": ASSETS + "/CornellBox/scene.gltf", "sponza": ASSETS + "/Sponza/Sponza.gltf"}

# LumenPTModelConverter::HeaderMaterial (LumenPTModelConverter.h:85-137), 136 bytes
MATERIAL = np.dtype([("color", "<f4", 4), ("emission", "<f4", 3), ("diffuse_texture", "<i4"), ("normal_texture", "<i4"), ("metallic_roughness_texture", "<i4"),
                     ("emissive_texture", "<i4"), ("transmission_texture", "<i4"), ("clear_coat_texture", "<i4"), ("clear_coat_roughness_texture", "<i4"), ("tint_texture", "<i4"),
                     ("transmission_factor", "<f4"), ("clear_coat_factor", "<f4"), ("clear_coat_roughness_factor", "<f4"), ("index_of_refraction", "<f4"),
                     ("specular_factor", "<f4"), ("specular_tint_factor", "<f4"), ("subsurface_factor", "<f4"), ("luminance", "<f4"), ("anisotropic", "<f4"),
                     ("sheen_factor", "<f4"), ("sheen_tint_factor", "<f4"), ("metallic_factor", "<f4"), ("roughness_factor", "<f4"), ("tint_factor", "<f4", 3),
                     ("transmittance", "<f4", 3)])


def read_dump(path):
    """Parses the record oracle/_ref/ref_gltf writes (layout in oracle/ref_shim/ref_gltf.cpp)."""
    b = open(path, "rb").read(); o = 0

    def u32():
        nonlocal o
        v, = struct.unpack_from("<I", b, o); o += 4
        return v
    n = u32(); size = u32()
    assert size == MATERIAL.itemsize, (size, MATERIAL.itemsize)
    mats = np.frombuffer(b, MATERIAL, n, o).copy(); o += n * size
    nt = u32(); tex = np.frombuffer(b, "<u8", nt * 3, o).reshape(nt, 3).copy(); o += nt * 24
    prims = []
    for mesh in range(u32()):
        for _ in range(u32()):
            vbytes, ibytes, isize, mat = struct.unpack_from("<4Q", b, o); o += 32
            v = np.frombuffer(b, "<f4", vbytes // 4, o).reshape(-1, 16).copy(); o += vbytes
            idx = np.frombuffer(b, "<u2" if isize == 2 else "<u4", ibytes // isize, o).astype(np.uint32); o += ibytes
            prims.append({"mesh": mesh, "vertices": v, "indices": idx, "material": int(mat), "index_size": int(isize)})
    nodes = []                                                # depth first, the order LoadNode (:953-992) emplaces them

    def node():
        nonlocal o
        t = np.frombuffer(b, "<f4", 16, o).copy(); o += 64
        mesh, children = struct.unpack_from("<iI", b, o); o += 8
        nodes.append((t, mesh, children))
        for _ in range(children):
            node()
    scenes = []
    for _ in range(u32()):
        roots = u32(); scenes.append(roots)
        for _ in range(roots):
            node()
    assert o == len(b)
    return mats, tex[:, 2].astype(np.int32), prims, nodes, scenes


def stream_hashes(positions, uvs, normals, tangents, indices):
    """SHA-256 of each stream; tangents only at the vertices the index buffer uses (the reference sizes its tangent buffer by the index
    count and leaves the rest of the vertex range unspecified)."""
    used = np.unique(indices)
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    return [h(positions.astype("<f4")), h(uvs.astype("<f4")), h(normals.astype("<f4")), h(tangents.astype("<f4")[used]), h(indices.astype("<u4"))]


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_gltf")
    gold = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, path in FILES.items():
            out = os.path.join(tmp, name + ".bin")
            ollad = os.path.join(tmp, name + ".ollad")        # the reference's cache file itself (GenerateHeader + OutputToFile)
            subprocess.check_call([tool, path, out, ollad], stdout=subprocess.DEVNULL)
            data = open(ollad, "rb").read()
            gold[f"{name}/ollad_sha256"] = np.array([hashlib.sha256(data).hexdigest(), str(len(data))])
            if name == "cornell":
                open(os.path.join(HERE, "cornell_reference.ollad"), "wb").write(data)
            mats, tex_types, prims, nodes, scenes = read_dump(out)
            gold[f"{name}/materials"] = mats
            gold[f"{name}/texture_types"] = tex_types
            gold[f"{name}/prim_mesh"] = np.array([p["mesh"] for p in prims], np.int32)
            gold[f"{name}/prim_material"] = np.array([p["material"] for p in prims], np.int32)
            gold[f"{name}/prim_counts"] = np.array([[len(p["vertices"]), len(p["indices"]), p["index_size"]] for p in prims], np.int64)
            gold[f"{name}/prim_hashes"] = np.array([stream_hashes(p["vertices"][:, 0:3], p["vertices"][:, 4:6], p["vertices"][:, 6:9], p["vertices"][:, 12:16], p["indices"]) for p in prims])
            gold[f"{name}/node_local"] = np.array([n[0] for n in nodes], np.float32).reshape(-1, 16)
            gold[f"{name}/node_mesh_children"] = np.array([[n[1], n[2]] for n in nodes], np.int32).reshape(-1, 2)
            gold[f"{name}/scene_roots"] = np.array(scenes, np.int32)
            if name == "cornell":
                for k, p in enumerate(prims):
                    gold[f"cornell/vertices{k}"] = p["vertices"]; gold[f"cornell/indices{k}"] = p["indices"]
        # every glTF / GLB asset the reference ships (29 files, 186 MB of caches): SHA-256 and size of the cache file its converter writes;
        # an empty hash where the reference's converter does not survive the file (Draco-compressed variants, a truncated sample)
        table = []
        for dirpath, _, files in sorted(os.walk(ASSETS)):
            for f in sorted(files):
                if not f.endswith((".gltf", ".glb")):
                    continue
                path = os.path.join(dirpath, f); ollad = os.path.join(tmp, "any.ollad")
                try:
                    ok = subprocess.run([tool, path, os.path.join(tmp, "any.bin"), ollad], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600).returncode == 0
                except subprocess.TimeoutExpired:
                    ok = False
                data = open(ollad, "rb").read() if ok else b""
                table.append([os.path.relpath(path, ASSETS), hashlib.sha256(data).hexdigest() if ok else "", str(len(data))])
                print(table[-1])
        gold["assets/ollad_sha256"] = np.array(table)
    np.savez_compressed(os.path.join(HERE, "gltf_reference_converter.npz"), **gold)
    print("wrote", os.path.join(HERE, "gltf_reference_converter.npz"))


if __name__ == "__main__":
    main()
