"""Drop-in boundary, reference side: include/lumen_b200_adapter.hpp (B200::Renderer : LumenRenderer) compiled against the reference's
own interface headers and driven by a program written purely against that interface (tests/adapter/adapter_driver.cpp: CreateTexture /
CreateMaterial / ILumenMaterial setters / CreatePrimitive / CreateMesh / m_Scene->AddMesh() / MeshInstance / Transform / Camera /
TraceFrame / GetOutputTexturePixels / GetLastFrameStats). The same scene is then rendered through the plain C ABI from Python with the
world matrices the reference's Transform / Camera classes produced; both must agree bit for bit.

CPU (`not gpu`): the adapter linked against the oracle library (same ABI, lo_ prefix) — needs /root/reference, skipped without it.
GPU (`gpu`): the executable prebuilt here against liblumen_b200.so travels to the GPU box in tests/adapter/_build/.
"""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "adapter"))
import build as adapter_build  # noqa: E402

from lumenrenderer_b200 import api, scenes, nanovdb  # noqa: E402
import nanovdb_tools  # noqa: E402

MATERIAL_FLOATS = ("transmission_factor", "clear_coat_factor", "clear_coat_roughness_factor", "index_of_refraction", "specular_factor", "specular_tint_factor",
                   "subsurface_factor", "luminance", "anisotropic", "sheen_factor", "sheen_tint_factor", "metallic_factor", "roughness_factor")
MATERIAL_TEXTURES = ("diffuse_texture", "normal_texture", "metallic_roughness_texture", "emissive_texture", "transmission_texture", "clear_coat_texture",
                     "clear_coat_roughness_texture", "tint_texture")


def dump_scene(path, scene, width, height, depth, restir, frames, interleaved):
    f32 = lambda a, n: np.ascontiguousarray(a, np.float32).reshape(-1)[:n].tobytes() if a is not None else np.zeros(n, np.float32).tobytes()
    with open(path, "wb") as f:
        f.write(struct.pack("<7I", 0x4353424C, width, height, depth, int(restir), frames, int(interleaved)))
        f.write(f32(scene.camera["position"], 3)); f.write(f32(scene.camera.get("rotation", (1, 0, 0, 0)), 4))
        f.write(struct.pack("<I", len(scene.textures)))
        for t in scene.textures:
            px = np.ascontiguousarray(t["pixels"], np.uint8)
            f.write(struct.pack("<3I", px.shape[1], px.shape[0], int(t.get("srgb", False)))); f.write(px.tobytes())
        f.write(struct.pack("<I", len(scene.materials)))
        for m in scene.materials:
            md = api.MaterialData(**{k: v for k, v in m.items() if not k.endswith("_texture")})
            f.write(f32(md.diffuse_color, 4)); f.write(f32(md.emission, 3))
            f.write(struct.pack(f"<{len(MATERIAL_FLOATS)}f", *[getattr(md, k) for k in MATERIAL_FLOATS]))
            f.write(f32(md.tint_factor, 3)); f.write(f32(md.transmittance, 3))
            f.write(struct.pack("<8i", *[(m.get(k) if m.get(k) is not None else -1) for k in MATERIAL_TEXTURES]))
        f.write(struct.pack("<I", len(scene.meshes)))
        for mesh in scene.meshes:
            f.write(struct.pack("<I", len(mesh)))
            for p in mesh:
                pos = np.ascontiguousarray(p["positions"], np.float32).reshape(-1, 3); n = len(pos)
                idx = np.ascontiguousarray(p["indices"]).reshape(-1)
                small = n < 65536 and (len(idx) // 3) % 2 == 0          # exercise both index widths
                f.write(struct.pack("<4I", n, len(idx), p["material"], 2 if small else 4))
                f.write(pos.tobytes()); f.write(f32(p.get("uvs"), 2 * n)); f.write(f32(p.get("normals"), 3 * n)); f.write(f32(p.get("tangents"), 4 * n))
                f.write(idx.astype(np.uint16 if small else np.uint32).tobytes())
        f.write(struct.pack("<I", len(scene.instances)))
        for inst in scene.instances:
            m = np.eye(4, dtype=np.float32) if inst.get("transform") is None else np.asarray(inst["transform"], np.float32).reshape(4, 4)
            f.write(struct.pack("<I", inst["mesh"])); f.write(m.tobytes())
            f.write(struct.pack("<i3ffi", inst.get("emission_mode", api.EMISSION_ENABLED), *inst.get("override_radiance", (0, 0, 0)), inst.get("emission_scale", 1.0),
                                inst.get("override_material", -1)))
        files = [v for v in scene.volumes if v.get("file")]
        if files:
            f.write(struct.pack("<I", len(files)))
            for v in files:
                path = os.fsencode(v["file"])
                m = np.eye(4, dtype=np.float32) if v.get("transform") is None else np.asarray(v["transform"], np.float32).reshape(4, 4)
                f.write(struct.pack("<I", len(path))); f.write(path); f.write(m.tobytes()); f.write(struct.pack("<f", v.get("instance_density", 0.001)))


def render_through_c_abi(bindings, scene, worlds, width, height, depth, restir, frames):
    """The same scene through the plain C ABI, with the matrices the reference classes computed."""
    st = api.Settings(width=width, height=height, depth=depth, restir=restir)
    with api.Renderer(bindings, st) as r:
        s2 = scenes.SceneDescription(textures=scene.textures, materials=scene.materials, meshes=scene.meshes, volumes=[], camera=None,
                                     instances=[dict(inst, transform=worlds[k]) for k, inst in enumerate(scene.instances)])
        # missing vertex streams are zero-filled by the driver (as the reference's interleaving does); do the same here
        for mesh in s2.meshes:
            for p in mesh:
                n = len(np.asarray(p["positions"]).reshape(-1, 3))
                for key, w in (("uvs", 2), ("normals", 3), ("tangents", 4)):
                    if p.get(key) is None:
                        p[key] = np.zeros((n, w), np.float32)
        r.load_scene(s2)
        for k, v in enumerate(v for v in scene.volumes if v.get("file")):
            world = worlds[len(scene.instances) + 1 + k]
            if bindings.prefix == "lb_":
                h = nanovdb.create_volume_from_file(r, v["file"])
            else:                                   # the oracle has no file reader: the numpy restatement's density box stands in
                g = nanovdb_tools.read_grid(open(v["file"], "rb").read())
                h = r.create_volume(g.density(), *g.volume_box())
            r.add_volume_instance(h, world, v.get("instance_density", 0.001))
        r.set_camera_matrix(worlds[len(scene.instances)])
        r.render_frames(frames)
        return r.read_hdr(), r.read_ldr()


def run_case(exe, bindings, tmp_path, scene, width, height, depth, restir, frames, interleaved):
    dump = str(tmp_path / "scene.bin"); out = str(tmp_path / "out")
    dump_scene(dump, scene, width, height, depth, restir, frames, interleaved)
    res = subprocess.run([exe, dump, out], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert f"frames {frames} resolution {width}x{height} instances {len(scene.instances)} frame-id {frames}" in res.stdout, res.stdout
    worlds = np.fromfile(out + ".worlds", np.float32).reshape(-1, 16)
    assert len(worlds) == len(scene.instances) + 1 + sum(1 for v in scene.volumes if v.get("file"))
    hdr = np.fromfile(out + ".hdr", np.float32).reshape(height, width, 4)
    ldr = np.fromfile(out + ".ldr", np.uint8).reshape(height, width, 4)
    ref_hdr, ref_ldr = render_through_c_abi(bindings, scene, worlds, width, height, depth, restir, frames)
    assert np.isfinite(hdr).all() and hdr[..., :3].sum() > 0
    assert np.array_equal(hdr, ref_hdr), "adapter and C-ABI renders differ"
    assert np.array_equal(ldr, np.asarray(ref_ldr).reshape(ldr.shape))
    return worlds


def moved_cornell():
    """Cornell box whose two boxes carry non-trivial instance transforms (rotation + non-uniform scale + translation): goes through
    Transform::operator=(mat4) -> Decompose -> GetWorldTransformationMatrix on the reference side."""
    s = scenes.cornell_box()
    s.instances[4]["transform"] = scenes.translate(0.05, 0.0, 0.1, scale=0.9, angle_y_deg=12.0)
    s.instances[5]["transform"] = scenes.translate(-0.03, 0.0, -0.05, scale=1.05, angle_y_deg=-8.0)
    s.instances[3].update(emission_mode=api.EMISSION_OVERRIDE, override_radiance=(3.0, 2.5, 2.0), emission_scale=1.5)
    return s


def room_with_nanovdb_volume():
    """The fog room with its medium loaded from a NanoVDB file through LumenRenderer::CreateVolume(path) (level-set sphere fixture written
    by the reference's own NanoVDB, tests/golden/make_golden_nanovdb.py), scaled and moved in front of the two boxes."""
    s = scenes.fog_room(grid=8)
    s.volumes = [{"file": os.path.join(ROOT, "tests", "golden", "nanovdb", "ls10_zip.vndb"), "instance_density": 0.8,
                  "transform": scenes.translate(-8.4, 4.2, -1.0, scale=0.4)}]
    return s


@pytest.mark.skipif(not adapter_build.available(), reason="needs the reference tree (/root/reference) to compile against")
@pytest.mark.parametrize("case", ["cornell_nee_separate_streams", "moved_restir_interleaved", "gallery_textures", "volume_from_nanovdb_file"])
def test_adapter_over_reference_interface_cpu(oracle, tmp_path, case):
    exe = adapter_build.build()["oracle"]
    if case == "cornell_nee_separate_streams":
        worlds = run_case(exe, oracle, tmp_path, scenes.cornell_box(), 48, 40, 3, False, 1, False)
        assert np.array_equal(worlds[0].reshape(4, 4), np.eye(4, dtype=np.float32))
        cam = worlds[-1].reshape(4, 4)          # Camera.cpp:128-140 for quat (w0,x0,y1,z0): right -x, up +y, forward -z, position (0,1,2)
        assert np.allclose(cam, [[-1, 0, 0, 0], [0, 1, 0, 1], [0, 0, -1, 2], [0, 0, 0, 1]], atol=1e-6)
    elif case == "moved_restir_interleaved":
        worlds = run_case(exe, oracle, tmp_path, moved_cornell(), 40, 32, 3, True, 2, True)
        want = np.asarray(scenes.translate(0.05, 0.0, 0.1, scale=0.9, angle_y_deg=12.0), np.float32).reshape(4, 4)
        assert np.allclose(worlds[4].reshape(4, 4), want, atol=1e-5)       # decompose + recompose reproduces the matrix to rounding
    elif case == "volume_from_nanovdb_file":
        run_case(exe, oracle, tmp_path, room_with_nanovdb_volume(), 64, 40, 3, False, 2, False)
    else:
        run_case(exe, oracle, tmp_path, scenes.material_gallery(), 40, 24, 3, True, 1, True)


@pytest.mark.gpu
def test_adapter_over_reference_interface_gpu(gpu, tmp_path):
    exe = os.path.join(ROOT, "tests", "adapter", "_build", "adapter_driver_b200")
    if not os.path.exists(exe):
        pytest.skip("tests/adapter/_build/adapter_driver_b200 was not prebuilt (needs /root/reference at build time)")
    run_case(exe, gpu, tmp_path, moved_cornell(), 96, 64, 3, True, 2, True)
    run_case(exe, gpu, tmp_path, room_with_nanovdb_volume(), 96, 64, 3, False, 2, False)


def run_ollad_case(exe, bindings, tmp_path, cache, width, height, depth, restir, frames, exact_worlds=True):
    """The reference's own LumenPTModelConverter::LoadFile (tests/adapter/ollad_driver.cpp) in front of the adapter, against the same
    `.ollad` file opened by lb_gltf_open and uploaded by lb_gltf_upload through the plain C ABI."""
    from lumenrenderer_b200.gltf import GltfDocument
    out = str(tmp_path / "ollad_out")
    cam = scenes.cornell_box().camera
    args = [str(v) for v in (width, height, depth, int(restir), frames, *cam["position"], *cam["rotation"])]
    res = subprocess.run([exe, cache, out, *args], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    worlds = np.fromfile(out + ".worlds", np.float32).reshape(-1, 4, 4)
    hdr = np.fromfile(out + ".hdr", np.float32).reshape(height, width, 4)
    ldr = np.fromfile(out + ".ldr", np.uint8).reshape(height, width, 4)
    with GltfDocument(cache) as doc, api.Renderer(bindings, api.Settings(width=width, height=height, depth=depth, restir=restir)) as r:
        info = doc.info
        assert f"loaded {info['materials']} materials {info['meshes']} meshes {info['instances']} instances, resolution {width}x{height}" in res.stdout, res.stdout
        assert len(worlds) == info["instances"] + 1
        # the node table through the reference's Transform hierarchy (LoadNode :275-317, Transform::AddChild / GetWorldTransformationMatrix)
        # against lb_gltf_instance: the mesh-node quirk included
        for i in range(info["instances"]):
            t = doc.instance(i)["transform"]
            if exact_worlds:
                assert np.array_equal(t.view(np.uint32), worlds[i].view(np.uint32)), f"instance {i}"
            else:
                assert np.allclose(t, worlds[i], rtol=0, atol=2e-6 * max(1.0, np.abs(t).max())), f"instance {i}"
        if bindings.prefix == "lb_":
            doc.upload(r)
        else:                                    # lb_gltf_upload is product code; the oracle gets the same document call by call
            r.load_scene(doc.to_scene_description())
        for k in range(info["instances"]):
            r.set_instance_transform(k, worlds[k])
        r.set_camera_matrix(worlds[-1])
        r.render_frames(frames)
        ref_hdr, ref_ldr = r.read_hdr(), np.asarray(r.read_ldr()).reshape(ldr.shape)
    assert np.isfinite(hdr).all() and hdr[..., :3].sum() > 0
    return hdr, ref_hdr, ldr, ref_ldr


@pytest.mark.skipif(not adapter_build.available(), reason="needs the reference tree (/root/reference) to compile against")
@pytest.mark.parametrize("case", ["reference_cornell_cache", "textured_hierarchy"])
def test_reference_ollad_loader_over_the_adapter_cpu(oracle, tmp_path, case):
    exe = adapter_build.build()["ollad_oracle"]
    if case == "reference_cornell_cache":
        cache = os.path.join(ROOT, "tests", "golden", "cornell_reference.ollad")
        hdr, ref_hdr, ldr, ref_ldr = run_ollad_case(exe, oracle, tmp_path, cache, 48, 40, 3, False, 1)
    else:
        import test_gltf
        from lumenrenderer_b200.gltf import GltfDocument
        src = test_gltf.build_test_document(str(tmp_path / "scene.glb"), "glb", reference_safe=True)
        cache = str(tmp_path / "scene.ollad")
        with GltfDocument(src) as doc:
            doc.save_ollad(cache)
        # parented, rotated and scaled nodes: the reference's Transform decomposes every matrix it is given and recomposes it (1 ulp apart)
        hdr, ref_hdr, ldr, ref_ldr = run_ollad_case(exe, oracle, tmp_path, cache, 48, 36, 3, True, 2, exact_worlds=False)
    assert np.array_equal(hdr, ref_hdr), f"reference loader + adapter and lb_gltf differ: {np.abs(hdr - ref_hdr).max()}"
    assert np.array_equal(ldr, ref_ldr)


@pytest.mark.gpu
def test_reference_ollad_loader_over_the_adapter_gpu(gpu, tmp_path):
    """The reference's LoadFile (prebuilt here against liblumen_b200.so) in front of the GPU renderer vs lb_gltf_open + lb_gltf_upload."""
    exe = os.path.join(ROOT, "tests", "adapter", "_build", "ollad_driver_b200")
    if not os.path.exists(exe):
        pytest.skip("tests/adapter/_build/ollad_driver_b200 was not prebuilt (needs /root/reference at build time)")
    import test_gltf
    from lumenrenderer_b200.gltf import GltfDocument
    hdr, ref_hdr, ldr, ref_ldr = run_ollad_case(exe, gpu, tmp_path, os.path.join(ROOT, "tests", "golden", "cornell_reference.ollad"), 96, 64, 3, False, 1)
    assert np.array_equal(hdr, ref_hdr) and np.array_equal(ldr, ref_ldr)
    src = test_gltf.build_test_document(str(tmp_path / "scene.glb"), "glb", reference_safe=True)
    cache = str(tmp_path / "scene.ollad")
    with GltfDocument(src) as doc:
        doc.save_ollad(cache)
    hdr, ref_hdr, ldr, ref_ldr = run_ollad_case(exe, gpu, tmp_path, cache, 96, 64, 3, True, 2, exact_worlds=False)
    assert np.array_equal(hdr, ref_hdr) and np.array_equal(ldr, ref_ldr)
