"""Mutation fuzzing of the host-only asset readers (csrc/lb_nanovdb.cpp, csrc/lb_gltf.cpp + lb_png.h / lb_json.h) under AddressSanitizer and
UndefinedBehaviorSanitizer: truncations, random bytes, random words, JSON token and digit substitutions of valid files. A malformed file
must be an error code — never a crash, an out-of-bounds access or an unbounded allocation (round 1 found and fixed two: an inflate stream
that kept decoding zero padding, and an accessor bound check that overflowed)."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "tests"))
CSRC = os.path.join(ROOT, "lumenrenderer_b200", "csrc")


def _build(tmp_path, name, source):
    exe = str(tmp_path / name)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-w", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", f"-I{ROOT}/include",
           os.path.join(ROOT, "tests", "fuzz", name + ".cpp"), os.path.join(CSRC, source), "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 and ("asan" in res.stderr.lower() or "sanitize" in res.stderr.lower()):
        pytest.skip("this toolchain has no sanitizer runtime")
    assert res.returncode == 0, res.stderr[-2000:]
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_nanovdb_reader_survives_mutated_files(tmp_path):
    exe = _build(tmp_path, "fuzz_nanovdb", "lb_nanovdb.cpp")
    files = [os.path.join(GOLDEN, "nanovdb", f) for f in ("fog5_raw.vndb", "fog12_zip.vndb", "ls10_zip.vndb")]
    res = subprocess.run([exe, "700", "20261017", *files], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "fuzzed 2100 inputs" in res.stdout, (res.stdout[-500:], res.stderr[-3000:])


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_gltf_loader_survives_mutated_files(tmp_path):
    import test_gltf
    exe = _build(tmp_path, "fuzz_gltf", "lb_gltf.cpp")
    files = [test_gltf.build_test_document(str(tmp_path / "scene.gltf"), "embedded"), test_gltf.build_test_document(str(tmp_path / "scene.glb"), "glb")]
    # the `.ollad` cache reader: the reference's own cache file of its Cornell box, and the cache of the synthetic document (images, names, hierarchy)
    from lumenrenderer_b200.gltf import GltfDocument
    with GltfDocument(files[1]) as doc:
        doc.save_ollad(str(tmp_path / "scene.ollad"))
    files += [os.path.join(GOLDEN, "cornell_reference.ollad"), str(tmp_path / "scene.ollad")]
    # a GLB whose BIN chunk holds the synthetic PNG corner cases (sub-byte depths, palettes, colour keys, Adam7): mutations reach every decoder path
    import json, struct
    import numpy as np
    cases = np.load(os.path.join(GOLDEN, "png_cases.npz"))
    blob = b""; views = []
    for name in sorted({k.split("/")[0] for k in cases.files}):
        data = cases[name + "/file"].tobytes()
        views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)}); blob += data + b"\0" * (-len(data) % 4)
    doc = json.dumps({"asset": {"version": "2.0"}, "buffers": [{"byteLength": len(blob)}], "bufferViews": views,
                      "images": [{"bufferView": i, "mimeType": "image/png"} for i in range(len(views))]}).encode()
    doc += b" " * (-len(doc) % 4)
    with open(tmp_path / "png_cases.glb", "wb") as f:
        f.write(struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(doc) + 8 + len(blob)) + struct.pack("<II", len(doc), 0x4E4F534A) + doc + struct.pack("<II", len(blob), 0x004E4942) + blob)
    with GltfDocument(str(tmp_path / "png_cases.glb")) as doc_:
        assert doc_.info["images"] == len(views) and doc_.info["undecoded_images"] == 0
    files.append(str(tmp_path / "png_cases.glb"))
    # the same for the JPEG decoder (csrc/lb_jpeg.h): baseline / progressive, every sampling ratio, restart intervals, CMYK, 16-bit tables
    cases = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    blob = b""; views = []
    for name in sorted({k.split("/")[0] for k in cases.files}):
        data = cases[name + "/file"].tobytes()
        views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)}); blob += data + b"\0" * (-len(data) % 4)
    doc = json.dumps({"asset": {"version": "2.0"}, "buffers": [{"byteLength": len(blob)}], "bufferViews": views,
                      "images": [{"bufferView": i, "mimeType": "image/jpeg"} for i in range(len(views))]}).encode()
    doc += b" " * (-len(doc) % 4)
    with open(tmp_path / "jpeg_cases.glb", "wb") as f:
        f.write(struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(doc) + 8 + len(blob)) + struct.pack("<II", len(doc), 0x4E4F534A) + doc + struct.pack("<II", len(blob), 0x004E4942) + blob)
    with GltfDocument(str(tmp_path / "jpeg_cases.glb")) as doc_:
        assert doc_.info["images"] == len(views) and doc_.info["undecoded_images"] == 0
    files.append(str(tmp_path / "jpeg_cases.glb"))
    env = dict(os.environ, FUZZ_TMP=str(tmp_path))
    res = subprocess.run([exe, "700", "4711", *files], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "fuzzed 4200 inputs" in res.stdout, (res.stdout[-500:], res.stderr[-3000:])
