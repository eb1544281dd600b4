"""csrc/lb_jpeg.h against the stb_image the reference decodes its textures with (Lumen/vendor/stb/stb_image.h v2.25 compiled in place ->
oracle/_ref/ref_stb; LumenPTModelConverter.cpp:105-131 stbi_load_from_memory(..., 4)). A JPEG file does not fix its pixels — inverse DCT,
chroma up-sampling and colour conversion are decoder choices — so the decoder is pinned pixel for pixel, like the PNG one."""
import hashlib
import os

import numpy as np
import pytest

from lumenrenderer_b200.gltf import GltfDocument
from conftest import GOLDEN
from test_gltf import REF_CORNELL


def _decode(tmp_path, name, data=None, link=None):
    img = os.path.join(tmp_path, name + ".jpg")
    if data is not None:
        open(img, "wb").write(data)
    else:
        os.symlink(link, img)
    path = os.path.join(tmp_path, name + ".gltf")
    open(path, "w").write('{"asset": {"version": "2.0"}, "images": [{"uri": "%s.jpg"}]}' % name)
    with GltfDocument(path) as doc:
        return doc.image(0)


def test_jpeg_decoder_corner_cases_against_stb_image(tmp_path):
    """22 small synthetic files (tests/golden/jpeg_cases.npz, written with Pillow / OpenCV by tests/golden/make_golden_jpeg.py) with the pixels
    the reference's stb_image decodes them to: baseline and progressive at 4:4:4 / 4:2:2 / 4:2:0 / 4:1:1 / 4:4:0, grey, sizes that leave
    partial MCUs, one-pixel / one-row / one-column images, restart intervals, optimised Huffman tables, 16-bit quantisation tables (quality 1),
    quality 100, Adobe CMYK. Needs no reference tree."""
    g = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) == 22
    for name in names:
        im = _decode(str(tmp_path), name, data=g[name + "/file"].tobytes())
        assert im["decoded"], name
        want = g[name + "/rgba"]
        assert im["pixels"].shape == want.shape, name
        assert np.array_equal(im["pixels"], want), f"{name}: {(im['pixels'] != want).any(axis=-1).sum()} of {want.shape[0] * want.shape[1]} pixels differ, max {np.abs(im['pixels'].astype(int) - want).max()}"


def test_jpeg_decoder_matches_stb_image_on_every_shipped_jpeg(tmp_path):
    """All 81 JPEG files under the reference's Sandbox/assets/models — Sponza's 65 textures (baseline and progressive, up to 2048 x 2048) among
    them — decode to the same RGBA8 pixels as stb_image (SHA-256 table: tests/golden/jpeg_reference.npz)."""
    root = os.path.dirname(os.path.dirname(REF_CORNELL))
    if not os.path.isdir(root):
        pytest.skip("the reference's Sandbox assets are not on this machine")
    table = np.load(os.path.join(GOLDEN, "jpeg_reference.npz"))["table"]
    assert len(table) == 81 and {r[3] for r in table} == {"0xc0", "0xc2"}
    seen = set()
    for k, (rel, w, h, sof, samp, sha) in enumerate(table):
        if sha in seen:
            continue
        seen.add(sha)
        im = _decode(str(tmp_path), f"img{k}", link=os.path.join(root, rel))
        assert im["decoded"], f"{rel} ({sof}, {samp}) was not decoded"
        assert im["pixels"].shape == (int(h), int(w), 4) and hashlib.sha256(im["pixels"].tobytes()).hexdigest() == sha, rel
    assert len(seen) >= 60


def test_malformed_jpeg_is_an_error_not_a_crash(tmp_path):
    g = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    data = g["prog_420/file"].tobytes()
    rng = np.random.default_rng(3)
    for trial in range(60):
        b = bytearray(data)
        if trial % 3 == 0:
            b = b[: int(rng.integers(4, len(b)))]                                   # truncation
        else:
            for _ in range(int(rng.integers(1, 8))):
                b[int(rng.integers(2, len(b)))] = int(rng.integers(0, 256))        # random bytes
        im = _decode(str(tmp_path), f"bad{trial}", data=bytes(b))
        assert im["pixels"].ndim == 3                                                # decoded to something or replaced by the 1x1 default
