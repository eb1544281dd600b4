"""Output stage behind the path: PNG screenshot of the 8-bit image and FrameStats JSON export.
CPU part: the oracle's mirror (stored-deflate PNG) through the same ABI. GPU part: the product's encoder (LZ77 + fixed Huffman,
Sub/Up filters) must decode to exactly the bytes lb_read_ldr returns, which in turn match the oracle's within one code value."""
import json
import os

import numpy as np
import pytest

import lumenrenderer_b200 as lr
from lumenrenderer_b200 import api, scenes
from conftest import read_png_rgba8


def _render(r, frames=1):
    r.load_scene(scenes.cornell_box())
    r.render_frames(frames)
    return r


def _check_stats(js, width, height, frames, stages=("raygen", "extend", "merge")):
    assert js["resolution"] == [width, height] and js["frame_id"] == frames
    assert js["counters"]["extend_rays"] >= width * height and js["counters"]["triangles"] == 32 and js["counters"]["lights"] == 2
    assert all(v >= 0 for v in js["times_us"].values()) and set(stages) <= set(js["times_us"])


def test_oracle_png_and_stats_json(oracle, tmp_path):
    with api.Renderer(oracle, lr.Settings(width=75, height=50, depth=2, restir=False)) as c:
        _render(c)
        path = os.path.join(tmp_path, "o.png")
        c.save_png(path)
        assert np.array_equal(read_png_rgba8(path), c.read_ldr())
        _check_stats(c.frame_stats_json(), 75, 50, 1)
        with pytest.raises(lr.LumenError):
            c.save_png(os.path.join(tmp_path, "no_such_dir", "x.png"))
        need = api.C.c_size_t(0)
        assert c.b.frame_stats_json(c._h, None, 0, api.C.byref(need)) == 0 and need.value > 50
        small = api.C.create_string_buffer(8)
        assert c.b.frame_stats_json(c._h, small, 8, None) == -1


@pytest.mark.gpu
def test_png_and_stats_json(oracle, tmp_path):
    st = lr.Settings(width=333, height=187, depth=3, restir=True)
    with lr.Renderer(st) as g, api.Renderer(oracle, st) as c:
        _render(g, 2); _render(c, 2)
        pg, pc = os.path.join(tmp_path, "g.png"), os.path.join(tmp_path, "c.png")
        g.save_png(pg); c.save_png(pc)
        img = read_png_rgba8(pg)
        assert np.array_equal(img, g.read_ldr()), "decoded screenshot differs from the output buffer"
        assert np.abs(img.astype(int) - read_png_rgba8(pc).astype(int)).max() <= 1
        assert os.path.getsize(pg) < 0.8 * os.path.getsize(pc), "the product's encoder should compress"
        js = g.frame_stats_json()
        _check_stats(js, 333, 187, 2, ("raygen", "extend", "shade", "shadow", "merge"))      # the fused device stages
        assert js["counters"]["kernel_launches"] > 0 and js["counters"]["visibility_rays"] > 0 and "restir_ris" in js["times_us"]
        assert json.dumps(js)
        with pytest.raises(lr.LumenError):
            g.save_png(os.path.join(tmp_path, "no_such_dir", "x.png"))
