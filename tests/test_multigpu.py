"""Multi-GPU inside the library (include/lumen_b200.h "multi-GPU inside the library", csrc/lb_multigpu.cpp): lb_band_settings / lb_shard_settings
(pure host arithmetic, checked here without a GPU against lumenrenderer_b200/sharding.py), lb_group_* (one process, n GPUs, ncclCommInitAll) and
lb_comm_* (one rank per process). The exchange itself needs NCCL and GPUs: the `gpu` tests below run on however many devices the box has — the
two-device cases skip on a single-GPU box and are run with `gpurun --gpus 2` (profiles/r02_multigpu_pytest.log). The same partitioning is
covered on CPU by the gloo tests (tests/test_sharding_gloo.py, tests/test_bands.py) with the oracle as the renderer."""
import ctypes
import os
import sys

import numpy as np
import pytest

import lumenrenderer_b200 as lr
from lumenrenderer_b200 import api, scenes, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_and_shard_settings_match_the_python_plan():
    for (w, h, ranks) in [(2560, 1440, 8), (2560, 1440, 4), (3840, 2160, 2), (48, 100, 2), (64, 288, 3), (640, 360, 1), (100, 77, 5)]:
        base = api.Settings(width=w, height=h, depth=4, restir=True)
        for rank in range(ranks):
            got, (y0, y1) = lr.band_settings(base, rank, ranks)
            want, (a, b, h0, h1) = sharding.band_settings(base, rank, ranks)
            assert got == want and (y0, y1) == (a, b), (w, h, ranks, rank)
            assert (got.band_row0 * w) % 256 == 0 and got.band_own_row0 == y0 and got.band_own_rows == y1 - y0
            assert lr.shard_settings(base, rank, ranks) == sharding.shard_settings(base, rank, ranks)
    with pytest.raises(lr.LumenError):
        lr.band_settings(api.Settings(width=64, height=64), 2, 2)
    with pytest.raises(lr.LumenError):
        lr.band_settings(api.Settings(width=64, height=3), 0, 4)                     # fewer rows than ranks


def test_group_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lr.LumenError) as e:
        lr.Group([0], api.Settings(width=32, height=32))
    assert "fallback" in str(e.value) or "NCCL" in str(e.value)


def test_own_rows_are_validated(oracle):
    with pytest.raises(api.LumenError):
        api.Renderer(oracle, api.Settings(width=64, height=16, band_row0=8, band_full_height=64, band_own_row0=4, band_own_rows=8))      # starts above the band
    with pytest.raises(api.LumenError):
        api.Renderer(oracle, api.Settings(width=64, height=16, band_row0=8, band_full_height=64, band_own_row0=16, band_own_rows=16))    # ends below it


def test_halo_rows_spawn_no_secondary_rays(oracle):
    """A band with owned rows traces fewer bounce / shadow rays than the same band without (the halo rows stop at the primary surface
    record), and its owned rows are unchanged."""
    base = api.Settings(width=64, height=256, depth=3, restir=True)
    st, (y0, y1, h0, h1) = sharding.band_settings(base, 1, 2)
    plain = api.Settings(**{**st.__dict__, "band_own_row0": 0, "band_own_rows": 0})
    out = []
    for s in (st, plain):
        with api.Renderer(oracle, s) as r:
            r.load_scene(scenes.cornell_box()); r.render_frames(2)
            out.append((r.read_hdr().copy(), r.frame_counters()))
    (a, ca), (b, cb) = out
    assert np.array_equal(a[y0 - h0:y1 - h0], b[y0 - h0:y1 - h0])
    assert ca["extend_rays"] < cb["extend_rays"] and ca["shadow_rays"] < cb["shadow_rays"] and ca["visibility_rays"] == cb["visibility_rays"]


# ------------------------------------------------------------------------------------------------------------------------------ GPU
def _devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["samples", "bands"])
def test_group_of_one_equals_the_plain_renderer(gpu, mode):
    scene = scenes.cornell_box()
    st = api.Settings(width=96, height=80, depth=3, restir=True)
    with lr.Group([0], st, mode) as g:
        g.load_scene(scene); g.render(3)
        if mode == "samples":
            g.reduce()
        img = g.read_hdr()
    with api.Renderer(gpu, api.Settings(**{**st.__dict__, "blend_output": mode == "samples"})) as r:
        r.load_scene(scene); r.render_frames(3)
        want = r.read_hdr()
    assert np.abs(want).sum() > 0 and np.array_equal(img, want)


@pytest.mark.gpu
def test_group_samples_on_two_gpus_equal_one_renderer_on_the_union_of_streams(gpu):
    if _devices() < 2:
        pytest.skip("needs two GPUs")
    scene = scenes.cornell_box()
    st = api.Settings(width=128, height=96, depth=3, restir=False)
    with lr.Group([0, 1], st, "samples") as g:
        g.load_scene(scene); g.render(3); g.reduce()
        img = g.read_hdr()
        g.render(1); g.reduce()                                # the members keep accumulating: a second reduce refines the same image (8 samples)
        img8 = g.read_hdr()
        g.reset(); g.render(1); g.reduce()                     # a new progressive image: the next two samples of the streams
        assert np.isfinite(g.read_hdr()).all()
    with api.Renderer(gpu, api.Settings(**{**st.__dict__, "blend_output": True})) as r:      # frameCount 1, 3, ..., 11: both streams interleaved
        r.load_scene(scene); r.render_frames(6)
        want = r.read_hdr()
        r.render_frames(2)
        want8 = r.read_hdr()
    assert np.allclose(img, want, rtol=1e-5, atol=1e-7) and np.abs(want).sum() > 0
    assert np.allclose(img8, want8, rtol=1e-5, atol=1e-7) and not np.array_equal(img8, img)


@pytest.mark.gpu
def test_group_bands_on_two_gpus_equal_the_full_frame(gpu):
    if _devices() < 2:
        pytest.skip("needs two GPUs")
    scene = scenes.cornell_box()
    st = api.Settings(width=256, height=320, depth=3, restir=True)
    with lr.Group([0, 1], st, "bands") as g, api.Renderer(gpu, st) as r:
        g.load_scene(scene); r.load_scene(scene)
        for frame in range(2):                                 # the first two frames after a history reset are bit-identical (tests/test_bands.py)
            g.render(1); r.render_frames(1)
            assert np.array_equal(g.read_hdr(), r.read_hdr()), f"frame {frame}"
        c = [m.frame_counters() for m in g.members]
        assert sum(x["extend_rays"] for x in c) < 1.6 * r.frame_counters()["extend_rays"]        # halo rows stop at the primary hit


def _comm_worker(rank, world, id_path, out_path, mode):
    sys.path.insert(0, ROOT)
    import time
    import torch
    import lumenrenderer_b200 as lr2
    from lumenrenderer_b200 import api as api2, scenes as scenes2
    torch.cuda.set_device(rank)
    if rank == 0:
        open(id_path + ".tmp", "wb").write(lr2.comm_unique_id()); os.rename(id_path + ".tmp", id_path)
    while not os.path.exists(id_path):
        time.sleep(0.01)
    uid = open(id_path, "rb").read()
    base = api2.Settings(width=128, height=192, depth=3, restir=(mode == "bands"))
    st = lr2.shard_settings(base, rank, world) if mode == "samples" else lr2.band_settings(base, rank, world)[0]
    st.device = rank
    r = lr2.Renderer(st)
    r.load_scene(scenes2.cornell_box())
    r.comm_init(uid, rank, world)
    r.render_frames(2)
    if mode == "samples":
        r.comm_reduce_accum(0, 2 * world)
        if rank == 0:
            np.save(out_path, r.read_hdr())
    else:
        full = torch.zeros((base.height, base.width, 4), device="cuda") if rank == 0 else None
        r.comm_gather_bands(0, full.data_ptr() if rank == 0 else 0)
        r.synchronize()
        if rank == 0:
            np.save(out_path, full.cpu().numpy())
    r.synchronize()
    r.comm_destroy(); r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["samples", "bands"])
def test_one_rank_per_process_on_two_gpus(gpu, tmp_path, mode):
    """lb_comm_*: two processes, one GPU each, the 128-byte NCCL id handed over through a file (any launcher works)."""
    if _devices() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "img.npy")
    mp.spawn(_comm_worker, args=(2, str(tmp_path / "nccl_id"), out, mode), nprocs=2, join=True)
    got = np.load(out)
    base = api.Settings(width=128, height=192, depth=3, restir=(mode == "bands"), blend_output=(mode == "samples"))
    with api.Renderer(gpu, base) as r:
        r.load_scene(scenes.cornell_box()); r.render_frames(4 if mode == "samples" else 2)
        want = r.read_hdr()
    if mode == "samples":
        assert np.allclose(got, want, rtol=1e-5, atol=1e-7)
    else:
        assert np.array_equal(got, want)
