"""The N > 1 path on CPU: two gloo ranks render disjoint sample streams with the ORACLE renderer, reduce their fp32
accumulation buffers with the same collective the GPU path uses (sum-reduce onto rank 0), and the result equals a
single-process render of the union of the streams."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lumenrenderer_b200 import api, scenes, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, FRAMES = 48, 40, 3


def _render_accum(settings, frames):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    r = api.Renderer(ge.oracle_bindings(), settings)
    r.load_scene(scenes.cornell_box())
    r.render_frames(frames)
    ptr, nbytes, n = r.accum_buffer()
    assert n == frames
    import ctypes
    acc = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(nbytes // 4,)).copy()
    r.close()
    return acc


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base = api.Settings(width=W, height=H, depth=3, restir=False)
    acc = torch.from_numpy(_render_accum(sharding.shard_settings(base, rank, world), FRAMES))
    sharding.reduce_accumulation(acc, dst=0)
    if rank == 0:
        np.save(out_path, (acc / (FRAMES * world)).numpy())
    dist.destroy_process_group()


def test_two_rank_sample_sharding_matches_single_process(tmp_path):
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(2, 29531 + os.getpid() % 200, out), nprocs=2, join=True)
    sharded = np.load(out)
    # single process rendering the union of both streams: frameCount 1,3,5,7,9,11 (stride 2 = the reference's own sequence)
    base = api.Settings(width=W, height=H, depth=3, restir=False, blend_output=True)
    single = _render_accum(base, 2 * FRAMES) / (2 * FRAMES)
    assert np.abs(single).sum() > 0
    assert np.allclose(sharded, single, rtol=1e-5, atol=1e-7)


def test_shard_plan():
    assert sharding.frame_counts(0, 1, 4) == [1, 3, 5, 7]                      # the reference's sequence (SURVEY hazard 10)
    streams = [sharding.frame_counts(r, 4, 5) for r in range(4)]
    flat = sorted(x for s in streams for x in s)
    assert flat == list(range(1, 41, 2))                                       # disjoint, gap-free cover of the odd numbers
    assert sharding.split_frames(64, 8) == [8] * 8 and sum(sharding.split_frames(10, 4)) == 10
    s = sharding.shard_settings(api.Settings(width=8, height=8), 3, 8)
    assert (s.first_frame_count, s.frame_count_stride, s.blend_output) == (6, 16, True)
    with pytest.raises(ValueError):
        sharding.shard_settings(api.Settings(), 8, 8)
    bands = sharding.band_partition(1440, 4)
    assert bands[0] == (0, 360, 0, 420) and bands[-1] == (1080, 1440, 1020, 1440)
    assert all(b[1] == n[0] for b, n in zip(bands, bands[1:]))
